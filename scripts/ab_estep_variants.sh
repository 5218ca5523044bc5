#!/bin/bash
# A/B of the three forms of the segment E-step's rate pass (profiles/sass_r1_estep_rate_pass.txt).
#
#   scripts/ab_estep_variants.sh build     # here (no GPU): three libraries under vlgp_b200/variants/ (git-ignored)
#   scripts/ab_estep_variants.sh bench     # on the GPU box: bench.py once per library, JSON lines to gpurun_out/
#
# Typical round trip:  scripts/ab_estep_variants.sh build && \
#   gpurun --timeout 900 -- 'bash scripts/ab_estep_variants.sh bench'
# The default library (vlgp_b200/libvlgp_b200.so) is rebuilt with the default flags at the end of `build`.
set -e
cd "$(dirname "$0")/.."
VAR=vlgp_b200/variants
case "$1" in
build)
    mkdir -p $VAR
    for v in "a2_from_smem:-DVLGP_ESTEP_A2_FROM_SMEM" "two_bins:-DVLGP_ESTEP_TWO_BINS" "default:"; do
        name=${v%%:*}; flags=${v#*:}
        echo "== $name ($flags)"
        VLGP_NVCC_DEFINES="$flags" python -m vlgp_b200.build > /dev/null
        cp vlgp_b200/libvlgp_b200.so $VAR/libvlgp_b200_$name.so
    done
    ls -la $VAR
    ;;
bench)
    mkdir -p gpurun_out
    for name in a2_from_smem default two_bins; do
        lib=$PWD/$VAR/libvlgp_b200_$name.so
        [ -f "$lib" ] || { echo "missing $lib (run: $0 build)"; exit 1; }
        echo "== $name"
        VLGP_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu | tee gpurun_out/ab_estep_$name.json | \
            python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], 'EM-iter/s; E-step', d['roofline']['ms_per_launch'], 'ms per launch; e2e', d['e2e']['value'])"
    done
    ;;
*)
    sed -n 2,10p "$0"
    ;;
esac
