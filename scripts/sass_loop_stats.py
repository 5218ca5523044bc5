#!/usr/bin/env python
"""Static look at the hot loops of one instantiation of the segment E-step kernel (no GPU needed).

    python scripts/sass_loop_stats.py [--L 5] [--nbmax 2] [--fast 1] [-D VLGP_ESTEP_TWO_BINS ...] [--maxreg-blocks]

Compiles ONLY estep_seg_kernel<L, NBMAX, FAST> for sm_100a into a temporary directory (a few seconds instead of the
minutes the full translation unit takes), prints ptxas' register / spill report and, for every loop of the SASS that
carries at least 50 DFMA, its instruction mix, the number and width of its shared-memory loads and the longest run of
consecutive DFMA that write the register they read -- i.e. whether two Horner chains that are interleaved in the source
are still interleaved after ptxas (DESIGN.md section 8 quotes these numbers).
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vlgp_b200", "csrc")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=5)
    ap.add_argument("--nbmax", type=int, default=2)
    ap.add_argument("--fast", type=int, default=1)
    ap.add_argument("-D", action="append", default=[], help="extra defines, e.g. -D VLGP_ESTEP_TWO_BINS")
    ap.add_argument("--min-dfma", type=int, default=50)
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "one.cu")
        with open(src, "w") as f:
            f.write('#include "estep_seg_impl.cuh"\nnamespace segk {\ntemplate __global__ void '
                    "estep_seg_kernel<%d, %d, %s>(SegArgs);\n}\n" % (args.L, args.nbmax, "true" if args.fast else "false"))
        obj = os.path.join(tmp, "one.o")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "--expt-relaxed-constexpr", "-Xptxas", "-v", "-I", CSRC] + ["-D" + d for d in args.D] + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        for line in r.stderr.splitlines():
            if "registers" in line or "spill" in line:
                print(line.strip())
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ins = []
    for line in sass.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("%d SASS instructions" % len(ins))
    for addr, text in ins:
        if "BRA" not in text:
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m or int(m.group(1), 16) >= addr:
            continue
        tgt = int(m.group(1), 16)
        body = [x for a, x in ins if tgt <= a <= addr]
        ops = Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for x in body)
        if ops["DFMA"] < args.min_dfma or len(body) > 400:
            continue
        run = best = 0
        prev = None
        for x in body:
            if x.startswith("DFMA"):
                d = x.split()[1].rstrip(",")
                run = run + 1 if d == prev else 1
                prev, best = d, max(best, run)
        lds = Counter(x.split()[0] for x in body if x.startswith("LDS"))
        fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"] + ops["DSETP"]
        waves = sum(n * (4 if k.endswith(".128") else 2 if k.endswith(".64") else 1) for k, n in lds.items())
        print("loop %#x-%#x: %d instructions, FP64 pipe %d (= %d issue cycles), shared loads %s = %d wavefronts per warp "
              "(x4 sub-partitions = %d SM-cycles), longest dependent DFMA run %d\n    mix %s"
              % (tgt, addr, len(body), fp64, 2 * fp64, dict(lds), waves, 4 * waves, best, dict(ops.most_common(10))))


if __name__ == "__main__":
    main()
