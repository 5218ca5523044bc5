"""Build libvlgp_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m vlgp_b200.build            # incremental
    python -m vlgp_b200.build --force

The shared library is the only native artefact of the package; it is git-ignored but travels to the GPU box with the
repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libvlgp_b200.so")
SOURCES = ["capi.cu", "hostpack.cpp", "comm.cu", "shmcomm.cu", "ichol.cu", "estep.cu", "estep_long.cu", "estep_seg.cu", "estep_seg_v_fast.cu", "estep_seg_v_gen.cu", "estep_seg_v_big.cu", "estep_seg_v_bigfast.cu", "estep_seg_v_fast32.cu", "estep_seg_v_bigfast32.cu", "mstep.cu", "hstep.cu", "hstep_dmma.cu", "hstep_wide.cu", "hstep_opt.cu", "p2p.cu", "regress.cu", "postcov.cu", "gpfa.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "vlgp_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _extra_flags():
    """Extra compiler flags from VLGP_NVCC_DEFINES (e.g. "-DVLGP_ESTEP_TWO_BINS -DVLGP_ESTEP_A2_FROM_SMEM": the A/B
    build options described in DESIGN.md section 8).  A change of flags rebuilds every object."""
    return [f for f in os.environ.get("VLGP_NVCC_DEFINES", "").split() if f.startswith("-D")]


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    extra = _extra_flags()
    stamp = os.path.join(OBJ, "flags.txt")
    prev = open(stamp).read() if os.path.exists(stamp) else ""
    if prev != " ".join(extra):
        force = True
        with open(stamp, "w") as f:
            f.write(" ".join(extra))
    hdr_m = _deps_mtime()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(LIB):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"])
    build_fastpack(force)
    return LIB


def build_fastpack(force: bool = False):
    """Optional CPython helper (host pointer tables); skipped silently when Python.h or gcc is missing."""
    import sysconfig

    src = os.path.join(CSRC, "fastpack.c")
    out = os.path.join(HERE, "_fastpack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    inc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(inc, "Python.h")) or not shutil.which("gcc"):
        return None
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I" + inc, "-o", out, src], capture_output=True, text=True)
        if r.returncode != 0:
            return None
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
