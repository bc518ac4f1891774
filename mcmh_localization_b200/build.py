"""Builds libmcl.so (hand-written sm_100a CUDA + the C ABI of include/mcl.h) in-tree with nvcc.

    python -m mcmh_localization_b200.build          # incremental
    python -m mcmh_localization_b200.build --force

The shared object lands next to this file so that it travels with the repo snapshot; it is
git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libmcl.so")
SOURCES = ["mcl_core.cu", "likelihood.cu", "motion.cu", "mh_softmax.cu", "resample.cu", "estimate.cu",
           "init_misc.cu", "filter.cu", "amh.cu", "kld.cu", "raycast.cu", "edt.cu", "fused.cu", "tail.cu", "alt.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--fmad=true", "-Xptxas", "-v"]


def _deps():
    d = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "seqsum.cuh"), os.path.join(os.path.dirname(HERE), "include", "mcl.h"),
         os.path.abspath(__file__)]
    return max(os.path.getmtime(p) for p in d)


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    dep_t = _deps()
    objs, rebuilt = [], False
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), dep_t):
            log = open(o + ".log", "w")
            procs.append((src, log, subprocess.Popen([NVCC] + FLAGS + ["-c", s, "-o", o], stdout=log,
                                                     stderr=subprocess.STDOUT)))
            rebuilt = True
    for src, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            sys.stderr.write(open(os.path.join(OBJ, src.replace(".cu", ".o")) + ".log").read())
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose:
            sys.stdout.write(open(os.path.join(OBJ, src.replace(".cu", ".o")) + ".log").read())
    if rebuilt or not os.path.exists(LIB):
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
