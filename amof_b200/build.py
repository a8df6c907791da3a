"""
Builds ``amof_b200/libamofb.so`` (the C-ABI CUDA library of include/amofb.h) in-tree with nvcc for sm_100a.

    python -m amof_b200.build [--force]

Flags that matter for parity: ``-fmad=false`` (device) and ``-ffp-contract=off`` (host) keep every
bin-deciding fp64 expression un-fused, in the operation order the oracle pins (oracle/amof_oracle.c P1-P8).
nvcc cross-compiles without a GPU, so this runs in the build container; the .so travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libamofb.so")
SOURCES = ["amofb.cu", "xyz_parse.cpp"]
DEPS = ["amofb.cu", "xyz_parse.cpp", "common.cuh", "prep.cuh", "pair.cuh", "pair_tiled.cuh", "neigh.cuh", "neigh_host.inl", "bad.cuh", "msd.cuh", "bad_host.inl", "msd_host.inl",
        os.path.join("..", "..", "include", "amofb.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libamofb.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2", "-shared",
           "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    extra = os.environ.get("AMOFB_NVCC_FLAGS", "").split()
    cmd[1:1] = extra
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
