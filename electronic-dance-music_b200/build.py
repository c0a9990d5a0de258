"""Builds libedm_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python electronic-dance-music_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands in electronic-dance-music_b200/lib/ (git-ignored,
but it travels to the GPU box with the snapshot).  Host code is compiled by the system g++ with FMA
contraction off: the McGDP tables and grid geometry are evaluated on the host and must come out
bit-identical to the reference's.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libedm_b200.so")
SOURCES = ["edm_grid.cu", "edm_bias.cu", "edm_pair.cu"]
HEADERS = ["edm_device.cuh", "edm_host.h", os.path.join("..", "..", "include", "edm_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "--expt-relaxed-constexpr",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            results = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs))
        for cmd, r in zip(jobs, results):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
