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
SOURCES = ["edm_grid.cu", "edm_bias.cu", "edm_pair.cu", "edm_comm.cu"]
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
            extra = os.environ.get("EDM_NVCC_EXTRA", "").split()  # e.g. -DEDM_RUN_REDUX=0 for A/B builds
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
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
        cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


HOST_SRCS = ["edm.cpp", "grid.cpp", "gaussian_grid.cpp", "edm_bias.cpp", "edm_bias_py.cpp"]
HOST_LIB = os.path.join(LIBDIR, "libedm.so")
HOST_TEST = os.path.join(LIBDIR, "edm_host_test")
FIX_DRIVER = os.path.join(LIBDIR, "fix_driver_test")
EXCHANGE_TEST = os.path.join(LIBDIR, "exchange_test")
TEXT_IO_DRIVER = os.path.join(LIBDIR, "text_io_driver")
CXX = "/usr/bin/g++"
CXXFLAGS = ["-std=c++11", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-sign-compare"]


def build_host_tests(force=False):
    """libedm.so (the EDM:: C++ API over the C ABI), the restated reference tests, and a compile
    check of the two LAMMPS fixes against the mock LAMMPS headers."""
    build_lib()
    edm_dir = os.path.join(HERE, "edm")
    srcs = [os.path.join(edm_dir, s) for s in HOST_SRCS]
    hdrs = [os.path.join(edm_dir, h) for h in ("edm.h", "grid.h", "gaussian_grid.h", "edm_bias.h")]
    link = ["-L" + LIBDIR, "-ledm_b200", "-Wl,-rpath,$ORIGIN"]
    if force or _stale(HOST_LIB, srcs + hdrs + [LIB]):
        subprocess.check_call([CXX] + CXXFLAGS + ["-shared", "-o", HOST_LIB] + srcs + link)
    test_src = os.path.join(HERE, "tests_host", "edm_host_test.cpp")
    if force or _stale(HOST_TEST, [test_src, HOST_LIB]):
        subprocess.check_call([CXX] + CXXFLAGS + ["-o", HOST_TEST, test_src, "-L" + LIBDIR, "-ledm", "-ledm_b200",
                                                  "-Wl,-rpath,$ORIGIN"])
    # LAMMPS fixes: <edm/edm_bias.h> resolves through the package directory itself
    lmp = os.path.join(HERE, "lammps")
    for f in ("fix_edm.cpp", "fix_edm_pair.cpp"):
        obj = os.path.join(LIBDIR, f.replace(".cpp", ".o"))
        src = os.path.join(lmp, f)
        if force or _stale(obj, [src, os.path.join(lmp, f.replace(".cpp", ".h"))] + hdrs):
            subprocess.check_call([CXX] + CXXFLAGS + ["-I" + HERE, "-I" + os.path.join(lmp, "mock"), "-I" + lmp,
                                                      "-c", src, "-o", obj])
    # the fake MD loop that drives the two fixes on the GPU
    drv_src = os.path.join(HERE, "tests_host", "fix_driver_test.cpp")
    objs = [os.path.join(LIBDIR, "fix_edm.o"), os.path.join(LIBDIR, "fix_edm_pair.o")]
    if force or _stale(FIX_DRIVER, [drv_src, HOST_LIB] + objs):
        subprocess.check_call([CXX] + CXXFLAGS + ["-I" + HERE, "-I" + os.path.join(lmp, "mock"), "-I" + lmp, "-o",
                                                  FIX_DRIVER, drv_src] + objs +
                              ["-L" + LIBDIR, "-ledm", "-ledm_b200", "-Wl,-rpath,$ORIGIN"])
    # the multi-device hill exchange driven from C++ (ncclCommInitAll, one host thread per device)
    xt_src = os.path.join(HERE, "tests_host", "exchange_test.cpp")
    if force or _stale(EXCHANGE_TEST, [xt_src, HOST_LIB]):
        subprocess.check_call([CXX] + CXXFLAGS + ["-pthread", "-I" + HERE, "-o", EXCHANGE_TEST, xt_src,
                                                  "-L" + LIBDIR, "-ledm", "-ledm_b200", "-Wl,-rpath,$ORIGIN"])
    # the text-I/O driver against this repo's EDM:: classes (the same source is linked against the unmodified
    # reference by the checker's Makefile target `text` to produce the golden files)
    td_src = os.path.join(HERE, "tests_host", "text_io_driver.cpp")
    if force or _stale(TEXT_IO_DRIVER, [td_src, HOST_LIB]):
        subprocess.check_call([CXX] + CXXFLAGS + ["-I" + edm_dir, "-o", TEXT_IO_DRIVER, td_src, "-L" + LIBDIR, "-ledm",
                                                  "-ledm_b200", "-Wl,-rpath,$ORIGIN"])
    return HOST_TEST


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--host" in sys.argv:
        print(build_host_tests(force="--force" in sys.argv))
