"""In-tree build of libtcmcmc.so (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# TC_LIBTCMCMC / TC_NVCC_EXTRA: development builds only (e.g. a -DTC_SUBPROF profiling variant beside the product library)
LIB = os.environ.get("TC_LIBTCMCMC") or os.path.join(HERE, "libtcmcmc.so")
SOURCES = ["tc_mcmc.cu"]
HEADERS = ["tc_device.cuh", "tc_warp.cuh", os.path.join("..", "..", "include", "tcmcmc.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtcmcmc.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(LIB):
        return True
    mt = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > mt for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libtcmcmc.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + os.environ.get("TC_NVCC_EXTRA", "").split() + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper; let nvcc pick the system host compiler
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd,
                         env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtcmcmc.so")
    with open(os.path.join(HERE, "libtcmcmc.ptxas.log"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
