"""Builds libehic_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the
working tree.  Invoked by ``__graft_entry__.build()`` and lazily by :mod:`._lib` when the
library is missing or older than its sources.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libehic_b200.so")
SOURCES = ["ehic_b200.cu", "kernels.cuh", "layout.hpp", "traj_cluster.cuh"]
HEADER = os.path.join(os.path.dirname(HERE), "include", "ehic_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    """Compile the library if needed; returns its path.  Safe when several ranks start at once:
    the build runs under an exclusive file lock into a temporary file that is renamed into place."""
    if not force and not _stale():
        return LIB
    import fcntl
    nvcc = nvcc_path()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libehic_b200.so (there is no CPU fallback)")
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():        # another process built it while we waited
                return LIB
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("EHIC_NVCC_EXTRA", "").split() + [os.path.join(CSRC, "ehic_b200.cu"), "-o", tmp]
            if verbose:
                print("[ensemble_hic_b200]", " ".join(cmd))
            subprocess.check_call(cmd, cwd=CSRC)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_fma_peak(force=False, verbose=False):
    """tools/libehic_fma_peak.so: the FMA issue-rate micro-benchmark bench.py uses as the roofline denominator of
    the FP64 / FP32 pipes (a measurement aid, kept out of the product library)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tools", "peak", "fma_peak.cu")
    out = os.path.join(root, "tools", "libehic_fma_peak.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = [nvcc_path()] + NVCC_FLAGS + [src, "-o", out + ".tmp.%d" % os.getpid()]
    if verbose:
        print("[ensemble_hic_b200]", " ".join(cmd))
    subprocess.check_call(cmd)
    os.replace(out + ".tmp.%d" % os.getpid(), out)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
