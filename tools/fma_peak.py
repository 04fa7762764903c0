"""ctypes face of tools/libehic_fma_peak.so (tools/peak/fma_peak.cu): FP64 / FP32 FMA issue-rate micro-benchmark,
the roofline denominator bench.py uses for the non-tensor pipes (not in MEASURED_PEAKS.json)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def fma_peak(fp64=True, iters=20, device=0):
    """(TFLOP/s, ms per launch)"""
    from ensemble_hic_b200 import build
    L = C.CDLL(build.build_fma_peak())
    L.fma_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    t, ms = C.c_double(), C.c_double()
    rc = L.fma_peak(int(device), 1 if fp64 else 0, int(iters), C.byref(t), C.byref(ms))
    if rc:
        raise RuntimeError("fma_peak failed with code %d" % rc)
    return t.value, ms.value


if __name__ == "__main__":
    for f in (True, False):
        print("fp64" if f else "fp32", fma_peak(f))
