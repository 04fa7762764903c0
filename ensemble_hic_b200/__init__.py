"""ensemble_hic_b200 -- the posterior hot path of ensemble_hic on NVIDIA B200 (sm_100a).

Drop-in for the reference's native kernels and the PDF classes that call them:
same module, class and function names as ``ensemble_hic`` for this path; the bodies
call hand-written CUDA kernels through the C ABI in ``include/ehic_b200.h``.
There is no CPU fallback: without a B200 the compute calls raise.
"""
__version__ = "0.1.0"
