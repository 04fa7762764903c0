// fma_peak.cu -- FP64 / FP32 FMA issue-rate micro-benchmark: the roofline denominators of bench.py for the
// non-tensor pipes (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only).  Measurement aid,
// deliberately NOT part of the product library: built into tools/libehic_fma_peak.so by ensemble_hic_b200/build.py.
#include <cuda_runtime.h>

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T a, T b)
{
    T v0 = (T)threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            v0 = v0 * a + b; v1 = v1 * a + b; v2 = v2 * a + b; v3 = v3 * a + b;
            v4 = v4 * a + b; v5 = v5 * a + b; v6 = v6 * a + b; v7 = v7 * a + b;
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

// returns 0 on success; *tflops = achieved TFLOP/s (2 flop per FMA) over `iters` launches after 2 warm-up launches
extern "C" int fma_peak(int device, int fp64, int iters, double *tflops, double *ms_per_launch)
{
    if (!tflops || cudaSetDevice(device) != cudaSuccess) return 1;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 2;
    const int blocks = sms * 8, threads = 256, inner = 4096;
    void *buf = nullptr;
    if (cudaMalloc(&buf, (size_t)blocks * threads * 8) != cudaSuccess) return 3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (iters < 1) iters = 1;
    for (int w = 0; w < 2 + iters; w++) {
        if (w == 2) cudaEventRecord(e0);
        if (fp64) fma_peak_kernel<double><<<blocks, threads>>>((double *)buf, inner, 1.0000001, 1e-9);
        else fma_peak_kernel<float><<<blocks, threads>>>((float *)buf, inner, 1.0000001f, 1e-9f);
    }
    cudaEventRecord(e1);
    int rc = cudaEventSynchronize(e1) == cudaSuccess ? 0 : 4;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * inner * (double)blocks * threads * iters;
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (ms_per_launch) *ms_per_launch = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    return rc;
}
