// FP64 pipe micro-benchmarks on B200 (development aid): what instruction mixes can the FP64
// pipe sustain?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt_nr(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double a = x * y0;
    double e = fma(-a, y0, 1.0);
    double p = fma(0.375, e, 0.5);
    double ye = y0 * e;
    return fma(ye, p, y0);
}

#define NCH 8
// 1: v = v*a + b (two operands shared by all instructions)
__global__ void k_shared(double *out, int iters, double a, double b)
{
    double v[NCH];
    for (int c = 0; c < NCH; c++) v[c] = threadIdx.x + c;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) v[c] = fma(v[c], a, b);
    double s = 0; for (int c = 0; c < NCH; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2: three distinct register operands per DFMA
__global__ void k_distinct(double *out, int iters, double a, double b)
{
    double v[NCH], w[NCH], z[NCH];
    for (int c = 0; c < NCH; c++) { v[c] = threadIdx.x + c; w[c] = a + c * 1e-9; z[c] = b + c * 1e-9; }
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) { v[c] = fma(v[c], w[c], z[c]); }
    double s = 0; for (int c = 0; c < NCH; c++) s += v[c] + w[c] + z[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3: DMUL two distinct
__global__ void k_mul(double *out, int iters, double a, double b)
{
    double v[NCH], w[NCH];
    for (int c = 0; c < NCH; c++) { v[c] = 1.0 + 1e-9 * (threadIdx.x + c); w[c] = a + c * 1e-12; }
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) v[c] = v[c] * w[c];
    double s = 0; for (int c = 0; c < NCH; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 4: rsqrt_nr chains
__global__ void k_rsqrt(double *out, int iters, double a, double b)
{
    double v[NCH];
    for (int c = 0; c < NCH; c++) v[c] = 1.0 + 1e-3 * (threadIdx.x + c);
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) v[c] = rsqrt_nr(v[c]) + a;
    double s = 0; for (int c = 0; c < NCH; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 5: the fast-mode contact pair function (backward form), operands in registers
template <int NP>
__global__ void k_pair(double *out, int iters, double alpha, double dc)
{
    double ox[NP], oy[NP], oz[NP], a0[NP], a1[NP], a2[NP];
    const double xb = 0.1 * threadIdx.x, yb = 0.2, zb = 0.3, w = 1e-3;
    for (int c = 0; c < NP; c++) { ox[c] = 1.0 + c; oy[c] = 2.0 + 0.5 * c; oz[c] = 3.0 - 0.1 * c; a0[c] = a1[c] = a2[c] = 0; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < NP; c++) {
            double dx = ox[c] - xb, dy = oy[c] - yb, dz = oz[c] - zb;
            double s = fma(dz, dz, fma(dy, dy, dx * dx));
            double rd = rsqrt_nr(s);
            double tt = alpha * (dc - s * rd);
            double rq = rsqrt_nr(fma(tt, tt, 1.0));
            double f = w * ((rq * rq) * (rq * rd));
            a0[c] = fma(f, dx, a0[c]); a1[c] = fma(f, dy, a1[c]); a2[c] = fma(f, dz, a2[c]);
            ox[c] += 1e-9;   // keep the loop from being hoisted
        }
    }
    double s = 0; for (int c = 0; c < NP; c++) s += a0[c] + a1[c] + a2[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); f(); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 16 * 512);
    const int iters = 2000;
    for (int threads : {128, 256, 512}) for (int bps : {1, 2, 4}) {
        int blocks = sms * bps;
        double n = (double)blocks * threads;
        float t1 = timeit([&] { k_shared<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t2 = timeit([&] { k_distinct<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t3 = timeit([&] { k_mul<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t4 = timeit([&] { k_rsqrt<<<blocks, threads>>>(out, iters, 0.5, 1e-9); });
        float t5 = timeit([&] { k_pair<4><<<blocks, threads>>>(out, iters * 4, 10.0, 1.25); });
        float t6 = timeit([&] { k_pair<8><<<blocks, threads>>>(out, iters * 2, 10.0, 1.25); });
        double ops = n * iters * 64.0;
        printf("threads=%d blocks/SM=%d (warps/SM=%d): Ginstr-lanes/s  dfma_shared %.0f  dfma_3distinct %.0f  dmul %.0f | rsqrt_nr: %.1f G/s | pair4 %.1f Gpairs/s  pair8 %.1f Gpairs/s\n",
               threads, bps, threads / 32 * bps, ops / t1 * 1e-6, ops / t2 * 1e-6, ops / t3 * 1e-6,
               ops / t4 * 1e-6, n * iters * 16.0 / t5 * 1e-6, n * iters * 16.0 / t6 * 1e-6);
    }
    printf("reference: 64 lanes/clk/SM * %d SMs * %.3f GHz = %.0f G lanes/s\n", sms, clk * 1e-6, 64.0 * sms * clk * 1e-6);
    return 0;
}
