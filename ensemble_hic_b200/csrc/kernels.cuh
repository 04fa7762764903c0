// kernels.cuh -- sm_100a kernels of the ensemble_hic posterior hot path.
//
// Not a dense contraction: no tensor cores.  The large configurations are bound by the
// FP64 pipe (64 DFMA lanes / clk / SM), the small ones by launch latency; see DESIGN.md.
//
// Working layout of the coordinates inside the kernels ("xs"): [R][S][3][Np] FP64, SoA per
// structure, Np = N rounded up to 32.  A warp whose lanes are consecutive beads of one
// structure then reads 256 contiguous bytes per component, and a warp that gathers the
// (nearly identical) partner index reads one or two sectors.  The public layout
// [R][S][N][3] (forward_models.py:78) is converted by pack_kernel (4.8 MB per replica at
// N=2000, S=100: noise next to the pair arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>

namespace ehic {

struct DevModel {
    int N, S, Np, n_slices;
    long long M, tot, wstride;   // wstride = tot + tail pad: per-replica stride of the weight buffer
    // pair view
    const int *pi, *pj, *slot_a, *slot_b;
    const double *pdc, *pcnt;
    const long long *perm;
    // row view
    const long long *slice_off;
    const int *slice_w;
    const int *ent_j;
    const double *ent_dc;
    const int *row_bead;    // [n_slices*32] lane slot -> bead (-1 = none)
    // per bead
    const double *radii;    // [N]
    const double *bb_ul;    // [N] upper limit of bond (b, b+1); +inf where there is no bond
    const double *bb_ll;    // [N] lower limit of bond (b, b+1); 0    where there is no bond
    double alpha, k_nb, k_bb, sph_R, sph_k, sph_cx, sph_cy, sph_cz;
    int sphere_active;
};

// ---------------------------------------------------------------------------------------
// arithmetic helpers

// Full-precision 1/sqrt(x) for normal x > 0: MUFU.RSQ64H seed (~2^-22) + one third-order
// step (error ~ (5/16) e^3 < 2^-60).  5 FP64 instructions, no slow path; x = 0 gives inf/NaN
// exactly where the reference divides by zero.
#ifndef EHIC_RSQRT_ORDER
#define EHIC_RSQRT_ORDER 3
#endif
__device__ __forceinline__ double rsqrt_nr(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double a = x * y0;
    double e = fma(-a, y0, 1.0);
#if EHIC_RSQRT_ORDER == 3
    double p = fma(0.375, e, 0.5);
    double ye = y0 * e;
    return fma(ye, p, y0);
#else   // second order: 4 FP64 instructions, relative error ~1.5 * seed_error^2 (~1e-13)
    double h = 0.5 * y0;
    return fma(h, e, y0);
#endif
}

__device__ __forceinline__ double warp_sum(double v)
{   // fixed butterfly order: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one elected thread moves a
// whole contiguous stage (up to 16 KB) global -> shared instead of hundreds of per-thread cp.async.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// FP32 mode: a coordinate is carried as two floats (hi = float(x), lo = float(x - hi)) in the 8 bytes of
// the double it replaces, so that differences x_i - x_j = (hi_i - hi_j) + (lo_i - lo_j) keep a relative
// error of ~1e-7 however far the beads are from the origin.
__device__ __forceinline__ double split_hilo(double x)
{
    const float hi = __double2float_rn(x);
    const float lo = __double2float_rn(x - (double)hi);
    return __hiloint2double(__float_as_int(lo), __float_as_int(hi));      // memory order: (hi, lo)
}
__device__ __forceinline__ float rsqrt_f32(float x)
{   // MUFU.RSQ + one Newton step: relative error ~1e-7
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    const float e = fmaf(-x * y, y, 1.0f);
    return fmaf(0.5f * y, e, y);
}
__device__ __forceinline__ float rsqrt_f32_raw(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_f32_raw(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------
// x[R][S][N][3] -> xs[R][S][3][Np]  and  xt[R][nkt][Np][KT][3]
//
// xt ("tile-AoS") is the gather kernel's copy: one record holds bead o of the KT structures of
// one structure tile (KT*24 B, contiguous), so a partner costs one address computation and
// KT*3/2 LDG.128 instead of KT*3 separately addressed LDG.64.
//
// With `lanes` set, xt is the LANE layout of the sparse path instead: xg[R][nkg][Np][3][32], one
// structure per lane (KT = 32, nkt = nkg), so a warp reads one bead of 32 structures as three
// fully coalesced 256-byte rows.
__global__ void pack_kernel(const double *__restrict__ x, double *__restrict__ xs, double *__restrict__ xt,
                            int N, int Np, int S, int KT, int nkt, long long rows /* R*S */, int mode)
{   // mode 0: tile-AoS doubles; 1: lane layout; 2: tile-AoS (hi, lo) float pairs of the FP32 mode
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = rows * Np;
    if (t >= total) return;
    long long row = t / Np;
    int i = (int)(t - row * Np);
    double a = 0.0, b = 0.0, c = 0.0;
    if (i < N) {
        const double *p = x + (row * N + i) * 3;
        a = p[0]; b = p[1]; c = p[2];
    }
    double *q = xs + row * 3 * (long long)Np + i;
    q[0] = a; q[Np] = b; q[2 * (long long)Np] = c;
    if (xt && i < N) {
        long long r = row / S;
        int k = (int)(row - r * S);
        if (mode == 1) {
            double *u = xt + ((r * nkt + k / 32) * Np + i) * 96 + (k % 32);
            u[0] = a; u[32] = b; u[64] = c;
        } else {
            double *u = xt + (((r * nkt + k / KT) * Np + i) * KT + (k % KT)) * 3;
            if (mode == 2) { a = split_hilo(a); b = split_hilo(b); c = split_hilo(c); }
            u[0] = a; u[1] = b; u[2] = c;
        }
    }
}

// Lane layout of the sparse path: x[R][S][N][3] -> xs[R][S][3][Np] and xg[R][nkg][Np][3][32] (structure <-> lane).
// A 32-structure x 32-bead tile goes through shared memory so that both the reads (96 consecutive
// doubles per structure) and the writes (96 consecutive doubles per bead) are coalesced; padding
// structures and beads are written as zeros.
__global__ void __launch_bounds__(256)
pack_lane_kernel(const double *__restrict__ x, double *__restrict__ xs, double *__restrict__ xg,
                 int N, int Np, int S, int nkg, int f32 /* xg holds (hi, lo) float pairs: FP32 mode */)
{
    __shared__ double t[32][97];
    const int i0 = blockIdx.x * 32, kg = blockIdx.y, r = blockIdx.z;
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
        const int k = idx / 96, q = idx - k * 96;
        const int kgl = kg * 32 + k;
        const long long col = 3LL * i0 + q;
        double v = 0.0;
        if (kgl < S && col < 3LL * N) v = x[((long long)r * S + kgl) * N * 3 + col];
        t[k][q] = v;
    }
    __syncthreads();
    double *g = xg + (((long long)r * nkg + kg) * Np + i0) * 96;
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
        const double v = t[idx & 31][idx >> 5];
        g[idx] = f32 ? split_hilo(v) : v;
    }
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
        const int b = idx & 31, kc = idx >> 5, k = kc / 3, c = kc - 3 * k;
        const int kgl = kg * 32 + k;
        if (kgl < S) xs[(((long long)r * S + kgl) * 3 + c) * Np + i0 + b] = t[k][3 * b + c];
    }
}

// xs[R][S][3][Np] -> x[R][S][N][3]
__global__ void unpack_kernel(const double *__restrict__ xs, double *__restrict__ x,
                              int N, int Np, long long rows)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = rows * N;
    if (t >= total) return;
    long long row = t / N;
    int i = (int)(t - row * N);
    const double *q = xs + row * 3 * (long long)Np + i;
    double *p = x + (row * N + i) * 3;
    p[0] = q[0]; p[1] = q[Np]; p[2] = q[2 * (long long)Np];
}

// ---------------------------------------------------------------------------------------
// Forward kernel: mock data md[m] = norm * sum_k s_km (evaluate_FWM_pureC.pxd:29-39), the
// Poisson weight w_m = 0.5 (1 - c_m/md_m) norm alpha (likelihoods_c.pyx:24-25,69) scattered
// into both row slots of the pair, and the per-block partial of
// log L = sum_m (c_m log md_m - md_m) (error_models.py:97).
//
// thread <-> pair m of the (i, j)-sorted list; the reduction over the ensemble is a serial
// loop over k in the reference's own order (deterministic; bit-exact in EXACT mode).
template <bool EXACT>
__global__ void __launch_bounds__(256)
forward_kernel(DevModel dm, const double *__restrict__ xs, const double *__restrict__ norm,
               double *__restrict__ md_out, double *__restrict__ wbuf,
               double *__restrict__ partA, int want_logl)
{
    const int r = blockIdx.y;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long plane = 3LL * dm.Np;
    const double *xr = xs + (long long)r * dm.S * plane;
    const double nrm = norm[r];
    double term = 0.0;
    if (m < dm.M) {
        const int i = dm.pi[m], j = dm.pj[m];
        const double dc = dm.pdc[m], c = dm.pcnt[m];
        const double alpha = dm.alpha;
        const double *pa = xr + i, *pb = xr + j;
        double md;
        if (EXACT) {
            const double aa = __dmul_rn(alpha, alpha);
            md = 0.0;
            for (int k = 0; k < dm.S; k++, pa += plane, pb += plane) {
                double d0 = __dsub_rn(pa[0], pb[0]);
                double d1 = __dsub_rn(pa[dm.Np], pb[dm.Np]);
                double d2 = __dsub_rn(pa[2 * dm.Np], pb[2 * dm.Np]);
                double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
                double d = __dsqrt_rn(s);
                double u = __dsub_rn(dc, d);
                double q = __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(__dmul_rn(aa, u), u)));
                double v = __dadd_rn(__ddiv_rn(__dmul_rn(alpha, u), q), 1.0);
                md = __dadd_rn(md, __dmul_rn(__dmul_rn(v, 0.5), nrm));
            }
        } else {
            double acc0 = 0.0, acc1 = 0.0;
            int k = 0;
            for (; k + 1 < dm.S; k += 2, pa += 2 * plane, pb += 2 * plane) {
                double ax = pa[0] - pb[0], ay = pa[dm.Np] - pb[dm.Np], az = pa[2 * dm.Np] - pb[2 * dm.Np];
                double bx = pa[plane] - pb[plane], by = pa[plane + dm.Np] - pb[plane + dm.Np],
                       bz = pa[plane + 2 * dm.Np] - pb[plane + 2 * dm.Np];
                double sa = fma(az, az, fma(ay, ay, ax * ax));
                double sb = fma(bz, bz, fma(by, by, bx * bx));
                double ra = rsqrt_nr(sa), rb = rsqrt_nr(sb);
                double ta = alpha * (dc - sa * ra), tb = alpha * (dc - sb * rb);
                double qa = rsqrt_nr(fma(ta, ta, 1.0)), qb = rsqrt_nr(fma(tb, tb, 1.0));
                acc0 += fma(ta, qa, 1.0);
                acc1 += fma(tb, qb, 1.0);
            }
            if (k < dm.S) {
                double ax = pa[0] - pb[0], ay = pa[dm.Np] - pb[dm.Np], az = pa[2 * dm.Np] - pb[2 * dm.Np];
                double sa = fma(az, az, fma(ay, ay, ax * ax));
                double ra = rsqrt_nr(sa);
                double ta = alpha * (dc - sa * ra);
                double qa = rsqrt_nr(fma(ta, ta, 1.0));
                acc0 += fma(ta, qa, 1.0);
            }
            md = (acc0 + acc1) * (0.5 * nrm);
        }
        if (md_out) md_out[(long long)r * dm.M + dm.perm[m]] = md;
        // numerator of f in likelihoods_c.pyx:69: ((0.5 * em) * norm) * alpha, em = 1 - c/md
        double em = __dsub_rn(1.0, __ddiv_rn(c, md));
        double w = __dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), alpha);
        if (wbuf) {
            double *wr = wbuf + (long long)r * dm.wstride;
            wr[dm.slot_a[m]] = w;
            wr[dm.slot_b[m]] = w;
        }
        if (want_logl) term = c * log(md) - md;
    }
    if (want_logl) {
        __shared__ double sh[8];
        double v = warp_sum(term);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) t += sh[q];
            partA[(long long)r * gridDim.x + blockIdx.x] = t;
        }
    }
}

// One tile-AoS record (KT*3 doubles, 16-byte aligned when KT is even) -> registers.
template <int KT>
__device__ __forceinline__ void load_record(const double *__restrict__ p, double (&v)[KT * 3])
{
    if constexpr (KT % 2 == 0) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int u = 0; u < KT * 3 / 2; u++) {
            double2 w = q[u];
            v[2 * u] = w.x; v[2 * u + 1] = w.y;
        }
    } else {
#pragma unroll
        for (int u = 0; u < KT * 3; u++) v[u] = p[u];
    }
}

// ---------------------------------------------------------------------------------------
// Forward kernel of the row path, fast mode: same mapping as forward_kernel (thread <-> pair of the
// (i, j)-sorted list, serial over the ensemble), but both beads are read as tile-AoS records:
// 12 LDG.128 per KT structures instead of 6 LDG.64 per structure.  On sparse lists, where every
// partner is its own cache line, that is KT times fewer sectors from L2 (the SoA version is bound
// by L2 sector traffic there, profiles/r01_sparse_rowpath_E.txt).
template <int KT>
__global__ void __launch_bounds__(256)
forward_rec_kernel(DevModel dm, const double *__restrict__ xt, const double *__restrict__ norm,
                   double *__restrict__ md_out, double *__restrict__ wbuf,
                   double *__restrict__ partA, int want_logl, int nkt)
{
    const int r = blockIdx.y;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double nrm = norm[r];
    double term = 0.0;
    if (m < dm.M) {
        const int i = dm.pi[m], j = dm.pj[m];
        const double dc = dm.pdc[m], c = dm.pcnt[m];
        const double alpha = dm.alpha;
        const double *ra = xt + ((long long)r * nkt * dm.Np + i) * (KT * 3);
        const double *rb = xt + ((long long)r * nkt * dm.Np + j) * (KT * 3);
        const long long step = (long long)dm.Np * (KT * 3);
        double acc[KT];
#pragma unroll
        for (int t = 0; t < KT; t++) acc[t] = 0.0;
        const int nfull = dm.S / KT;
        for (int kt = 0; kt < nfull; kt++, ra += step, rb += step) {
            double a[KT * 3], b[KT * 3];
            load_record<KT>(ra, a); load_record<KT>(rb, b);
#pragma unroll
            for (int t = 0; t < KT; t++) {
                double dx = a[3 * t] - b[3 * t], dy = a[3 * t + 1] - b[3 * t + 1], dz = a[3 * t + 2] - b[3 * t + 2];
                double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                double rd = rsqrt_nr(s2);
                double tt = alpha * (dc - s2 * rd);
                double rq = rsqrt_nr(fma(tt, tt, 1.0));
                acc[t] += fma(tt, rq, 1.0);
            }
        }
        if (nfull < nkt) {
            double a[KT * 3], b[KT * 3];
            load_record<KT>(ra, a); load_record<KT>(rb, b);
            const int nvalid = dm.S - nfull * KT;
#pragma unroll
            for (int t = 0; t < KT; t++) {
                if (t < nvalid) {
                    double dx = a[3 * t] - b[3 * t], dy = a[3 * t + 1] - b[3 * t + 1], dz = a[3 * t + 2] - b[3 * t + 2];
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    double rd = rsqrt_nr(s2);
                    double tt = alpha * (dc - s2 * rd);
                    double rq = rsqrt_nr(fma(tt, tt, 1.0));
                    acc[t] += fma(tt, rq, 1.0);
                }
            }
        }
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < KT; t++) sum += acc[t];
        const double md = sum * (0.5 * nrm);
        if (md_out) md_out[(long long)r * dm.M + dm.perm[m]] = md;
        double em = __dsub_rn(1.0, __ddiv_rn(c, md));
        double w = __dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), alpha);
        if (wbuf) {
            double *wr = wbuf + (long long)r * dm.wstride;
            wr[dm.slot_a[m]] = w;
            wr[dm.slot_b[m]] = w;
        }
        if (want_logl) term = c * log(md) - md;
    }
    if (want_logl) {
        __shared__ double sh[8];
        double v = warp_sum(term);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) t += sh[q];
            partA[(long long)r * gridDim.x + blockIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Bounding boxes of the 32-bead slices of every structure: bbox[R*S][n_slices][6] = (lo, hi).
// Consecutive beads of a chain are neighbours in space, so the all-pairs sweep of the prior kernel
// can discard whole 32 x 32 slice pairs whose boxes are farther apart than the largest contact
// range (a warp-uniform test) -- the same pruning a cell grid gives the reference (c/nblist.c),
// without sorting and without changing which pairs contribute or in which order.
__global__ void __launch_bounds__(32)
bbox_kernel(const double *__restrict__ xs, double *__restrict__ bbox, int N, int Np, int n_slices)
{
    const long long rk = blockIdx.x;              // R*S can exceed the 65535 limit of grid.y: structures on grid.x
    const int sl = blockIdx.y, lane = threadIdx.x;
    const int b = min(sl * 32 + lane, N - 1);
    const double *xk = xs + rk * 3LL * Np;
    double lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double v = xk[c * Np + b];
        double mn = v, mx = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        lo[c] = mn; hi[c] = mx;
    }
    if (lane == 0) {
        double *o = bbox + (rk * n_slices + sl) * 6;
        o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
    }
}

// ---------------------------------------------------------------------------------------
// Prior kernel (north-star kernel 2): repulsive nonbonded term as a tiled all-pairs gather,
// with the backbone and sphere terms fused into the same pass.
//
// thread <-> (replica r, structure k, bead b); a CTA is 256 consecutive beads of ONE structure.
// Partner beads are staged through shared memory in tiles of 256 as (x, y), (z, radius) pairs
// and read back with warp-uniform (broadcast) LDS.128.  Every pair is visited from both ends
// (gather form): no atomics, fixed ascending-partner summation order -- the order in which
// forcefield_c.pyx:47-57 accumulates into res[b].
// Output: gprior[R][S][N][3] = beta * dE_nb/dx + d(-log p_bb)/dx + d(-log p_sphere)/dx and
// per-CTA energy partials partP[(r*S+k)*gridDim.x + blockIdx.x][3] = (sum a^4, sum bb, sum sph).
#define EHIC_PRIOR_MAX_THREADS 256
#define EHIC_PRIOR_TILE 512
#define EHIC_CELL_MAX 4096

// Structures are routed per evaluation: compact ones (grid of contact-range cells fits EHIC_CELL_MAX)
// go to cell_nonbonded_kernel, extended ones to the box-pruned all-pairs sweep of prior_kernel.
// Both kernels call this with the same slice boxes, so they always agree.  Executed by warp 0.
__device__ __forceinline__ bool structure_fits_cells(const double *__restrict__ bbk, int n_slices, double r_max,
                                                     int lane, double (&lo)[3], double (&hi)[3])
{
#pragma unroll
    for (int c = 0; c < 3; c++) { lo[c] = 1e300; hi[c] = -1e300; }
    for (int q = lane; q < n_slices; q += 32)
#pragma unroll
        for (int c = 0; c < 3; c++) { lo[c] = fmin(lo[c], bbk[q * 6 + c]); hi[c] = fmax(hi[c], bbk[q * 6 + 3 + c]); }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    const double inv = 1.0 / (2.0 * r_max * (1.0 + 1e-9));
    double cells = 1.0;
#pragma unroll
    for (int c = 0; c < 3; c++) cells *= floor((hi[c] - lo[c]) * inv) + 1.0;
    return cells <= (double)EHIC_CELL_MAX;      // false for NaN extents
}

// Backbone and sphere restraints of one bead (IEEE arithmetic and the reference's operation order in both
// modes); adds to (g0, g1, g2) and accumulates the squared violations.  Shared by prior_kernel and
// verlet_force_kernel.
template <class XAt>
__device__ __forceinline__ void bonded_terms_t(const DevModel &dm, XAt xat, int b, bool bead_ok, int bb,
                                             double xb0, double xb1, double xb2, double rb, int terms,
                                             double &g0, double &g1, double &g2, double &e_bb, double &e_sp)
{
    // backbone: bonds (b-1, b) then (b, b+1), the reference's visiting order (backbone_prior_c.pyx:26-42)
    if (terms & 4) {
        const bool has_prev = bead_ok && b > 0 && !isinf(dm.bb_ul[b - 1]);
        const bool has_next = bead_ok && !isinf(dm.bb_ul[bb]);
        double res0 = 0.0, res1 = 0.0, res2 = 0.0;
        if (has_prev) {   // i = b, i-1 = b-1
            const double ul = dm.bb_ul[b - 1], ll = dm.bb_ll[b - 1];
            double d0 = __dsub_rn(xb0, xat(b - 1, 0)), d1 = __dsub_rn(xb1, xat(b - 1, 1)), d2 = __dsub_rn(xb2, xat(b - 1, 2));
            double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            if (s > __dmul_rn(ul, ul)) {
                double d = __dsqrt_rn(s), a = __dsub_rn(d, ul);
                res0 = __ddiv_rn(__dmul_rn(a, d0), d); res1 = __ddiv_rn(__dmul_rn(a, d1), d); res2 = __ddiv_rn(__dmul_rn(a, d2), d);
                e_bb += a * a;
            } else if (s < __dmul_rn(ll, ll)) {
                double d = __dsqrt_rn(s), a = __dsub_rn(ll, d);
                res0 = -__ddiv_rn(__dmul_rn(a, d0), d); res1 = -__ddiv_rn(__dmul_rn(a, d1), d); res2 = -__ddiv_rn(__dmul_rn(a, d2), d);
                e_bb += a * a;
            }
        }
        if (has_next) {   // i = b+1, i-1 = b
            const double ul = dm.bb_ul[bb], ll = dm.bb_ll[bb];
            double d0 = __dsub_rn(xat(b + 1, 0), xb0), d1 = __dsub_rn(xat(b + 1, 1), xb1), d2 = __dsub_rn(xat(b + 1, 2), xb2);
            double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            if (s > __dmul_rn(ul, ul)) {
                double d = __dsqrt_rn(s), a = __dsub_rn(d, ul);
                res0 = __dsub_rn(res0, __ddiv_rn(__dmul_rn(a, d0), d));
                res1 = __dsub_rn(res1, __ddiv_rn(__dmul_rn(a, d1), d));
                res2 = __dsub_rn(res2, __ddiv_rn(__dmul_rn(a, d2), d));
            } else if (s < __dmul_rn(ll, ll)) {
                double d = __dsqrt_rn(s), a = __dsub_rn(ll, d);
                res0 = __dadd_rn(res0, __ddiv_rn(__dmul_rn(a, d0), d));
                res1 = __dadd_rn(res1, __ddiv_rn(__dmul_rn(a, d1), d));
                res2 = __dadd_rn(res2, __ddiv_rn(__dmul_rn(a, d2), d));
            }
        }
        g0 = __dadd_rn(g0, __dmul_rn(dm.k_bb, res0));
        g1 = __dadd_rn(g1, __dmul_rn(dm.k_bb, res1));
        g2 = __dadd_rn(g2, __dmul_rn(dm.k_bb, res2));
    }

    // sphere (sphere_prior_c.pyx:25-36)
    if ((terms & 8) && dm.sphere_active) {
        const double R0 = dm.sph_R;
        // radius2 + bead_radii2[i] - 2 * radius * bead_radii[i]
        const double lim = __dsub_rn(__dadd_rn(__dmul_rn(R0, R0), __dmul_rn(rb, rb)), __dmul_rn(__dmul_rn(2.0, R0), rb));
        double d0 = __dsub_rn(xb0, dm.sph_cx), d1 = __dsub_rn(xb1, dm.sph_cy), d2 = __dsub_rn(xb2, dm.sph_cz);
        double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        if (!(s < lim)) {
            double d = __dsqrt_rn(s);
            double a = __dsub_rn(__dadd_rn(d, rb), R0);
            double f = __ddiv_rn(a, d);
            g0 = __dadd_rn(g0, __dmul_rn(dm.sph_k, __dmul_rn(d0, f)));
            g1 = __dadd_rn(g1, __dmul_rn(dm.sph_k, __dmul_rn(d1, f)));
            g2 = __dadd_rn(g2, __dmul_rn(dm.sph_k, __dmul_rn(d2, f)));
            if (a > 0.0) e_sp += a * a;      // sphere_prior.py:70-72 penalises d + r > R only
        }
    }

}

// coordinates in the SoA working copy xk[3][Np]
__device__ __forceinline__ void bonded_terms(const DevModel &dm, const double *__restrict__ xk, int Np, int b, bool bead_ok, int bb,
                                             double xb0, double xb1, double xb2, double rb, int terms,
                                             double &g0, double &g1, double &g2, double &e_bb, double &e_sp)
{
    bonded_terms_t(dm, [=](int o, int c) { return xk[c * Np + o]; }, b, bead_ok, bb, xb0, xb1, xb2, rb, terms, g0, g1, g2, e_bb, e_sp);
}
// coordinates as (x, y, z, radius) records (verlet_step_kernel's shared-memory copy)
__device__ __forceinline__ void bonded_terms_rec(const DevModel &dm, const double4 *rec, int b,
                                                 double xb0, double xb1, double xb2, double rb, int terms,
                                                 double &g0, double &g1, double &g2, double &e_bb, double &e_sp)
{
    bonded_terms_t(dm, [=](int o, int c) { const double4 v = rec[o]; return c == 0 ? v.x : (c == 1 ? v.y : v.z); },
                   b, true, b, xb0, xb1, xb2, rb, terms, g0, g1, g2, e_bb, e_sp);
}

template <bool EXACT, int EHIC_PRIOR_THREADS>
__global__ void __launch_bounds__(EHIC_PRIOR_THREADS)
prior_kernel(DevModel dm, const double *__restrict__ xs, const double *__restrict__ beta,
             double *__restrict__ gprior, double *__restrict__ partP, int terms, double r_max,
             const double *__restrict__ bbox, int cells_enabled, const int *__restrict__ vl_state = nullptr)
{
    const int k = blockIdx.y, r = blockIdx.z;
    // HMC trajectories: structures with a usable Verlet list are evaluated by verlet_force_kernel (nonbonded,
    // backbone and sphere terms in one pass); this kernel keeps those whose list overflowed (state 2)
    if (vl_state && vl_state[(long long)r * dm.S + k] != 2) {          // CTA-uniform
        if (partP && threadIdx.x == 0) {
            double *o = partP + (((long long)r * dm.S + k) * gridDim.x + blockIdx.x) * 3;
            o[0] = 0.0; o[1] = 0.0; o[2] = 0.0;
        }
        return;
    }
    const int b = blockIdx.x * EHIC_PRIOR_THREADS + threadIdx.x;
    const bool bead_ok = b < dm.N;
    const int bb = bead_ok ? b : dm.N - 1;
    const int Np = dm.Np;
    const double *xk = xs + ((long long)r * dm.S + k) * 3LL * Np;
    const double xb0 = xk[bb], xb1 = xk[Np + bb], xb2 = xk[2 * Np + bb];
    const double rb = dm.radii[bb];
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0;

    if ((terms & 2) && cells_enabled) {       // compact structure: cell_nonbonded_kernel takes the nonbonded term
        __shared__ int s_fits;
        if (threadIdx.x < 32) {
            double lo[3], hi[3];
            bool f = structure_fits_cells(bbox + ((long long)r * dm.S + k) * dm.n_slices * 6, dm.n_slices, r_max, threadIdx.x, lo, hi);
            if (threadIdx.x == 0) s_fits = f ? 1 : 0;
        }
        __syncthreads();
        if (s_fits) terms &= ~2;
    }
    if (terms & 2) {
        // partner tile: EHIC_PRIOR_TILE beads as (x, y), (z, radius), independent of the CTA size so
        // that small systems need one or two barriers per structure
        __shared__ double2 sxy[EHIC_PRIOR_TILE];
        __shared__ double2 szr[EHIC_PRIOR_TILE];
        __shared__ double sbox[EHIC_PRIOR_TILE / 32][6];
        const double rb2 = __dmul_rn(rb, rb);
        const double coarse = (rb + r_max) * (rb + r_max);
        // this warp's own slice box, grown by the largest contact range (+ slack for the <= tests)
        const double reach = 2.0 * r_max * (1.0 + 1e-12);
        // bbox == NULL (Verlet trajectories: only structures whose list overflowed land here, compact ones):
        // no slice boxes were computed, nothing is pruned
        const bool prune = bbox != nullptr;
        const double *bbk = prune ? bbox + ((long long)r * dm.S + k) * dm.n_slices * 6 : nullptr;
        const int my_sl = min((int)((blockIdx.x * EHIC_PRIOR_THREADS + threadIdx.x) >> 5), dm.n_slices - 1);
        const double BIG = 1e300;
        const double mlo0 = prune ? bbk[my_sl * 6] - reach : -BIG, mlo1 = prune ? bbk[my_sl * 6 + 1] - reach : -BIG,
                     mlo2 = prune ? bbk[my_sl * 6 + 2] - reach : -BIG;
        const double mhi0 = prune ? bbk[my_sl * 6 + 3] + reach : BIG, mhi1 = prune ? bbk[my_sl * 6 + 4] + reach : BIG,
                     mhi2 = prune ? bbk[my_sl * 6 + 5] + reach : BIG;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int t0 = 0; t0 < dm.N; t0 += EHIC_PRIOR_TILE) {
            __syncthreads();
            for (int q = threadIdx.x; q < EHIC_PRIOR_TILE; q += EHIC_PRIOR_THREADS) {
                int o = t0 + q;
                int oo = o < dm.N ? o : dm.N - 1;
                sxy[q] = make_double2(xk[oo], xk[Np + oo]);
                szr[q] = make_double2(xk[2 * Np + oo], dm.radii[oo]);
            }
            for (int q = threadIdx.x; q < (EHIC_PRIOR_TILE / 32) * 6; q += EHIC_PRIOR_THREADS) {
                const int psl = min(t0 / 32 + q / 6, dm.n_slices - 1);
                sbox[q / 6][q % 6] = prune ? bbk[psl * 6 + q % 6] : 0.0;
            }
            __syncthreads();
            const int cnt = min(EHIC_PRIOR_TILE, dm.N - t0);
            if (!EXACT) {
                // fast mode: the all-pairs sweep only TESTS (6 FP64 + compare per pair, branch-free,
                // 32 partners per bit mask); the few partners within the coarse contact range are
                // then evaluated from the mask, in ascending order
                for (int q0 = 0; q0 < cnt; q0 += 32) {
                    {   // warp-uniform: partner slice entirely out of reach of this warp's slice?
                        const double *pb = sbox[q0 >> 5];
                        if (pb[0] > mhi0 || pb[3] < mlo0 || pb[1] > mhi1 || pb[4] < mlo1 || pb[2] > mhi2 || pb[5] < mlo2) continue;
                    }
                    unsigned hits = 0u;
#pragma unroll
                    for (int qq = 0; qq < 32; qq++) {
                        const double2 pxy = sxy[q0 + qq], pzr = szr[q0 + qq];     // tile is padded: always readable
                        double dx = pxy.x - xb0, dy = pxy.y - xb1, dz = pzr.x - xb2;
                        double s = fma(dz, dz, fma(dy, dy, dx * dx));
                        hits |= (s < coarse) ? (1u << qq) : 0u;
                    }
                    const int self = b - (t0 + q0);
                    if (self >= 0 && self < 32) hits &= ~(1u << self);
                    if (q0 + 32 > cnt) hits &= (cnt - q0 >= 32) ? 0xffffffffu : ((1u << (cnt - q0)) - 1u);
                    while (hits) {
                        const int qq = __ffs(hits) - 1;
                        hits &= hits - 1u;
                        const double2 pxy = sxy[q0 + qq], pzr = szr[q0 + qq];
                        double dx = pxy.x - xb0, dy = pxy.y - xb1, dz = pzr.x - xb2;
                        double s = fma(dz, dz, fma(dy, dy, dx * dx));
                        double d0s = rb + pzr.y;
                        if (s < d0s * d0s) {
                            double rd = rsqrt_nr(s);
                            double a = d0s - s * rd;
                            double aa = a * a;
                            double g = aa * a * rd;
                            a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
                            e_nb += aa * aa;
                        }
                    }
                }
            }
#pragma unroll 4
            for (int q = 0; EXACT && q < cnt; q++) {
                if ((q & 31) == 0) {   // skipping a slice pair that cannot touch changes neither the terms nor their order
                    const double *pb = sbox[q >> 5];
                    if (pb[0] > mhi0 || pb[3] < mlo0 || pb[1] > mhi1 || pb[4] < mlo1 || pb[2] > mhi2 || pb[5] < mlo2) { q += 31; continue; }
                }
                const double2 pxy = sxy[q], pzr = szr[q];
                const int o = t0 + q;
                if (EXACT) {
                    double d0 = __dsub_rn(xb0, pxy.x), d1 = __dsub_rn(xb1, pxy.y), d2 = __dsub_rn(xb2, pzr.x);
                    double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
                    const double ro = pzr.y;
                    // forcefield_c.pyx:50: d < radii2[i] + radii2[j] + 2 * radii[i] * radii[j]
                    double lim = __dadd_rn(__dadd_rn(rb2, __dmul_rn(ro, ro)), __dmul_rn(__dmul_rn(2.0, rb), ro));
                    if (s < lim && o != b) {
                        double d = __dsqrt_rn(s);
                        double a = __dsub_rn(__dadd_rn(rb, ro), d);
                        double g = __ddiv_rn(__dmul_rn(__dmul_rn(a, a), a), d);
                        a0 = __dadd_rn(a0, __dmul_rn(__dsub_rn(pxy.x, xb0), g));
                        a1 = __dadd_rn(a1, __dmul_rn(__dsub_rn(pxy.y, xb1), g));
                        a2 = __dadd_rn(a2, __dmul_rn(__dsub_rn(pzr.x, xb2), g));
                        double aa = a * a;
                        e_nb += aa * aa;
                    }
                }
            }
        }
        // forcefield_c.pyx:59: force_constant * 2.0 * res ; nonbonded_prior.py:119-121: * beta
        const double k2 = __dmul_rn(dm.k_nb, 2.0);
        const double bet = beta ? beta[r] : 1.0;
        g0 = __dmul_rn(bet, __dmul_rn(k2, a0));
        g1 = __dmul_rn(bet, __dmul_rn(k2, a1));
        g2 = __dmul_rn(bet, __dmul_rn(k2, a2));
    }

    bonded_terms(dm, xk, Np, b, bead_ok, bb, xb0, xb1, xb2, rb, terms, g0, g1, g2, e_bb, e_sp);
    if (bead_ok && gprior) {
        const long long at = (((long long)r * dm.S + k) * dm.N + b) * 3;
        gprior[at] = g0; gprior[at + 1] = g1; gprior[at + 2] = g2;
    }
    if (partP) {
        __shared__ double sh[3][EHIC_PRIOR_THREADS / 32];
        if (!bead_ok) { e_nb = 0.0; e_bb = 0.0; e_sp = 0.0; }
        e_nb = warp_sum(e_nb); e_bb = warp_sum(e_bb); e_sp = warp_sum(e_sp);
        if ((threadIdx.x & 31) == 0) {
            sh[0][threadIdx.x >> 5] = e_nb; sh[1][threadIdx.x >> 5] = e_bb; sh[2][threadIdx.x >> 5] = e_sp;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int q = 0; q < EHIC_PRIOR_THREADS / 32; q++) { t0 += sh[0][q]; t1 += sh[1][q]; t2 += sh[2][q]; }
            double *o = partP + (((long long)r * dm.S + k) * gridDim.x + blockIdx.x) * 3;
            o[0] = t0; o[1] = t1; o[2] = t2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Cell-grid nonbonded kernel (fast mode): the GPU counterpart of the reference's production path,
// a cubic cell grid of edge >= the largest contact range (c/nblist.c:112-303), rebuilt for every
// evaluation like the reference does, but per structure inside ONE CTA:
//   1. bounding box of the structure, grid dimensions (cell edge grows until the grid fits
//      EHIC_CELL_MAX cells, as nblist.c:133-138 inflates its cells),
//   2. counting sort of the beads by cell in shared memory; every cell's bead list is then sorted
//      by bead index, so the layout -- and the summation order -- is deterministic,
//   3. thread <-> bead: the 27 surrounding cells are scanned, contacts evaluated with the same
//      arithmetic as prior_kernel (gather form: each pair seen from both ends, no atomics on FP data).
// gprior += beta * dE_nb/dx (prior_kernel wrote the backbone / sphere part first);
// partC[(r*S + k)*3] = sum a^4 over both visits (slots +1, +2: backbone / sphere, zero here).
#define EHIC_CELL_THREADS 256

__global__ void __launch_bounds__(EHIC_CELL_THREADS)
cell_nonbonded_kernel(DevModel dm, const double *__restrict__ xs, const double *__restrict__ beta,
                      double *__restrict__ gprior, double *__restrict__ partC, double r_max,
                      int *__restrict__ scratch /* [R*S][N] sorted bead lists */, const double *__restrict__ bbox)
{
    __shared__ int s_start[EHIC_CELL_MAX + 1];
    __shared__ int s_fill[EHIC_CELL_MAX];
    __shared__ double s_red[6][EHIC_CELL_THREADS / 32];
    __shared__ double s_box[7];
    __shared__ int s_part[EHIC_CELL_THREADS];
    __shared__ int s_G[3];
    const int k = blockIdx.x, r = blockIdx.y;
    const long long rk = (long long)r * dm.S + k;
    const int N = dm.N, Np = dm.Np, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *xk = xs + rk * 3LL * Np;
    int *list = scratch + rk * N;

    // ---- 1. bounding box from the slice boxes; extended structures are left to prior_kernel --------
    if (tid < 32) {
        double lo[3], hi[3];
        const bool fits = structure_fits_cells(bbox + rk * dm.n_slices * 6, dm.n_slices, r_max, tid, lo, hi);
        if (tid == 0) {
            const double edge = 2.0 * r_max * (1.0 + 1e-9), inv_e = 1.0 / edge;
            for (int c = 0; c < 3; c++) {
                s_box[c] = lo[c]; s_box[3 + c] = hi[c];
                s_G[c] = fits ? (int)(floor((hi[c] - lo[c]) * inv_e) + 1.0) : 0;
            }
            s_box[6] = edge;
        }
    }
    __syncthreads();
    if (s_G[0] == 0) {
        if (partC && tid == 0) { partC[rk * 3] = 0.0; partC[rk * 3 + 1] = 0.0; partC[rk * 3 + 2] = 0.0; }
        return;
    }
    const double edge = s_box[6], inv = 1.0 / edge;
    const int G[3] = {s_G[0], s_G[1], s_G[2]};
    const int ncell = G[0] * G[1] * G[2];
    auto cell_of = [&](double x, double y, double z, int &ci, int &cj, int &ck) {
        ci = min(max((int)floor((x - s_box[0]) * inv), 0), G[0] - 1);
        cj = min(max((int)floor((y - s_box[1]) * inv), 0), G[1] - 1);
        ck = min(max((int)floor((z - s_box[2]) * inv), 0), G[2] - 1);
    };

    // ---- 2. counting sort by cell ---------------------------------------------------------------
    for (int q = tid; q < ncell; q += EHIC_CELL_THREADS) s_fill[q] = 0;
    __syncthreads();
    for (int b = tid; b < N; b += EHIC_CELL_THREADS) {
        int ci, cj, ck; cell_of(xk[b], xk[Np + b], xk[2 * Np + b], ci, cj, ck);
        atomicAdd(&s_fill[(ci * G[1] + cj) * G[2] + ck], 1);       // integer counts: order-independent
    }
    __syncthreads();
    {   // exclusive scan of the counts (each thread a contiguous span, then the span totals)
        const int span = (ncell + EHIC_CELL_THREADS - 1) / EHIC_CELL_THREADS;
        const int q0 = min(tid * span, ncell), q1 = min(q0 + span, ncell);
        int sum = 0;
        for (int q = q0; q < q1; q++) sum += s_fill[q];
        s_part[tid] = sum;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int q = 0; q < EHIC_CELL_THREADS; q++) { int v = s_part[q]; s_part[q] = run; run += v; } s_start[ncell] = N; }
        __syncthreads();
        int run = s_part[tid];
        for (int q = q0; q < q1; q++) { s_start[q] = run; run += s_fill[q]; }
    }
    __syncthreads();
    for (int q = tid; q < ncell; q += EHIC_CELL_THREADS) s_fill[q] = 0;
    __syncthreads();
    for (int b = tid; b < N; b += EHIC_CELL_THREADS) {
        int ci, cj, ck; cell_of(xk[b], xk[Np + b], xk[2 * Np + b], ci, cj, ck);
        const int cell = (ci * G[1] + cj) * G[2] + ck;
        list[s_start[cell] + atomicAdd(&s_fill[cell], 1)] = b;
    }
    __syncthreads();
    for (int q = tid; q < ncell; q += EHIC_CELL_THREADS) {      // deterministic order inside every cell
        const int a = s_start[q], e = s_start[q + 1];
        for (int u = a + 1; u < e; u++) {
            int v = list[u], w = u - 1;
            while (w >= a && list[w] > v) { list[w + 1] = list[w]; w--; }
            list[w + 1] = v;
        }
    }
    __syncthreads();

    // ---- 3. contacts from the 27 surrounding cells ------------------------------------------------
    const double bet = beta ? beta[r] : 1.0;
    const double sc = bet * (dm.k_nb * 2.0);
    double e_nb = 0.0;
    for (int b = tid; b < N; b += EHIC_CELL_THREADS) {
        const double xb0 = xk[b], xb1 = xk[Np + b], xb2 = xk[2 * Np + b];
        const double rb = dm.radii[b];
        int ci, cj, ck; cell_of(xb0, xb1, xb2, ci, cj, ck);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int di = max(ci - 1, 0); di <= min(ci + 1, G[0] - 1); di++)
            for (int dj = max(cj - 1, 0); dj <= min(cj + 1, G[1] - 1); dj++) {
                const int base = (di * G[1] + dj) * G[2];
                const int u0 = s_start[base + max(ck - 1, 0)], u1 = s_start[base + min(ck + 1, G[2] - 1) + 1];
                for (int u = u0; u < u1; u++) {          // the (up to) 3 cells along z are contiguous in the list
                    const int o = list[u];
                    if (o == b) continue;
                    double dx = xk[o] - xb0, dy = xk[Np + o] - xb1, dz = xk[2 * Np + o] - xb2;
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double d0 = rb + dm.radii[o];
                    if (s2 < d0 * d0) {
                        double rd = rsqrt_nr(s2);
                        double a = d0 - s2 * rd, aa = a * a;
                        double g = aa * a * rd;
                        a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
                        e_nb += aa * aa;
                    }
                }
            }
        if (gprior) {
            double *o = gprior + (rk * N + b) * 3;
            o[0] = fma(sc, a0, o[0]); o[1] = fma(sc, a1, o[1]); o[2] = fma(sc, a2, o[2]);
        }
    }
    if (partC) {
        e_nb = warp_sum(e_nb);
        __syncthreads();
        if (lane == 0) s_red[0][warp] = e_nb;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int q = 0; q < EHIC_CELL_THREADS / 32; q++) t += s_red[0][q];
            partC[rk * 3] = t; partC[rk * 3 + 1] = 0.0; partC[rk * 3 + 2] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Verlet lists for the nonbonded term inside HMC trajectories (fast mode, N < EHIC_CELL_MAX).
//
// The reference rebuilds its neighbour list for every energy and every gradient (forcefields.py:214-224,
// c/nblist.c).  Along a leapfrog trajectory beads move by a small fraction of their radius per step, so the
// list is built with a skin -- partners within r_i + r_j + skin -- and reused until some bead of the
// structure has moved more than skin / 2 from where it was at build time; until then no pair outside the
// list can be in contact, so the force and energy are exactly those of the all-pairs sweep (same pairs,
// same ascending partner order, same arithmetic as prior_kernel's fast path).  Everything is decided on
// the device, per structure, so the trajectory stays one CUDA graph:
//   verlet_check_kernel   state[rk]: 0 never built, 1 valid, 2 overflow (sticky: prior_kernel keeps the
//                         structure); stale[rk] = list missing or some bead beyond skin / 2
//   verlet_build_kernel   stale structures only: thread <-> bead, sweep over all partners (slice boxes
//                         prune as in prior_kernel), list[bead][slot] ascending, x_ref = x
//   verlet_force_kernel   CTA <-> structure: gprior = beta * 2k * sum (r_i + r_j - d)^3 / d * (x_o - x_b)
//                         + backbone and sphere terms; partC[rk*3 + {0,1,2}] = energy sums of the three terms
#define EHIC_VL_CAP 64
struct VerletDev {
    int *state, *stale;     // [R*S]
    int *cnt;               // [R*S][Np]
    unsigned short *idx;    // [R*S][Np][EHIC_VL_CAP]: a bead's list is one 128-byte line (N < 4096 fits 16 bits)
    double *xref;           // [R*S][3][Np]
    unsigned short *perm;   // [R*S][Np] force-phase work order of verlet_step_kernel: beads by list length, longest first
    double skin;
};

// A list overflow (> EHIC_VL_CAP partners: a transient collapse, NaN coordinates) must not exile the structure to
// the all-pairs sweep for the rest of the model's life: at the start of every trajectory overflowed structures
// get one new attempt (state 2 -> 0: rebuilt at the first step, back to 2 only if the list overflows again).
__global__ void verlet_retry_kernel(VerletDev vl, long long n)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n && vl.state[q] == 2) vl.state[q] = 0;
}

__global__ void __launch_bounds__(256)
verlet_check_kernel(DevModel dm, VerletDev vl, const double *__restrict__ xs)
{
    const long long rk = blockIdx.x;
    __shared__ double sh[8];
    const int st = vl.state[rk];
    if (st == 2) { if (threadIdx.x == 0) vl.stale[rk] = 0; return; }
    double mx = 0.0;
    if (st == 1) {
        const double *xk = xs + rk * 3LL * dm.Np, *xr = vl.xref + rk * 3LL * dm.Np;
        for (int b = threadIdx.x; b < dm.N; b += 256) {
            const double d0 = xk[b] - xr[b], d1 = xk[dm.Np + b] - xr[dm.Np + b], d2 = xk[2 * dm.Np + b] - xr[2 * dm.Np + b];
            const double d = fma(d2, d2, fma(d1, d1, d0 * d0));
            mx = fmax(mx, (d == d) ? d : 1e300);       // NaN coordinates: rebuild (and let the sweep see them)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 1; q < 8; q++) mx = fmax(mx, sh[q]);
        vl.stale[rk] = (st == 0 || mx > 0.25 * vl.skin * vl.skin) ? 1 : 0;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
verlet_build_kernel(DevModel dm, VerletDev vl, const double *__restrict__ xs, double r_max, const double *__restrict__ bbox)
{
    const int k = blockIdx.y, r = blockIdx.z;
    const long long rk = (long long)r * dm.S + k;
    if (!vl.stale[rk]) return;                                     // CTA-uniform
    const int b = blockIdx.x * THREADS + threadIdx.x;
    const bool bead_ok = b < dm.N;
    const int bb = bead_ok ? b : dm.N - 1;
    const int Np = dm.Np;
    const double *xk = xs + rk * 3LL * Np;
    const double xb0 = xk[bb], xb1 = xk[Np + bb], xb2 = xk[2 * Np + bb];
    const double rb = dm.radii[bb];
    __shared__ double2 sxy[EHIC_PRIOR_TILE];
    __shared__ double2 szr[EHIC_PRIOR_TILE];
    __shared__ double sbox[EHIC_PRIOR_TILE / 32][6];
    const double reach = (2.0 * r_max + vl.skin) * (1.0 + 1e-12);
    const bool prune = bbox != nullptr;                            // rebuilds are rare: without boxes the sweep is plain all-pairs
    const double *bbk = prune ? bbox + rk * dm.n_slices * 6 : nullptr;
    const int my_sl = min((int)((blockIdx.x * THREADS + threadIdx.x) >> 5), dm.n_slices - 1);
    const double BIG = 1e300;
    const double mlo0 = prune ? bbk[my_sl * 6] - reach : -BIG, mlo1 = prune ? bbk[my_sl * 6 + 1] - reach : -BIG,
                 mlo2 = prune ? bbk[my_sl * 6 + 2] - reach : -BIG;
    const double mhi0 = prune ? bbk[my_sl * 6 + 3] + reach : BIG, mhi1 = prune ? bbk[my_sl * 6 + 4] + reach : BIG,
                 mhi2 = prune ? bbk[my_sl * 6 + 5] + reach : BIG;
    unsigned short *my_idx = vl.idx + (rk * Np + bb) * (long long)EHIC_VL_CAP;
    int n = 0;
    bool over = false;
    for (int t0 = 0; t0 < dm.N; t0 += EHIC_PRIOR_TILE) {
        __syncthreads();
        for (int q = threadIdx.x; q < EHIC_PRIOR_TILE; q += THREADS) {
            int o = t0 + q;
            int oo = o < dm.N ? o : dm.N - 1;
            sxy[q] = make_double2(xk[oo], xk[Np + oo]);
            szr[q] = make_double2(xk[2 * Np + oo], dm.radii[oo]);
        }
        for (int q = threadIdx.x; q < (EHIC_PRIOR_TILE / 32) * 6; q += THREADS) {
            const int psl = min(t0 / 32 + q / 6, dm.n_slices - 1);
            sbox[q / 6][q % 6] = prune ? bbk[psl * 6 + q % 6] : 0.0;
        }
        __syncthreads();
        const int cnt = min(EHIC_PRIOR_TILE, dm.N - t0);
        for (int q0 = 0; q0 < cnt; q0 += 32) {
            const double *pb = sbox[q0 >> 5];
            if (pb[0] > mhi0 || pb[3] < mlo0 || pb[1] > mhi1 || pb[4] < mlo1 || pb[2] > mhi2 || pb[5] < mlo2) continue;
            const int qe = min(32, cnt - q0);
            for (int qq = 0; qq < qe; qq++) {
                const double2 pxy = sxy[q0 + qq], pzr = szr[q0 + qq];
                const double dx = pxy.x - xb0, dy = pxy.y - xb1, dz = pzr.x - xb2;
                const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const double lim = rb + pzr.y + vl.skin;
                const int o = t0 + q0 + qq;
                if (!(s2 >= lim * lim) && o != b && bead_ok) {      // !(>=): NaN distances are listed, like the sweep would see them
                    if (n < EHIC_VL_CAP) my_idx[n] = (unsigned short)o;
                    else over = true;
                    n++;
                }
            }
        }
    }
    if (bead_ok) {
        vl.cnt[rk * Np + b] = min(n, EHIC_VL_CAP);
        double *xr = vl.xref + rk * 3LL * Np;
        xr[b] = xb0; xr[Np + b] = xb1; xr[2 * Np + b] = xb2;
    }
    if (over) atomicExch(&vl.state[rk], 2);                        // integer flag: same outcome in any order
}

__global__ void __launch_bounds__(512, 2)
verlet_force_kernel(DevModel dm, VerletDev vl, const double *__restrict__ xs, const double *__restrict__ beta,
                    double *__restrict__ gprior, double *__restrict__ partC, int terms)
{
    const long long rk = blockIdx.x;
    const int r = (int)(rk / dm.S);
    __shared__ double sh[3][32];
    const int nth = blockDim.x;                                    // one pass over the beads when N <= 1024
    __shared__ int s_bin[EHIC_VL_CAP + 2];
    __shared__ short s_perm[EHIC_CELL_MAX];                        // N < EHIC_CELL_MAX on this path
    const int st = vl.state[rk];
    if (st == 2) {                                                 // prior_kernel evaluated this structure
        if (partC && threadIdx.x == 0) { partC[rk * 3] = 0.0; partC[rk * 3 + 1] = 0.0; partC[rk * 3 + 2] = 0.0; }
        return;
    }
    const int Np = dm.Np, N = dm.N;
    const double *xk = xs + rk * 3LL * Np;
    const int *cnt = vl.cnt + rk * Np;
    const unsigned short *idx = vl.idx + rk * (long long)Np * EHIC_VL_CAP;
    const double sc = (beta ? beta[r] : 1.0) * (dm.k_nb * 2.0);
    // The kernel is latency bound (index -> coordinates -> arithmetic, a dozen times per bead), so the beads are
    // handed out sorted by list length (counting sort, longest first): the lanes of a warp then finish together.
    // Which thread takes which bead affects no gradient element; when the energy is wanted (first and last step
    // of a trajectory) the beads keep their natural order so that the energy sum has a fixed order too.
    const bool sorted = partC == nullptr;
    if (sorted) {
        for (int q = threadIdx.x; q < EHIC_VL_CAP + 2; q += nth) s_bin[q] = 0;
        __syncthreads();
        for (int b = threadIdx.x; b < N; b += nth) atomicAdd(&s_bin[EHIC_VL_CAP - cnt[b] + 1], 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int q = 1; q < EHIC_VL_CAP + 2; q++) { const int c = s_bin[q]; s_bin[q] = run; run += c; }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < N; b += nth) s_perm[atomicAdd(&s_bin[EHIC_VL_CAP - cnt[b] + 1], 1)] = (short)b;
        __syncthreads();
    }
    double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0;
    for (int t = threadIdx.x; t < N; t += nth) {
        const int b = sorted ? (int)s_perm[t] : t;
        const double xb0 = xk[b], xb1 = xk[Np + b], xb2 = xk[2 * Np + b];
        const double rb = dm.radii[b];
        const int n = cnt[b];
        const uint4 *row = reinterpret_cast<const uint4 *>(idx + (long long)b * EHIC_VL_CAP);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        // ascending partner order, as in prior_kernel.  The whole list of a bead is one cache line: 8 partner
        // indices per 16-byte load, then the coordinates of 2 partners in flight at a time.
        for (int u0 = 0; u0 < n; u0 += 8) {
            const uint4 w = row[u0 >> 3];
            const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int ub = u0 + 2 * h;
                if (ub < n) {
                    const int o[2] = {(int)(ws[h] & 0xffffu), (int)(ws[h] >> 16)};
                    double px[2], py[2], pz[2], pr[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int oo = (ub + q < n) ? o[q] : b;          // slots past the end of the list are undefined
                        px[q] = xk[oo]; py[q] = xk[Np + oo]; pz[q] = xk[2 * Np + oo]; pr[q] = dm.radii[oo];
                    }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const double dx = px[q] - xb0, dy = py[q] - xb1, dz = pz[q] - xb2;
                        const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        const double d0 = rb + pr[q];
                        if (ub + q < n && s2 < d0 * d0) {
                            const double rd = rsqrt_nr(s2);
                            const double a = d0 - s2 * rd, aa = a * a;
                            const double g = aa * a * rd;
                            a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
                            e_nb += aa * aa;
                        }
                    }
                }
            }
        }
        // backbone and sphere terms of the same bead: the prior gradient is written once, nothing is read back
        double g0 = sc * a0, g1 = sc * a1, g2 = sc * a2;
        bonded_terms(dm, xk, Np, b, true, b, xb0, xb1, xb2, rb, terms, g0, g1, g2, e_bb, e_sp);
        if (gprior) {
            double *o = gprior + (rk * N + b) * 3;
            o[0] = g0; o[1] = g1; o[2] = g2;
        }
    }
    if (threadIdx.x == 0 && st == 0) vl.state[rk] = 1;             // built in this step without overflow
    if (partC) {
        e_nb = warp_sum(e_nb); e_bb = warp_sum(e_bb); e_sp = warp_sum(e_sp);
        if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = e_nb; sh[1][threadIdx.x >> 5] = e_bb; sh[2][threadIdx.x >> 5] = e_sp; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int q = 0; q < (nth >> 5); q++) { t0 += sh[0][q]; t1 += sh[1][q]; t2 += sh[2][q]; }
            partC[rk * 3] = t0; partC[rk * 3 + 1] = t1; partC[rk * 3 + 2] = t2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Gather kernel (second half of north-star kernel 1 + the leapfrog epilogue of kernel 3).
//
// CTA <-> (replica r, slice of 32 consecutive beads, group of 4 structure tiles);
// warp <-> one tile of KT structures; lane <-> bead.
// The slice's rows of the sliced-ELL row view (partner index, contact distance, Poisson
// weight: 20 B per entry per lane) are streamed ONCE per CTA through a 3-stage cp.async
// shared-memory ring and shared by the 4 warps; partner coordinates are gathered from the
// SoA working copy (L1/L2 resident: 48 KB per structure at N=2000).  Each pair is recomputed
// from coordinates instead of being read back from [S][M] scratch as the reference does
// (likelihoods_c.pyx:36-43, 3.2 GB at config D).
//   g_b += f * (x_o - x_b),  f = w_m / (g q d)                          (likelihoods_c.pyx:62-73)
// Epilogue: grad = lammda * g + gprior; optional leapfrog kick + drift (HMC trajectories keep
// x and p on the device) and kinetic-energy partials.
// No atomics anywhere: every output element has exactly one writer and a fixed summation order.
#define EHIC_GATHER_MAX_WARPS 16
#define EHIC_CHUNK 16      // row entries per pipeline stage
#define EHIC_STAGES 3

struct GatherArgs {
    const double *xt;       // [R][nkt][Np][KT][3] positions the gradient is evaluated at (tile-AoS)
    const double *wbuf;     // [R][tot+pad] per-slot Poisson weights (from forward_kernel)
    const double *lammda;   // [R] or NULL
    const double *gprior;   // [R][S][N][3] or NULL
    double *grad;           // [R][S][N][3] or NULL
    double *partK;          // [items] kinetic partials or NULL
    int contact;            // evaluate the contact term
    int nb_fused;           // tile path: nonbonded force evaluated inside backward_tile_kernel
    const double *beta;     // [R] or NULL (only read when nb_fused)
    // leapfrog epilogue (all NULL/0 outside HMC)
    double *p;              // [R][S][N][3] momenta, updated in place:  p -= kick * eps * g
    const double *eps;      // [R]
    double kick;            // 0.5 or 1.0
    double *xs_next;        // [R][S][3][Np]: x + eps * p (after the kick), or NULL (no drift)
    double *xt_next;        // the same in tile-AoS layout
    const double *xs_cur;   // [R][S][3][Np] current positions (SoA), read by combine_kernel for the drift
    int kinetic_mode;       // partK = sum p^2 before the kick (1) / after the kick (2) / none (0)
};

template <bool EXACT, int KT>
__global__ void __launch_bounds__(EHIC_GATHER_MAX_WARPS * 32)
gather_kernel(DevModel dm, GatherArgs ga, int nkt, int nktg)
{
    __shared__ __align__(16) int s_o[EHIC_STAGES][EHIC_CHUNK * 32];
    __shared__ __align__(16) double s_dc[EHIC_STAGES][EHIC_CHUNK * 32];
    __shared__ __align__(16) double s_w[EHIC_STAGES][EHIC_CHUNK * 32];
    __shared__ __align__(8) unsigned long long full[EHIC_STAGES];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ktg = blockIdx.x % nktg;
    const int tmp = blockIdx.x / nktg;
    const int sl = tmp % dm.n_slices;
    const int r = tmp / dm.n_slices;
    const int n_warps = blockDim.x >> 5;
    const int kt = ktg * n_warps + warp;
    const bool kt_ok = kt < nkt;
    const int b = dm.row_bead[sl * 32 + lane];
    const bool bead_ok = b >= 0;
    const int bb = bead_ok ? b : dm.N - 1;
    const int Np = dm.Np;
    const long long plane = 3LL * Np;
    const int k0 = (kt_ok ? kt : 0) * KT;

    // records of this structure tile: bead o -> KT*3 contiguous doubles
    const double *xrec = ga.xt + ((long long)r * nkt + (kt_ok ? kt : 0)) * Np * (KT * 3);
    bool k_ok[KT];
    double xb[KT][3], acc[KT][3];
    {
        double v[KT * 3];
        load_record<KT>(xrec + (long long)bb * (KT * 3), v);
#pragma unroll
        for (int t = 0; t < KT; t++) {
            k_ok[t] = kt_ok && (k0 + t) < dm.S;
            xb[t][0] = v[3 * t]; xb[t][1] = v[3 * t + 1]; xb[t][2] = v[3 * t + 2];
            acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
        }
    }

    if (ga.contact && dm.M > 0) {
        const long long off = dm.slice_off[sl];
        const int width = dm.slice_w[sl];
        const int nchunks = (width + EHIC_CHUNK - 1) / EHIC_CHUNK;
        const int *gj = dm.ent_j + off;
        const double *gdc = dm.ent_dc + off;
        const double *gw = ga.wbuf + (long long)r * dm.wstride + off;
        const double alpha = dm.alpha;
        const double aa = __dmul_rn(alpha, alpha);
        // one chunk = 16 entries x 32 lanes: 2 KB of partner indices + 4 KB + 4 KB of doubles, moved by
        // three TMA bulk copies issued by thread 0 and completing on the stage's mbarrier
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < EHIC_STAGES; q++) mbar_init(&full[q], 1);
            mbar_fence_init();
        }
        __syncthreads();
        auto issue = [&](int c) {
            if (c < nchunks && threadIdx.x == 0) {
                const int st = c % EHIC_STAGES;
                const long long base = (long long)c * EHIC_CHUNK * 32;
                mbar_expect_tx(&full[st], EHIC_CHUNK * 32 * 20u);
                bulk_g2s(&s_o[st][0], gj + base, EHIC_CHUNK * 32 * 4u, &full[st]);
                bulk_g2s(&s_dc[st][0], gdc + base, EHIC_CHUNK * 32 * 8u, &full[st]);
                bulk_g2s(&s_w[st][0], gw + base, EHIC_CHUNK * 32 * 8u, &full[st]);
            }
        };
#pragma unroll
        for (int c = 0; c < EHIC_STAGES - 1; c++) issue(c);
        for (int c = 0; c < nchunks; c++) {
            mbar_wait(&full[c % EHIC_STAGES], (unsigned)((c / EHIC_STAGES) & 1));
            __syncthreads();                       // everyone is done with chunk c-1: its stage may be refilled
            issue(c + EHIC_STAGES - 1);
            const int st = c % EHIC_STAGES;
            const int cnt = min(EHIC_CHUNK, width - c * EHIC_CHUNK);
#pragma unroll 2
            for (int e = 0; e < cnt; e++) {
                const int o = s_o[st][e * 32 + lane];
                const double dc = s_dc[st][e * 32 + lane];
                const double w = s_w[st][e * 32 + lane];
                double v[KT * 3];
                load_record<KT>(xrec + (long long)o * (KT * 3), v);
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    const double ox = v[3 * t], oy = v[3 * t + 1], oz = v[3 * t + 2];
                    if (EXACT) {
                        // likelihoods_c.pyx:64-73 with d, sqrtdenoms recomputed as in pxd:33-37
                        double d0 = __dsub_rn(xb[t][0], ox), d1 = __dsub_rn(xb[t][1], oy), d2 = __dsub_rn(xb[t][2], oz);
                        double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
                        double d = __dsqrt_rn(s);
                        double u = __dsub_rn(dc, d);
                        double g = __dadd_rn(1.0, __dmul_rn(__dmul_rn(aa, u), u));
                        double q = __dsqrt_rn(g);
                        double f = __ddiv_rn(w, __dmul_rn(__dmul_rn(g, q), d));
                        acc[t][0] = __dadd_rn(acc[t][0], __dmul_rn(__dsub_rn(ox, xb[t][0]), f));
                        acc[t][1] = __dadd_rn(acc[t][1], __dmul_rn(__dsub_rn(oy, xb[t][1]), f));
                        acc[t][2] = __dadd_rn(acc[t][2], __dmul_rn(__dsub_rn(oz, xb[t][2]), f));
                    } else {
                        double dx = ox - xb[t][0], dy = oy - xb[t][1], dz = oz - xb[t][2];
                        double s = fma(dz, dz, fma(dy, dy, dx * dx));
                        double rd = rsqrt_nr(s);
                        double tt = alpha * (dc - s * rd);
                        double rq = rsqrt_nr(fma(tt, tt, 1.0));
                        double f = w * ((rq * rq) * (rq * rd));
                        acc[t][0] = fma(f, dx, acc[t][0]);
                        acc[t][1] = fma(f, dy, acc[t][1]);
                        acc[t][2] = fma(f, dz, acc[t][2]);
                    }
                }
            }
        }
    }

    // ---- epilogue -------------------------------------------------------------------------
    const double lam = ga.lammda ? ga.lammda[r] : 1.0;
    double kin = 0.0;
#pragma unroll
    for (int t = 0; t < KT; t++) {
        if (!(k_ok[t] && bead_ok)) continue;
        const int k = k0 + t;
        const long long at = (((long long)r * dm.S + k) * dm.N + b) * 3;
        double g0 = __dmul_rn(lam, acc[t][0]), g1 = __dmul_rn(lam, acc[t][1]), g2 = __dmul_rn(lam, acc[t][2]);
        if (ga.gprior) {
            g0 = __dadd_rn(g0, ga.gprior[at]); g1 = __dadd_rn(g1, ga.gprior[at + 1]); g2 = __dadd_rn(g2, ga.gprior[at + 2]);
        }
        if (ga.grad) { ga.grad[at] = g0; ga.grad[at + 1] = g1; ga.grad[at + 2] = g2; }
        if (ga.p) {
            const double eps = ga.eps[r];
            double p0 = ga.p[at], p1 = ga.p[at + 1], p2 = ga.p[at + 2];
            if (ga.kinetic_mode == 1) kin += p0 * p0 + p1 * p1 + p2 * p2;
            const double ke = ga.kick * eps;
            p0 = fma(-ke, g0, p0); p1 = fma(-ke, g1, p1); p2 = fma(-ke, g2, p2);
            ga.p[at] = p0; ga.p[at + 1] = p1; ga.p[at + 2] = p2;
            if (ga.kinetic_mode == 2) kin += p0 * p0 + p1 * p1 + p2 * p2;
            if (ga.xs_next) {
                const double n0 = fma(eps, p0, xb[t][0]), n1 = fma(eps, p1, xb[t][1]), n2 = fma(eps, p2, xb[t][2]);
                double *xn = ga.xs_next + ((long long)r * dm.S + k) * plane;
                xn[b] = n0; xn[Np + b] = n1; xn[2 * Np + b] = n2;
                double *un = ga.xt_next + ((((long long)r * nkt + kt) * Np + b) * KT + t) * 3;
                un[0] = n0; un[1] = n1; un[2] = n2;
            }
        }
    }
    if (ga.partK) {
        kin = warp_sum(kin);
        if (lane == 0) ga.partK[(long long)blockIdx.x * n_warps + warp] = kin;
    }
}

// ---------------------------------------------------------------------------------------
// verlet_step_kernel: the whole prior of one leapfrog step in ONE launch, CTA <-> (replica, structure).
//
// Replaces the chain verlet_check -> verlet_build -> prior_kernel (early exit) -> verlet_force of a trajectory step:
// four launches of which three usually do nothing, and a force kernel that gathered partner coordinates from L2 a
// pair at a time after sorting its beads with shared-memory atomics (nora2012, 302 replicas: 0.31 ms of a 1.25 ms step).
// Here the structure's coordinates and radii are staged in shared memory once (coalesced), and the CTA
//   1. measures the largest displacement since the list was built (block maximum);
//   2. if the list is missing or stale, rebuilds it from shared memory -- thread <-> bead, all partners in ascending
//      order, same test as verlet_build_kernel -- and stores the new reference positions.  More than EHIC_VL_CAP
//      partners for some bead: state 2, and until the next trajectory (verlet_retry_kernel) the structure is swept
//      all-pairs from shared memory, which is what prior_kernel's fast path computes (same pairs, ascending order);
//   3. evaluates the nonbonded force of every bead from its list, then the backbone and sphere terms (bonded_terms),
//      writes the prior gradient once and, on request, the three energy sums of the structure.
// Lists are read back only by threads of the CTA that wrote them (after a CTA barrier), so no grid-wide ordering is
// needed.  Determinism: a bead's force does not depend on which thread computes it; the energy terms go through a
// per-bead shared-memory table and are summed in bead order, so H has the same bits in every run although the work
// order (a counting sort by list length with shared-memory integer atomics) may differ between runs.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 128 ? 8 : THREADS <= 256 ? 4 : THREADS <= 320 ? 3 : 2)
verlet_step_kernel(DevModel dm, VerletDev vl, const double *__restrict__ xs, const double *__restrict__ beta,
                   double *__restrict__ gprior, double *__restrict__ partC, int terms)
{
    extern __shared__ __align__(16) double vsm[];              // [Np] records (x, y, z, radius), [3][Np] energies, [Np] ints (list lengths)
    __shared__ double sh[3][THREADS / 32];
    __shared__ int s_bin[EHIC_VL_CAP + 2];
    const long long rk = blockIdx.x;
    const int r = (int)(rk / dm.S);
    const int Np = dm.Np, N = dm.N, tid = threadIdx.x;
    const double *xk = xs + rk * 3LL * Np;
    double4 *sb = reinterpret_cast<double4 *>(vsm);
    double *s_e = vsm + 4 * Np;                                // [3][Np] per-bead energy terms (energy steps only)
    int *s_cnt = reinterpret_cast<int *>(vsm + 7 * Np);
    int st = vl.state[rk];
    int *cnt = vl.cnt + rk * Np;
    unsigned short *idx = vl.idx + rk * (long long)Np * EHIC_VL_CAP;
    unsigned short *perm = vl.perm + rk * Np;
    // The kernel is bound by round trips to L2 (state -> coordinates -> work order -> list length and row: half of its
    // stall samples, profiles/r02_nora_final_kernels_R302.txt), so the list of this thread's first bead is requested up
    // front: the work order with the state, the length and the first 16 entries as soon as the bead is known, before the
    // displacement check.  A rebuild in this step overwrites them; they are then read again.
    int pre_b = min((int)perm[min(tid, N - 1)], N - 1);        // (never-built lists: any valid bead)
    double *xr = vl.xref + rk * 3LL * Np;
    double mx = 0.0;
    for (int b = tid; b < Np; b += THREADS) {
        const double x0 = xk[b], x1 = xk[Np + b], x2 = xk[2 * Np + b];
        sb[b] = make_double4(x0, x1, x2, dm.radii[min(b, N - 1)]);
        if (st == 1 && b < N) {
            const double d0 = x0 - xr[b], d1 = x1 - xr[Np + b], d2 = x2 - xr[2 * Np + b];
            const double d = fma(d2, d2, fma(d1, d1, d0 * d0));
            mx = fmax(mx, (d == d) ? d : 1e300);               // NaN coordinates: rebuild (and let the sweep see them)
        }
    }
    int pre_n = cnt[pre_b];
    uint4 pre_w0 = reinterpret_cast<const uint4 *>(idx + (long long)pre_b * EHIC_VL_CAP)[0];
    uint4 pre_w1 = reinterpret_cast<const uint4 *>(idx + (long long)pre_b * EHIC_VL_CAP)[1];
    const bool moved = mx > 0.25 * vl.skin * vl.skin;
    const bool stale = st == 0 || (st == 1 && __syncthreads_or(moved));      // st is CTA-uniform
    if (st != 1) __syncthreads();                                            // shared coordinates complete
    if (stale) {
        // The membership test runs on the FP32 pipe (the N^2 tests of a rebuild cost several normal steps).  A list only has
        // to be a SUPERSET of the pairs within r_i + r_j + skin: every radius is inflated by a bound on the rounding error
        // of the bead's float coordinates (2.5e-7 |x|_1, four times the worst case), so no pair inside the FP64 range is
        // missed; the few extra members contribute exact zeros in the force phase, whose own test stays FP64.
        bool over = false;
        float4 *sf = reinterpret_cast<float4 *>(s_e);            // the energy table is free until the end of the step
        for (int b = tid; b < N; b += THREADS) {
            const double4 me = sb[b];
            const float fx = (float)me.x, fy = (float)me.y, fz = (float)me.z;
            sf[b] = make_float4(fx, fy, fz, (float)me.w * 1.000001f + 2.5e-7f * (fabsf(fx) + fabsf(fy) + fabsf(fz)));
        }
        __syncthreads();
        const float skin = (float)vl.skin * 1.000001f + 1e-6f;
        for (int b = tid; b < N; b += THREADS) {
            const double4 me = sb[b];
            const float4 mf = sf[b];
            unsigned short *row = idx + (long long)b * EHIC_VL_CAP;
            int n = 0;
            for (int o = 0; o < N; o++) {
                const float4 p = sf[o];
                const float dx = p.x - mf.x, dy = p.y - mf.y, dz = p.z - mf.z;
                const float s2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const float lim = mf.w + p.w + skin;
                if (!(s2 >= lim * lim) && o != b) {             // !(>=): NaN distances are listed, like the sweep would see them
                    if (n < EHIC_VL_CAP) row[n] = (unsigned short)o;
                    n++;
                }
            }
            cnt[b] = min(n, EHIC_VL_CAP);
            s_cnt[b] = min(n, EHIC_VL_CAP);
            over = over || n > EHIC_VL_CAP;
            xr[b] = me.x; xr[Np + b] = me.y; xr[2 * Np + b] = me.z;
        }
        st = __syncthreads_or(over) ? 2 : 1;                  // (also: s_cnt complete)
        if (tid == 0) vl.state[rk] = st;
        // work order of the force phase: beads sorted by list length, longest first (the lanes of a warp then walk lists of
        // similar length: a warp costs its longest list).  Counting sort with integer shared-memory atomics: which of two
        // beads with equal list lengths comes first depends on the scheduling, but no result does -- a bead's force does
        // not depend on which thread computes it, and the energies are summed per bead in bead order (below).
        for (int q = tid; q < EHIC_VL_CAP + 2; q += THREADS) s_bin[q] = 0;
        __syncthreads();
        for (int b = tid; b < N; b += THREADS) atomicAdd(&s_bin[EHIC_VL_CAP - s_cnt[b] + 1], 1);
        __syncthreads();
        if (tid < 32) {                                         // exclusive prefix over the 66 bins, three per lane
            int v[3], sum = 0;
#pragma unroll
            for (int u = 0; u < 3; u++) { const int q = 3 * tid + u; v[u] = q < EHIC_VL_CAP + 2 ? s_bin[q] : 0; sum += v[u]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
            int run = incl - sum;
#pragma unroll
            for (int u = 0; u < 3; u++) { const int q = 3 * tid + u; if (q < EHIC_VL_CAP + 2) s_bin[q] = run; run += v[u]; }
        }
        __syncthreads();
        for (int b = tid; b < N; b += THREADS) perm[atomicAdd(&s_bin[EHIC_VL_CAP - s_cnt[b] + 1], 1)] = (unsigned short)b;
        __syncthreads();                                        // perm is read by other threads below (same CTA)
        pre_b = (int)perm[min(tid, N - 1)];
        pre_n = cnt[pre_b];
        pre_w0 = reinterpret_cast<const uint4 *>(idx + (long long)pre_b * EHIC_VL_CAP)[0];
        pre_w1 = reinterpret_cast<const uint4 *>(idx + (long long)pre_b * EHIC_VL_CAP)[1];
    }
    const double sc = (beta ? beta[r] : 1.0) * (dm.k_nb * 2.0);
    for (int t = tid; t < N; t += THREADS) {
        double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0;
        const bool first = t == tid;
        const int b = st == 2 ? t : (first ? pre_b : (int)perm[t]);
        const double4 me = sb[b];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        auto visit = [&](int o, bool live) {
            // branch-free: contacts are common inside compact structures, so a warp takes the contact arithmetic for
            // nearly every visit anyway; a miss contributes through a = 0
            const double4 p = sb[o];
            const double dx = p.x - me.x, dy = p.y - me.y, dz = p.z - me.z;
            const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const double d0 = me.w + p.w;
            const double rd = rsqrt_nr(s2);
            const bool hit = live && s2 < d0 * d0;
            const double a = hit ? d0 - s2 * rd : 0.0, rdm = hit ? rd : 0.0, aa = a * a;
            const double g = aa * a * rdm;
            a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
            e_nb += aa * aa;
        };
        if (st == 2) {
            for (int o = 0; o < N; o++) visit(o == b ? (b + 1 < N ? b + 1 : b - 1) : o, o != b);
        } else {
            // ascending partner order, as in prior_kernel; the list of a bead is one 128-byte line, 8 indices per load
            const int n = first ? pre_n : cnt[b];
            const uint4 *row = reinterpret_cast<const uint4 *>(idx + (long long)b * EHIC_VL_CAP);
            for (int u0 = 0; u0 < n; u0 += 8) {
                const uint4 w = (first && u0 == 0) ? pre_w0 : ((first && u0 == 8) ? pre_w1 : row[u0 >> 3]);
                const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const bool live = u0 + u < n;
                    if (u < 4 || u0 + 4 < n)                   // second half of a block only when the list reaches into it
                        visit(live ? (int)((ws[u >> 1] >> (16 * (u & 1))) & 0xffffu) : (b > 0 ? b - 1 : 1), live);
                }
            }
        }
        double g0 = sc * a0, g1 = sc * a1, g2 = sc * a2;
        bonded_terms_rec(dm, sb, b, me.x, me.y, me.z, me.w, terms, g0, g1, g2, e_bb, e_sp);
        if (gprior) {
            double *o = gprior + (rk * N + b) * 3;
            o[0] = g0; o[1] = g1; o[2] = g2;
        }
        if (partC) { s_e[b] = e_nb; s_e[Np + b] = e_bb; s_e[2 * Np + b] = e_sp; }
    }
    if (partC) {
        __syncthreads();
        double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0;             // bead order, whatever the work order was
        for (int b = tid; b < N; b += THREADS) { e_nb += s_e[b]; e_bb += s_e[Np + b]; e_sp += s_e[2 * Np + b]; }
        e_nb = warp_sum(e_nb); e_bb = warp_sum(e_bb); e_sp = warp_sum(e_sp);
        if ((tid & 31) == 0) { sh[0][tid >> 5] = e_nb; sh[1][tid >> 5] = e_bb; sh[2][tid >> 5] = e_sp; }
        __syncthreads();
        if (tid == 0) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int q = 0; q < THREADS / 32; q++) { t0 += sh[0][q]; t1 += sh[1][q]; t2 += sh[2][q]; }
            partC[rk * 3] = t0; partC[rk * 3 + 1] = t1; partC[rk * 3 + 2] = t2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// LANE path (large sparse contact lists, fast arithmetic): lane <-> structure.
//
// A warp works on ONE bead pair for the 32 structures of a structure group at a time, so the pair's
// metadata (partner index, contact distance, weight) is warp-uniform (broadcast loads), the
// coordinates of a bead are three coalesced 256-byte rows of xg[R][nkg][Np][3][32] however scattered
// the list is, and the ensemble reduction of the forward model is a warp shuffle reduction.  (The
// row path reads one 24*KT-byte record per lane: 32 different cache lines per load instruction,
// which leaves it L1-wavefront / L2-sector bound on large sparse lists,
// profiles/r01_sparse_rowpath_E.txt.)  On lists without locality the bound is the L2 -> SM fabric
// (24 bytes per pair and structure), not the FP64 pipe.  Tried and rejected: moving every 768-byte
// partner row with its own TMA bulk copy (the per-copy cost of the TMA unit makes it 3x slower than
// plain coalesced loads) and L1 prefetch instructions (slower).
// No atomics; every output has one writer and a fixed order.
struct __align__(16) LanePair { int i, j; double dc; };
struct __align__(16) LaneEnt { double dc; int j; int pad; };
struct __align__(16) LanePair32 { int i, j; float dc; int pad; };    // FP32 mode: the same views with float contact distances
struct __align__(8) LaneEnt32 { float dc; int j; };

struct LaneDev {
    int nkg;                        // structure groups of 32 lanes
    const LanePair *pairs;          // [M]  (i, j)-sorted pair view
    const int *slot_a, *slot_b;     // [M]  CSR positions of (i -> j) and (j -> i)
    const long long *row_off;       // [N+1] CSR: both directions of every pair
    const LaneEnt *ent;             // [2M]
    const int *order;               // [N] beads in work order: identity, or longest rows first for ragged lists (cluster trajectory kernel)
    const int4 *items;              // [4 n_quads] backward work items (bead or -1, first entry, end entry, lead | n_seg << 8):
                                    // one warp each; the rows of medium systems are split into up to 4 segments that sit in the
                                    // same quad (lead = warp of the first segment), so that a long row is not one warp's serial walk
    const LanePair32 *pairs32;      // [M]  FP32 mode only
    const LaneEnt32 *ent32;         // [2M] FP32 mode only
};

__device__ __forceinline__ void reduce8(double (&r)[8], int lane);      // defined with the tile kernels below

// Forward, lane form: CTA <-> 256 consecutive pairs of the sorted list (8 warps x 32 pairs), in
// batches of 8 pairs per warp: the 8 per-lane sums over the structure groups are reduced across the
// warp with ONE halving butterfly (reduce8) per batch, and after 4 batches every lane finishes one
// pair (mock count, Poisson weight into both CSR slots, log-likelihood term).  Bead i of
// consecutive pairs is the same most of the time and stays in registers.
//
// Work item = "virtual block" (replica r, block of 256 consecutive pairs).  Two launch shapes run the same body:
// the plain grid (one CTA per virtual block) and the REPLICA-RESIDENT one (lane_resident below).
struct LaneSched {
    int n_sub;           // 256-thread (forward) / 128-thread (backward) groups per CTA of the resident launch
    long long n_virtual; // virtual blocks of the whole launch
    int per_replica;     // virtual blocks per replica (forward: nA; backward: nkg * n_quads)
    const long long *cum;   // [per_replica + 1] cumulative work of a replica's virtual blocks (backward: CSR entries), or NULL
};

__device__ __forceinline__ void group_sync(int group, int n_threads)
{   // named barrier of one thread group of a resident CTA (barrier 0 is __syncthreads)
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(n_threads) : "memory");
}

// Replica-resident launch: ONE CTA per SM, each walking a CONTIGUOUS range of virtual blocks -- i.e. of one or two
// replicas -- with all its thread groups.  The lane copy of one (replica, structure group) of a medium system
// (nora2012 3 kb: 308 beads x 768 B = 236 KB) then stays in that SM's L1 for the whole pass and every partner row is
// fetched from L2 once per replica instead of once per pair (the plain grid spreads consecutive blocks of a replica over
// all SMs: 6.5 - 8 TB/s of L2 -> SM traffic, which is what bounds it; profiles/r01_nora_lanepath.txt).  Ranges hold equal
// WORK (cumulative row lengths), not equal block counts: the backward pass walks the beads longest row first.
__device__ __forceinline__ void lane_resident_range(const LaneSched &sc, long long &v0, long long &v1)
{
    const int G = gridDim.x, g = blockIdx.x;
    if (!sc.cum) { v0 = sc.n_virtual * g / G; v1 = sc.n_virtual * (g + 1) / G; return; }
    const long long per = sc.cum[sc.per_replica];
    const long long n_rep = sc.n_virtual / sc.per_replica;
    auto locate = [&](int q) -> long long {                        // first virtual block at or after work fraction q / G
        if (q >= G) return sc.n_virtual;
        const long long tgt = (long long)((double)per * (double)n_rep * ((double)q / (double)G));
        const long long r = min(tgt / max(per, 1LL), n_rep - 1), w = tgt - r * per;
        int lo = 0, hi = sc.per_replica;                           // smallest b with cum[b] >= w
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sc.cum[mid] >= w) hi = mid; else lo = mid + 1; }
        return r * sc.per_replica + lo;
    };
    v0 = locate(g); v1 = locate(g + 1);
}

__device__ __forceinline__ void
forward_lane_body(const DevModel &dm, const LaneDev &ld, const double *__restrict__ xg, const double *__restrict__ norm,
                  double *__restrict__ md_out, double *__restrict__ wbuf,
                  double *__restrict__ partA, int want_logl, int vblock, int r, int tid, int group, double *sh /* [8] */,
                  int4 *stage /* [256] */)
{
    const int lane = tid & 31, warp = tid >> 5;
    const long long m0 = ((long long)vblock * 8 + warp) * 32;
    const double nrm = norm[r], alpha = dm.alpha;
    const long long gstride = (long long)dm.Np * 96;
    const double *xr = xg + (long long)r * ld.nkg * gstride + lane;
    double term = 0.0;
    if (m0 < dm.M) {                                    // warp-uniform
        // the warp's 32 pair records: one coalesced load, parked in the warp's shared-memory slot, read back as broadcasts
        int4 *se = stage + warp * 32;
        __syncwarp();
        se[lane] = __ldg(reinterpret_cast<const int4 *>(ld.pairs + min(m0 + lane, dm.M - 1)));
        __syncwarp();
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const long long mb = m0 + b * 8;
            if (mb < dm.M) {                            // warp-uniform
                double acc[8];
#pragma unroll
                for (int e = 0; e < 8; e++) acc[e] = 0.0;
                const int i_first = se[b * 8].x;
                const bool one_row = i_first == se[b * 8 + 7].x;      // sorted by (i, j): the 8 pairs share bead i (the usual case)
                for (int kg = 0; kg < ld.nkg; kg++) {
                    const double *xk = xr + kg * gstride;
                    const bool kv = kg * 32 + lane < dm.S;
                    const double *p0 = xk + (long long)i_first * 96;
                    double a0 = p0[0], a1 = p0[32], a2 = p0[64];
                    if (one_row) {
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const int4 v = se[b * 8 + e];
                            const double *q = xk + (long long)v.y * 96;
                            const double dx = a0 - q[0], dy = a1 - q[32], dz = a2 - q[64];
                            const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                            const double rd = rsqrt_nr(s2);
                            const double tt = alpha * (__hiloint2double(v.w, v.z) - s2 * rd);
                            const double rq = rsqrt_nr(fma(tt, tt, 1.0));
                            const double sv = fma(tt, rq, 1.0);
                            acc[e] += kv ? sv : 0.0;
                        }
                    } else {
                        int ci = i_first;
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const int4 v = se[b * 8 + e];
                            if (v.x != ci) {            // warp-uniform
                                ci = v.x;
                                const double *p = xk + (long long)ci * 96;
                                a0 = p[0]; a1 = p[32]; a2 = p[64];
                            }
                            const double *q = xk + (long long)v.y * 96;
                            const double dx = a0 - q[0], dy = a1 - q[32], dz = a2 - q[64];
                            const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                            const double rd = rsqrt_nr(s2);
                            const double tt = alpha * (__hiloint2double(v.w, v.z) - s2 * rd);
                            const double rq = rsqrt_nr(fma(tt, tt, 1.0));
                            const double sv = fma(tt, rq, 1.0);
                            acc[e] += kv ? sv : 0.0;
                        }
                    }
                }
                reduce8(acc, lane);                     // lane L holds the total of pair mb + (L >> 2)
                tot[b] = acc[0];
            }
        }
        // lane L finishes pair m0 + 8 (L & 3) + (L >> 2)
        const int bsel = lane & 3;
        const double sum = bsel == 0 ? tot[0] : (bsel == 1 ? tot[1] : (bsel == 2 ? tot[2] : tot[3]));
        const long long m = m0 + 8 * bsel + (lane >> 2);
        if (m < dm.M) {
            const double c = dm.pcnt[m];
            const double md = sum * (0.5 * nrm);
            if (md_out) md_out[(long long)r * dm.M + dm.perm[m]] = md;
            const double em = __dsub_rn(1.0, __ddiv_rn(c, md));
            const double w = __dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), alpha);
            if (wbuf) {
                double *wr = wbuf + (long long)r * dm.wstride;
                wr[ld.slot_a[m]] = w;
                wr[ld.slot_b[m]] = w;
            }
            if (want_logl) term = c * log(md) - md;
        }
    }
    if (want_logl) {
        double v = warp_sum(term);
        if (lane == 0) sh[warp] = v;
        group_sync(group, 256);
        if (tid == 0)
            partA[(long long)r * ((dm.M + 255) / 256) + vblock] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
        group_sync(group, 256);                          // sh is reused by the group's next virtual block
    }
}

__global__ void __launch_bounds__(256, 3)       // 80 registers: 24 warps per SM hide the L2 latency of the partner rows better than 16
forward_lane_kernel(DevModel dm, LaneDev ld, const double *__restrict__ xg, const double *__restrict__ norm,
                    double *__restrict__ md_out, double *__restrict__ wbuf,
                    double *__restrict__ partA, int want_logl)
{
    __shared__ double sh[8];
    __shared__ int4 stage[256];
    forward_lane_body(dm, ld, xg, norm, md_out, wbuf, partA, want_logl, blockIdx.x, blockIdx.y, threadIdx.x, 0, sh, stage);
}

#ifndef EHIC_LANE_FWD_GROUPS
#define EHIC_LANE_FWD_GROUPS 3                   // resident launch: 3 x 256 threads per SM
#endif
__global__ void __launch_bounds__(EHIC_LANE_FWD_GROUPS * 256, 1)
forward_lane_resident_kernel(DevModel dm, LaneDev ld, LaneSched sc, const double *__restrict__ xg, const double *__restrict__ norm,
                             double *__restrict__ md_out, double *__restrict__ wbuf,
                             double *__restrict__ partA, int want_logl)
{
    __shared__ double sh[EHIC_LANE_FWD_GROUPS][8];
    __shared__ int4 stage[EHIC_LANE_FWD_GROUPS][256];
    const int group = threadIdx.x >> 8, tid = threadIdx.x & 255;
    long long v0, v1;
    lane_resident_range(sc, v0, v1);
    for (long long v = v0 + group; v < v1; v += EHIC_LANE_FWD_GROUPS) {
        const int r = (int)(v / sc.per_replica), vb = (int)(v - (long long)r * sc.per_replica);
        forward_lane_body(dm, ld, xg, norm, md_out, wbuf, partA, want_logl, vb, r, tid, group, sh[group], stage[group]);
    }
}

#define EHIC_LANE_WARPS 4
#define EHIC_PRAGMA(x) _Pragma(#x)
#define EHIC_UNROLL(n) EHIC_PRAGMA(unroll n)
#ifndef EHIC_LANE_BWD_UNROLL
#define EHIC_LANE_BWD_UNROLL 4
#endif
#ifndef EHIC_LANE_BWD_MIN_CTAS
#define EHIC_LANE_BWD_MIN_CTAS 8      // 64 registers, 32 warps per SM: measured best of 5..9 CTAs x unroll 4 / 6 / 8 on nora2012 and config E
#endif
// Epilogue shared by the FP64 and FP32 backward lane kernels: lammda * g + priors, optional leapfrog kick + drift
// (the drifted positions go to both working copies; F32: the lane copy holds (hi, lo) float pairs), kinetic partials.
template <bool F32>
__device__ __forceinline__ void lane_epilogue(const DevModel &dm, const LaneDev &ld, const GatherArgs &ga, long long vblock, int r, int kg, int k, int b,
                                              int lane, int warp, bool live, long long gstride,
                                              double x0, double x1, double x2, double acc0, double acc1, double acc2)
{
    const double lam = ga.lammda ? ga.lammda[r] : 1.0;
    double kin = 0.0;
    if (live) {
        const long long at = (((long long)r * dm.S + k) * dm.N + b) * 3;
        double g0 = __dmul_rn(lam, acc0), g1 = __dmul_rn(lam, acc1), g2 = __dmul_rn(lam, acc2);
        if (ga.gprior) {
            g0 = __dadd_rn(g0, ga.gprior[at]); g1 = __dadd_rn(g1, ga.gprior[at + 1]); g2 = __dadd_rn(g2, ga.gprior[at + 2]);
        }
        if (ga.grad) { ga.grad[at] = g0; ga.grad[at + 1] = g1; ga.grad[at + 2] = g2; }
        if (ga.p) {
            const double eps = ga.eps[r];
            double p0 = ga.p[at], p1 = ga.p[at + 1], p2 = ga.p[at + 2];
            if (ga.kinetic_mode == 1) kin += p0 * p0 + p1 * p1 + p2 * p2;
            const double ke = ga.kick * eps;
            p0 = fma(-ke, g0, p0); p1 = fma(-ke, g1, p1); p2 = fma(-ke, g2, p2);
            ga.p[at] = p0; ga.p[at + 1] = p1; ga.p[at + 2] = p2;
            if (ga.kinetic_mode == 2) kin += p0 * p0 + p1 * p1 + p2 * p2;
            if (ga.xs_next) {
                if (F32) {                                   // the lane copy holds float pairs: exact positions from the SoA copy
                    const double *xc = ga.xs_cur + ((long long)r * dm.S + k) * 3LL * dm.Np;
                    x0 = xc[b]; x1 = xc[dm.Np + b]; x2 = xc[2 * dm.Np + b];
                }
                const double n0 = fma(eps, p0, x0), n1 = fma(eps, p1, x1), n2 = fma(eps, p2, x2);
                double *xn = ga.xs_next + ((long long)r * dm.S + k) * 3LL * dm.Np;
                xn[b] = n0; xn[dm.Np + b] = n1; xn[2 * dm.Np + b] = n2;
                double *un = ga.xt_next + ((long long)r * ld.nkg + kg) * gstride + (long long)b * 96 + lane;
                un[0] = F32 ? split_hilo(n0) : n0; un[32] = F32 ? split_hilo(n1) : n1; un[64] = F32 ? split_hilo(n2) : n2;
            }
        }
    }
    if (ga.partK) {
        kin = warp_sum(kin);
        if (lane == 0) ga.partK[vblock * EHIC_LANE_WARPS + warp] = kin;
    }
}

// Segments of one bead within a quad: the partial forces go through shared memory and are added by the leader warp in
// segment order (deterministic).  Returns whether this warp finishes the bead.
__device__ __forceinline__ bool lane_combine_segments(double (*part)[3][32] /* [EHIC_LANE_WARPS] */, int group, int warp, int lane,
                                                      int lead, int nseg, double &acc0, double &acc1, double &acc2)
{
    if (nseg > 1) { part[warp][0][lane] = acc0; part[warp][1][lane] = acc1; part[warp][2][lane] = acc2; }
    group_sync(group, EHIC_LANE_WARPS * 32);
    if (nseg > 1 && warp == lead) {
        acc0 = part[lead][0][lane]; acc1 = part[lead][1][lane]; acc2 = part[lead][2][lane];
        for (int q = 1; q < nseg; q++) { acc0 += part[lead + q][0][lane]; acc1 += part[lead + q][1][lane]; acc2 += part[lead + q][2][lane]; }
    }
    return warp == lead;
}

// Backward, lane form: warp <-> (replica, structure group, bead b); lane <-> structure.  The warp
// walks the CSR row of b -- (partner, contact distance) and the weight are broadcast loads, the
// partner's coordinates three coalesced rows -- and accumulates g_b of its 32 structures in
// registers: a gather, every pair evaluated from both ends, no atomics, fixed order.  CTA = 4
// consecutive beads of one structure group, so neighbouring rows share partner lines in L1.
// Epilogue as in gather_kernel (lammda * g + priors, optional leapfrog kick + drift).
__device__ __forceinline__ void
backward_lane_body(const DevModel &dm, const LaneDev &ld, const GatherArgs &ga, int n_quads, long long vblock, int tid, int group,
                   int4 *stage_ent /* [EHIC_LANE_WARPS * 32] */, double *stage_w /* [EHIC_LANE_WARPS * 32] */,
                   double (*part)[3][32] /* [EHIC_LANE_WARPS] */)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int quad = (int)(vblock % n_quads);
    const int tmp = (int)(vblock / n_quads);
    const int kg = tmp % ld.nkg;
    const int r = tmp / ld.nkg;
    const int4 item = __ldg(ld.items + quad * EHIC_LANE_WARPS + warp);
    const bool bead_ok = item.x >= 0;
    const int b = bead_ok ? item.x : dm.N - 1;
    const int bb = b;
    const int k = kg * 32 + lane;
    const bool k_ok = k < dm.S;
    const long long gstride = (long long)dm.Np * 96;
    const double *xk = ga.xt + ((long long)r * ld.nkg + kg) * gstride + lane;
    const double x0 = xk[(long long)bb * 96], x1 = xk[(long long)bb * 96 + 32], x2 = xk[(long long)bb * 96 + 64];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    if (ga.contact && dm.M > 0 && bead_ok) {            // warp-uniform
        const long long e0 = item.y, e1 = item.z;        // this warp's segment of the bead's CSR row
        const double *wr = ga.wbuf + (long long)r * dm.wstride;
        const double alpha = dm.alpha;
        // The row's metadata travels in blocks of 32 entries: ONE coalesced load brings (dc, partner) and the weight of
        // entry eb + lane, the block is parked in the warp's shared-memory slot and read back as broadcasts.  The
        // partner index of every entry of the block is then known up front, so the coordinate loads of the next entries
        // are not chained behind a metadata load each (that chain was 17 % of the kernel's stall samples), and the next
        // block is requested while this one is worked on.
        int4 *se = stage_ent + (tid >> 5) * 32;
        double *sw = stage_w + (tid >> 5) * 32;
        int4 nv = make_int4(0, 0, 0, 0); double nw = 0.0;
        if (e0 < e1) { const long long e = min(e0 + lane, e1 - 1); nv = __ldg(reinterpret_cast<const int4 *>(ld.ent + e)); nw = wr[e]; }
        for (long long eb = e0; eb < e1; eb += 32) {
            __syncwarp();                                // everybody is done reading the previous block
            se[lane] = nv; sw[lane] = nw;
            __syncwarp();
            if (eb + 32 < e1) { const long long e = min(eb + 32 + lane, e1 - 1); nv = __ldg(reinterpret_cast<const int4 *>(ld.ent + e)); nw = wr[e]; }
            const int n = (int)min(32LL, e1 - eb);
EHIC_UNROLL(EHIC_LANE_BWD_UNROLL)
            for (int qe = 0; qe < n; qe++) {
                const int4 v = se[qe];
                const double dc = __hiloint2double(v.y, v.x);
                const double w = sw[qe];
                const double *q = xk + (long long)v.z * 96;
                const double dx = q[0] - x0, dy = q[32] - x1, dz = q[64] - x2;
                const double s = fma(dz, dz, fma(dy, dy, dx * dx));
                const double rd = rsqrt_nr(s);
                const double tt = alpha * (dc - s * rd);
                const double rq = rsqrt_nr(fma(tt, tt, 1.0));
                const double f = w * ((rq * rq) * (rq * rd));
                acc0 = fma(f, dx, acc0);
                acc1 = fma(f, dy, acc1);
                acc2 = fma(f, dz, acc2);
            }
        }
    }
    const bool leader = lane_combine_segments(part, group, warp, lane, item.w & 255, item.w >> 8, acc0, acc1, acc2);
    lane_epilogue<false>(dm, ld, ga, vblock, r, kg, k, b, lane, warp, k_ok && bead_ok && leader, gstride, x0, x1, x2, acc0, acc1, acc2);
}

__global__ void __launch_bounds__(EHIC_LANE_WARPS * 32, EHIC_LANE_BWD_MIN_CTAS)
backward_lane_kernel(DevModel dm, LaneDev ld, GatherArgs ga, int n_quads)
{
    __shared__ int4 s_ent[EHIC_LANE_WARPS * 32];
    __shared__ double s_w[EHIC_LANE_WARPS * 32];
    __shared__ double s_part[EHIC_LANE_WARPS][3][32];
    backward_lane_body(dm, ld, ga, n_quads, blockIdx.x, threadIdx.x, 0, s_ent, s_w, s_part);
}

#define EHIC_LANE_BWD_GROUPS 8                   // resident launch: 8 x 128 threads per SM
__global__ void __launch_bounds__(EHIC_LANE_BWD_GROUPS * EHIC_LANE_WARPS * 32, 1)
backward_lane_resident_kernel(DevModel dm, LaneDev ld, LaneSched sc, GatherArgs ga, int n_quads)
{
    __shared__ int4 s_ent[EHIC_LANE_BWD_GROUPS][EHIC_LANE_WARPS * 32];
    __shared__ double s_w[EHIC_LANE_BWD_GROUPS][EHIC_LANE_WARPS * 32];
    const int group = threadIdx.x / (EHIC_LANE_WARPS * 32), tid = threadIdx.x % (EHIC_LANE_WARPS * 32);
    long long v0, v1;
    lane_resident_range(sc, v0, v1);
    __shared__ double s_part[EHIC_LANE_BWD_GROUPS][EHIC_LANE_WARPS][3][32];
    for (long long v = v0 + group; v < v1; v += EHIC_LANE_BWD_GROUPS) {
        backward_lane_body(dm, ld, ga, n_quads, v, tid, group, s_ent[group], s_w[group], s_part[group]);
        group_sync(group, EHIC_LANE_WARPS * 32);         // the leader has read the partials before the group's next block overwrites them
    }
}

// ---------------------------------------------------------------------------------------
// TILE path (dense contact lists, fast arithmetic): every unordered bead pair is evaluated ONCE
// in the backward pass.  See layout.hpp (3) for the slot geometry.
struct TileDev {
    int n_rb, n_tiles, n_sub, n_sl;
    const int *tile_I, *tile_B, *rb_begin, *sub_tile, *sub_rb, *exists;
    const unsigned *mask;
    const double *dc, *cnt;
    const long long *orig;
    long long slots;          // n_tiles * 4096: per-replica stride of the tile weights
};

// Forward, tile form: CTA <-> (replica, tile of 128 row beads x 32 column beads), 8 warps =
// 4 row slices x 2 column halves; lane <-> row bead.  Per structure tile (KT structures) the
// tile-AoS records of the 128 row beads (12 KB, contiguous) and of the 32 column beads (3 KB,
// contiguous) are staged in shared memory by a 3-stage cp.async ring shared by the whole CTA;
// column beads are then read with warp-uniform (broadcast) LDS.128.  The S structures are
// walked in order with one accumulator per column: no cross-lane reduction, the reference's
// summation order over the ensemble.  Writes md (data_points order), the Poisson weight of
// every present slot, and per-block partials of log L.
#define EHIC_FT_STAGES 2
template <int KT>
__global__ void __launch_bounds__(256, 3)
forward_tile_kernel(DevModel dm, TileDev td, const double *__restrict__ xt, const double *__restrict__ norm,
                    const double *__restrict__ lammda, double *__restrict__ md_out, double *__restrict__ tw,
                    double *__restrict__ partA, int want_logl, int nkt)
{
    constexpr int QC = 16;
    constexpr int REC = KT * 3;                       // doubles per record
    extern __shared__ __align__(16) double smem[];
    double *s_dc = smem;                               // [4096]
    double *s_row = s_dc + 4096;                       // [STAGES][128 * REC]
    double *s_col = s_row + EHIC_FT_STAGES * 128 * REC;   // [STAGES][32 * REC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rb = warp & 3, h = warp >> 2;
    const int tile = blockIdx.x, r = blockIdx.y;
    const int I = td.tile_I[tile], B = td.tile_B[tile];
    const int Np = dm.Np;
    const int row0 = 128 * I, col0 = 32 * B;           // Np is padded to 32; rows may run past it: clamp the copy
    const double *recs = xt + (long long)r * nkt * Np * REC;
    const unsigned mask = td.mask[((long long)tile * 4 + rb) * 32 + lane] >> (h * QC);
    const bool warp_live = __ballot_sync(0xffffffffu, mask != 0u) != 0u;
    const double alpha = dm.alpha;
    const int n_row = min(128, Np - row0);             // row records that exist in xt

    // TMA bulk copies completing on mbarriers: full[st] for the per-structure-tile records,
    // full[STAGES] for the tile's contact distances (32 KB, once)
    unsigned long long *full = reinterpret_cast<unsigned long long *>(s_col + EHIC_FT_STAGES * 32 * REC);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q <= EHIC_FT_STAGES; q++) mbar_init(&full[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&full[EHIC_FT_STAGES], 4096 * (unsigned)sizeof(double));
        bulk_g2s(s_dc, td.dc + (long long)tile * 4096, 16384, &full[EHIC_FT_STAGES]);
        bulk_g2s(s_dc + 2048, td.dc + (long long)tile * 4096 + 2048, 16384, &full[EHIC_FT_STAGES]);
    }
    auto issue = [&](int kt) {
        if (kt < nkt && threadIdx.x == 0) {
            const int st = kt % EHIC_FT_STAGES;
            const double *src_row = recs + ((long long)kt * Np + row0) * REC;
            const double *src_col = recs + ((long long)kt * Np + col0) * REC;
            const unsigned row_bytes = (unsigned)(n_row * REC * sizeof(double)), col_bytes = (unsigned)(32 * REC * sizeof(double));
            mbar_expect_tx(&full[st], row_bytes + col_bytes);
            bulk_g2s(&s_row[st * 128 * REC], src_row, row_bytes, &full[st]);
            bulk_g2s(&s_col[st * 32 * REC], src_col, col_bytes, &full[st]);
        }
    };
#pragma unroll
    for (int c = 0; c < EHIC_FT_STAGES - 1; c++) issue(c);
    mbar_wait(&full[EHIC_FT_STAGES], 0u);

    double acc[QC];
#pragma unroll
    for (int q = 0; q < QC; q++) acc[q] = 0.0;
    const int nfull = dm.S / KT;
    for (int kt = 0; kt < nkt; kt++) {
        mbar_wait(&full[kt % EHIC_FT_STAGES], (unsigned)((kt / EHIC_FT_STAGES) & 1));
        __syncthreads();                                  // structure tile kt-1 is consumed: its stage may be refilled
        issue(kt + EHIC_FT_STAGES - 1);
        if (!warp_live) continue;
        const int st = kt % EHIC_FT_STAGES;
        const int nvalid = kt < nfull ? KT : dm.S - nfull * KT;
        double own[REC];
        {
            const double *p = &s_row[st * 128 * REC + (rb * 32 + lane) * REC];
#pragma unroll
            for (int u = 0; u < REC; u++) own[u] = p[u];
        }
        const double *pc = &s_col[st * 32 * REC + h * QC * REC];
        const double *pd = &s_dc[(h * QC * 4 + rb) * 32 + lane];
        if (nvalid == KT) {
#pragma unroll
            for (int q = 0; q < QC; q++) {
                const double dc = pd[q * 128];
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    double dx = pc[q * REC + 3 * t] - own[3 * t], dy = pc[q * REC + 3 * t + 1] - own[3 * t + 1],
                           dz = pc[q * REC + 3 * t + 2] - own[3 * t + 2];
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    double rd = rsqrt_nr(s2);
                    double tt = alpha * (dc - s2 * rd);
                    double rq = rsqrt_nr(fma(tt, tt, 1.0));
                    acc[q] += fma(tt, rq, 1.0);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < QC; q++) {
                const double dc = pd[q * 128];
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    if (t < nvalid) {
                        double dx = pc[q * REC + 3 * t] - own[3 * t], dy = pc[q * REC + 3 * t + 1] - own[3 * t + 1],
                               dz = pc[q * REC + 3 * t + 2] - own[3 * t + 2];
                        double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        double rd = rsqrt_nr(s2);
                        double tt = alpha * (dc - s2 * rd);
                        double rq = rsqrt_nr(fma(tt, tt, 1.0));
                        acc[q] += fma(tt, rq, 1.0);
                    }
                }
            }
        }
    }
    double term = 0.0;
    const double nrm = norm[r];
    const double lam = lammda ? lammda[r] : 1.0;       // the backward tile kernel consumes lammda * w
    const long long pos0 = (((long long)tile * 32 + h * QC) * 4 + rb) * 32 + lane;
#pragma unroll
    for (int q = 0; q < QC; q++) {
        if ((mask >> q) & 1u) {
            const long long pos = pos0 + q * 128;
            const double md = acc[q] * (0.5 * nrm);
            const double c = td.cnt[pos];
            if (md_out) md_out[(long long)r * dm.M + td.orig[pos]] = md;
            if (tw) {
                double em = __dsub_rn(1.0, __ddiv_rn(c, md));
                tw[(long long)r * td.slots + pos] = lam * __dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), alpha);
            }
            if (want_logl) term += c * log(md) - md;
        }
    }
    if (want_logl) {
        __shared__ double sh[8];
        double v = warp_sum(term);
        if (lane == 0) sh[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0)
            partA[(long long)r * gridDim.x + blockIdx.x] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
    }
}

// Halving butterfly: 8 per-lane values -> the warp-wide sums of the 8 columns; after the call
// r[0] of lane L is the total of column ((L>>2)&1) + 2*((L>>3)&1) + 4*((L>>4)&1).
// 9 shuffle-adds instead of 40; fixed order (deterministic).
__device__ __forceinline__ void reduce8(double (&r)[8], int lane)
{
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double send = hi ? r[i] : r[i + 4], keep = hi ? r[i + 4] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double send = hi ? r[i] : r[i + 2], keep = hi ? r[i + 2] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = (lane & 4) != 0;
    {
        double send = hi ? r[0] : r[1], keep = hi ? r[1] : r[0];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 2);
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
}

// Same for 4 values: after the call r[0] of lane L is the total of column ((L>>3)&1) + 2*((L>>4)&1).
__device__ __forceinline__ void reduce4(double (&r)[4], int lane)
{
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double send = hi ? r[i] : r[i + 2], keep = hi ? r[i + 2] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
    {
        double send = hi ? r[0] : r[1], keep = hi ? r[1] : r[0];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 4);
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 2);
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
}

// Backward, tile form: CTA <-> (replica, row block of 128 beads, group of 4 structures); warp <->
// structure; each lane owns 4 row beads (accumulated in registers over all column slices of the row
// block).  The (contact distance, weight) slots do not depend on the structure: chunks of 8 columns
// (1024 slots, 16 KB) stream once per CTA through a 3-stage cp.async ring shared by the 4 warps.
// Column beads live one per lane and are broadcast with shuffles; their partial reactions are
// summed over the warp with reduce8 and written to the per-row-block buffer P[r][k][I][bead][3].
// No atomics: action and reaction of every pair are evaluated ONCE.
//
// The weights arrive already multiplied by lammda.  When the tiles cover every bead pair (NB),
// the repulsive nonbonded force is fused in: the pair distance is already there, the contact test
// is one compare against a per-row-bead coarse bound, and the rare hits are handled in a
// warp-uniform branch (beta * 2 k (r_b + r_o - d)^3 / d, forcefield_c.pyx:51-57).
// "Clean" tiles (strictly above the diagonal blocks, every slot present, no clamped beads) skip
// all masking.
#define EHIC_BT_STAGES 3

struct BackwardTileArgs {
    const double *xs, *tw, *beta;
    double *Grow, *P;
    const unsigned char *tile_clean;   // [n_tiles]
    const double *bbox;                // [R*S][n_slices][6] slice bounding boxes (NB only)
    int n_kg, nb;                      // structure groups per row block; fuse the nonbonded term
    double r_max;
};

#ifndef EHIC_BT_NC
#define EHIC_BT_NC 8        // columns whose reactions are reduced together (8 or 4)
#endif

template <bool MASKED, bool NB>
__device__ __forceinline__ void backward_chunk(const DevModel &dm, const double *__restrict__ sd, int qbase, int lane,
                                               double cx, double cy, double cz, double cr, const unsigned (&m)[4],
                                               int row0, int col0, const double (&xb)[4][3],
                                               double coarse, double nbs,
                                               double (&acc)[4][3], double (&rx)[EHIC_BT_NC], double (&ry)[EHIC_BT_NC],
                                               double (&rz)[EHIC_BT_NC])
{
    // qbase = first column of this group inside the tile; sd points at the group's slots;
    // row0 = first bead of this lane's rows (row rb is row0 + 32 rb); coarse = (2 r_max)^2
    const double alpha = dm.alpha;
    unsigned hits = 0u;                       // bit (qq * 4 + rb): pair within the coarse contact range
#pragma unroll
    for (int qq = 0; qq < EHIC_BT_NC; qq++) {
        const int q = qbase + qq;
        const double ox = __shfl_sync(0xffffffffu, cx, q), oy = __shfl_sync(0xffffffffu, cy, q),
                     oz = __shfl_sync(0xffffffffu, cz, q);
        double fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll
        for (int rb = 0; rb < 4; rb++) {
            const double dc = sd[(qq * 4 + rb) * 32], w = sd[1024 + (qq * 4 + rb) * 32];
            double dx = ox - xb[rb][0], dy = oy - xb[rb][1], dz = oz - xb[rb][2];
            double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
            double rd = rsqrt_nr(s2);
            double tt = alpha * (dc - s2 * rd);
            double rq = rsqrt_nr(fma(tt, tt, 1.0));
            double f = w * ((rq * rq) * (rq * rd));
            if (MASKED) f = ((m[rb] >> q) & 1u) ? f : 0.0;
            acc[rb][0] = fma(f, dx, acc[rb][0]); acc[rb][1] = fma(f, dy, acc[rb][1]); acc[rb][2] = fma(f, dz, acc[rb][2]);
            fx = fma(-f, dx, fx); fy = fma(-f, dy, fy); fz = fma(-f, dz, fz);
            if (NB) {
                bool h = s2 < coarse;
                if (MASKED) h = h && (col0 + q > row0 + 32 * rb) && (col0 + q < dm.N) && (row0 + 32 * rb < dm.N);
                hits |= h ? (1u << (qq * 4 + rb)) : 0u;
            }
        }
        rx[qq] = fx; ry[qq] = fy; rz[qq] = fz;
    }
    if (NB) {
        // rare: some pair of this chunk is within the coarse contact range.  One warp-wide OR
        // decides; only the flagged (column, row) combinations are redone, in warp-uniform branches.
        const unsigned any = __reduce_or_sync(0xffffffffu, hits);
        if (any != 0u) {
#pragma unroll
            for (int qq = 0; qq < EHIC_BT_NC; qq++) {
                if (((any >> (qq * 4)) & 15u) == 0u) continue;
                const int q = qbase + qq;
                const double ox = __shfl_sync(0xffffffffu, cx, q), oy = __shfl_sync(0xffffffffu, cy, q),
                             oz = __shfl_sync(0xffffffffu, cz, q), ro = __shfl_sync(0xffffffffu, cr, q);
#pragma unroll
                for (int rb = 0; rb < 4; rb++) {
                    if (((any >> (qq * 4 + rb)) & 1u) == 0u) continue;
                    double dx = ox - xb[rb][0], dy = oy - xb[rb][1], dz = oz - xb[rb][2];
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double d0 = dm.radii[min(row0 + 32 * rb, dm.N - 1)] + ro;
                    if (((hits >> (qq * 4 + rb)) & 1u) && s2 < d0 * d0) {
                        const double rd = rsqrt_nr(s2);
                        const double a = d0 - s2 * rd;
                        const double g = nbs * ((a * a) * (a * rd));
                        acc[rb][0] = fma(g, dx, acc[rb][0]); acc[rb][1] = fma(g, dy, acc[rb][1]); acc[rb][2] = fma(g, dz, acc[rb][2]);
                        rx[qq] = fma(-g, dx, rx[qq]); ry[qq] = fma(-g, dy, ry[qq]); rz[qq] = fma(-g, dz, rz[qq]);
                    }
                }
            }
        }
    }
}

template <bool NB>
__global__ void __launch_bounds__(128, 3)
backward_tile_kernel(DevModel dm, TileDev td, BackwardTileArgs ba)
{
    extern __shared__ __align__(16) double smem[];        // [STAGES][2][1024]: dc, w
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // row block outermost, ascending: row block I walks the tiles of columns >= 128 I, so the CTAs with the most
    // work of ALL replicas start first and the short ones fill the end of the launch (longest-first scheduling)
    const int kg = blockIdx.x % ba.n_kg;
    const int n_rep = gridDim.x / (ba.n_kg * td.n_rb);
    const int r = (blockIdx.x / ba.n_kg) % n_rep;
    const int I = blockIdx.x / (ba.n_kg * n_rep);
    const int k = kg * 4 + warp;
    const bool k_ok = k < dm.S;
    const long long rk = (long long)r * dm.S + (k_ok ? k : dm.S - 1);
    const int Np = dm.Np;
    const double *xk = ba.xs + rk * 3LL * Np;
    const int row0 = 128 * I + lane;
    double xb[4][3], acc[4][3];
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
        const int b0 = row0 + 32 * rb;
        const int b = b0 < dm.N ? b0 : 0;                // rows past the last bead compute on bead 0, masked out
        xb[rb][0] = xk[b]; xb[rb][1] = xk[Np + b]; xb[rb][2] = xk[2 * Np + b];
        acc[rb][0] = acc[rb][1] = acc[rb][2] = 0.0;
    }
    const double coarse = 4.0 * ba.r_max * ba.r_max;
    const double nbs = NB ? (ba.beta ? ba.beta[r] : 1.0) * 2.0 * dm.k_nb : 0.0;
    const double *twr = ba.tw + (long long)r * td.slots;
    double *Pk = ba.P + (rk * td.n_rb + I) * (long long)dm.N * 3;
    const int t0 = td.rb_begin[I], n_tiles = td.rb_begin[I + 1] - t0;
    const int n_chunks = n_tiles * 4;
    // stage ring filled by TMA bulk copies: thread 0 arms the stage's mbarrier with the byte count and
    // issues two 8 KB copies (contact distances, weights); everybody waits on the barrier's phase
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + EHIC_BT_STAGES * 2048);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < EHIC_BT_STAGES; q++) mbar_init(&full[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        if (c < n_chunks && threadIdx.x == 0) {
            const int st = c % EHIC_BT_STAGES;
            const long long base = (long long)(t0 + (c >> 2)) * 4096 + (c & 3) * 1024;
            double *d0 = smem + st * 2048;
            mbar_expect_tx(&full[st], 2 * 1024 * (unsigned)sizeof(double));
            bulk_g2s(d0, td.dc + base, 1024 * (unsigned)sizeof(double), &full[st]);
            bulk_g2s(d0 + 1024, twr + base, 1024 * (unsigned)sizeof(double), &full[st]);
        }
    };
#pragma unroll
    for (int c = 0; c < EHIC_BT_STAGES - 1; c++) issue(c);

    double cx = 0.0, cy = 0.0, cz = 0.0, cr = 0.0;
    unsigned m[4] = {0u, 0u, 0u, 0u};
    int col0 = 0;
    bool clean = false, nb_tile = false;
    // bounding box of this warp's 128 row beads, grown by the largest contact range: column slices
    // whose box lies outside cannot hold a nonbonded contact, their tiles skip the test altogether
    double rlo0 = 0.0, rlo1 = 0.0, rlo2 = 0.0, rhi0 = 0.0, rhi1 = 0.0, rhi2 = 0.0;
    const double *bbk = NB ? ba.bbox + rk * td.n_sl * 6 : nullptr;
    if (NB) {
        const double reach = 2.0 * ba.r_max * (1.0 + 1e-12);
        const int s0 = 4 * I, s1 = min(4 * I + 3, td.n_sl - 1);
        rlo0 = bbk[s0 * 6]; rlo1 = bbk[s0 * 6 + 1]; rlo2 = bbk[s0 * 6 + 2];
        rhi0 = bbk[s0 * 6 + 3]; rhi1 = bbk[s0 * 6 + 4]; rhi2 = bbk[s0 * 6 + 5];
        for (int q = s0 + 1; q <= s1; q++) {
            rlo0 = fmin(rlo0, bbk[q * 6]); rlo1 = fmin(rlo1, bbk[q * 6 + 1]); rlo2 = fmin(rlo2, bbk[q * 6 + 2]);
            rhi0 = fmax(rhi0, bbk[q * 6 + 3]); rhi1 = fmax(rhi1, bbk[q * 6 + 4]); rhi2 = fmax(rhi2, bbk[q * 6 + 5]);
        }
        rlo0 -= reach; rlo1 -= reach; rlo2 -= reach; rhi0 += reach; rhi1 += reach; rhi2 += reach;
    }
    for (int c = 0; c < n_chunks; c++) {
        mbar_wait(&full[c % EHIC_BT_STAGES], (unsigned)((c / EHIC_BT_STAGES) & 1));
        __syncthreads();                                  // everyone is done with chunk c-1: its stage may be refilled
        issue(c + EHIC_BT_STAGES - 1);
        const int qg = c & 3;
        if (qg == 0) {                                   // new tile: column beads (one per lane) and masks
            const int t = t0 + (c >> 2);
            col0 = 32 * td.tile_B[t];
            const int oc = min(col0 + lane, dm.N - 1);
            cx = xk[oc]; cy = xk[Np + oc]; cz = xk[2 * Np + oc];
            if (NB) {
                cr = dm.radii[oc];
                const double *pb = bbk + td.tile_B[t] * 6;
                nb_tile = !(pb[0] > rhi0 || pb[3] < rlo0 || pb[1] > rhi1 || pb[4] < rlo1 || pb[2] > rhi2 || pb[5] < rlo2);
            }
            clean = ba.tile_clean[t] != 0;
            if (!clean) {
#pragma unroll
                for (int rb = 0; rb < 4; rb++) m[rb] = td.mask[((long long)t * 4 + rb) * 32 + lane];
            }
        }
#pragma unroll
        for (int h = 0; h < 8 / EHIC_BT_NC; h++) {
            const int qbase = qg * 8 + h * EHIC_BT_NC;
            const double *sd = smem + (c % EHIC_BT_STAGES) * 2048 + h * EHIC_BT_NC * 128 + lane;
            double rx[EHIC_BT_NC], ry[EHIC_BT_NC], rz[EHIC_BT_NC];
            if (NB && nb_tile) {
                if (clean) backward_chunk<false, true>(dm, sd, qbase, lane, cx, cy, cz, cr, m, row0, col0, xb, coarse, nbs, acc, rx, ry, rz);
                else backward_chunk<true, true>(dm, sd, qbase, lane, cx, cy, cz, cr, m, row0, col0, xb, coarse, nbs, acc, rx, ry, rz);
            } else {
                if (clean) backward_chunk<false, false>(dm, sd, qbase, lane, cx, cy, cz, cr, m, row0, col0, xb, coarse, nbs, acc, rx, ry, rz);
                else backward_chunk<true, false>(dm, sd, qbase, lane, cx, cy, cz, cr, m, row0, col0, xb, coarse, nbs, acc, rx, ry, rz);
            }
#if EHIC_BT_NC == 8
            reduce8(rx, lane); reduce8(ry, lane); reduce8(rz, lane);
            const bool writer = (lane & 3) == 0;
            const int sub = ((lane >> 2) & 1) + 2 * ((lane >> 3) & 1) + 4 * ((lane >> 4) & 1);
#else
            reduce4(rx, lane); reduce4(ry, lane); reduce4(rz, lane);
            const bool writer = (lane & 7) == 0;
            const int sub = ((lane >> 3) & 1) + 2 * ((lane >> 4) & 1);
#endif
            if (writer && k_ok) {
                const int col = col0 + qbase + sub;
                if (col < dm.N) { double *o = Pk + (long long)col * 3; o[0] = rx[0]; o[1] = ry[0]; o[2] = rz[0]; }
            }
        }
    }
    if (k_ok) {
#pragma unroll
        for (int rb = 0; rb < 4; rb++) {
            const int b = row0 + 32 * rb;
            if (b < dm.N) { double *o = ba.Grow + (rk * dm.N + b) * 3; o[0] = acc[rb][0]; o[1] = acc[rb][1]; o[2] = acc[rb][2]; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Optional FP32 mode of the tile path (EHIC_FLAG_FP32; north star: 1e-5 relative).  Same mapping,
// staging and summation structure as the FP64 tile kernels; the pair arithmetic runs on the FP32
// pipe (128 lanes/clk/SM instead of 64) with MUFU.RSQ reciprocal square roots:
//   * coordinates are (hi, lo) float pairs (split_hilo), differences formed as (hi-hi) + (lo-lo);
//   * contact distances and weights are staged as floats (dc32 static, tw32 written by the forward);
//   * the smooth contact function of far-apart pairs (t < 0) is evaluated as
//     0.5 / (q^2 (1 + |t|/q)) instead of 0.5 (1 + t/q): no cancellation, so mock counts stay
//     relatively accurate where FP32 would otherwise lose every digit;
//   * per-column sums over the ensemble stay in FP32 (S terms); force accumulators are flushed
//     into FP64 after every tile (128 pairs per lane).
// Inputs, outputs, weights of the Poisson term, energies and the priors stay FP64.
template <typename T>
__device__ __forceinline__ void reduce8_t(T (&r)[8], int lane)
{
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        T send = hi ? r[i] : r[i + 4], keep = hi ? r[i + 4] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        T send = hi ? r[i] : r[i + 2], keep = hi ? r[i + 2] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = (lane & 4) != 0;
    {
        T send = hi ? r[0] : r[1], keep = hi ? r[1] : r[0];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 2);
    r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
}

static __host__ __device__ constexpr size_t forward_tile32_smem_bytes(int kt)
{
    return 4096 * sizeof(float) + sizeof(double) * ((size_t)EHIC_FT_STAGES * (128 + 32) * kt * 3) + 64;
}

template <int KT>
__global__ void __launch_bounds__(256, 3)
forward_tile32_kernel(DevModel dm, TileDev td, const float *__restrict__ dc32, const double *__restrict__ xt,
                      const double *__restrict__ norm, const double *__restrict__ lammda,
                      double *__restrict__ md_out, float *__restrict__ tw32,
                      double *__restrict__ partA, int want_logl, int nkt)
{
    constexpr int QC = 16;
    constexpr int REC = KT * 3;                       // (hi, lo) pairs per record
    extern __shared__ __align__(16) double smem[];
    float *s_dc = reinterpret_cast<float *>(smem);     // [4096]
    double *s_row = smem + 2048;                       // [STAGES][128 * REC]
    double *s_col = s_row + EHIC_FT_STAGES * 128 * REC;   // [STAGES][32 * REC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rb = warp & 3, h = warp >> 2;
    const int tile = blockIdx.x, r = blockIdx.y;
    const int I = td.tile_I[tile], B = td.tile_B[tile];
    const int Np = dm.Np;
    const int row0 = 128 * I, col0 = 32 * B;
    const double *recs = xt + (long long)r * nkt * Np * REC;
    const unsigned mask = td.mask[((long long)tile * 4 + rb) * 32 + lane] >> (h * QC);
    const bool warp_live = __ballot_sync(0xffffffffu, mask != 0u) != 0u;
    const float alpha = (float)dm.alpha;
    const int n_row = min(128, Np - row0);

    unsigned long long *full = reinterpret_cast<unsigned long long *>(s_col + EHIC_FT_STAGES * 32 * REC);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q <= EHIC_FT_STAGES; q++) mbar_init(&full[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&full[EHIC_FT_STAGES], 4096 * (unsigned)sizeof(float));
        bulk_g2s(s_dc, dc32 + (long long)tile * 4096, 16384, &full[EHIC_FT_STAGES]);
    }
    auto issue = [&](int kt) {
        if (kt < nkt && threadIdx.x == 0) {
            const int st = kt % EHIC_FT_STAGES;
            const double *src_row = recs + ((long long)kt * Np + row0) * REC;
            const double *src_col = recs + ((long long)kt * Np + col0) * REC;
            const unsigned row_bytes = (unsigned)(n_row * REC * sizeof(double)), col_bytes = (unsigned)(32 * REC * sizeof(double));
            mbar_expect_tx(&full[st], row_bytes + col_bytes);
            bulk_g2s(&s_row[st * 128 * REC], src_row, row_bytes, &full[st]);
            bulk_g2s(&s_col[st * 32 * REC], src_col, col_bytes, &full[st]);
        }
    };
#pragma unroll
    for (int c = 0; c < EHIC_FT_STAGES - 1; c++) issue(c);
    mbar_wait(&full[EHIC_FT_STAGES], 0u);

    float acc[QC];
#pragma unroll
    for (int q = 0; q < QC; q++) acc[q] = 0.0f;
    const int nfull = dm.S / KT;
    for (int kt = 0; kt < nkt; kt++) {
        mbar_wait(&full[kt % EHIC_FT_STAGES], (unsigned)((kt / EHIC_FT_STAGES) & 1));
        __syncthreads();
        issue(kt + EHIC_FT_STAGES - 1);
        if (!warp_live) continue;
        const int st = kt % EHIC_FT_STAGES;
        const int nvalid = kt < nfull ? KT : dm.S - nfull * KT;
        float2 own[REC];
        {
            const float2 *p = reinterpret_cast<const float2 *>(&s_row[st * 128 * REC + (rb * 32 + lane) * REC]);
#pragma unroll
            for (int u = 0; u < REC; u++) own[u] = p[u];
        }
        const float2 *pc = reinterpret_cast<const float2 *>(&s_col[st * 32 * REC + h * QC * REC]);
        const float *pd = &s_dc[(h * QC * 4 + rb) * 32 + lane];
#pragma unroll
        for (int q = 0; q < QC; q++) {
            const float dc = pd[q * 128];
#pragma unroll
            for (int t = 0; t < KT; t++) {
                if (t < nvalid) {                       // warp-uniform; always true except in the last structure tile
                    const float2 c0 = pc[q * REC + 3 * t], c1 = pc[q * REC + 3 * t + 1], c2 = pc[q * REC + 3 * t + 2];
                    const float dx = (c0.x - own[3 * t].x) + (c0.y - own[3 * t].y);
                    const float dy = (c1.x - own[3 * t + 1].x) + (c1.y - own[3 * t + 1].y);
                    const float dz = (c2.x - own[3 * t + 2].x) + (c2.y - own[3 * t + 2].y);
                    const float s2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float rd = rsqrt_f32(s2);
                    const float tt = alpha * (dc - s2 * rd);
                    const float u = fmaf(tt, tt, 1.0f);
                    const float rq = rsqrt_f32_raw(u);
                    const float a = fabsf(tt) * rq;      // |t| / q in [0, 1)
                    const float den = 1.0f + a;
                    acc[q] += tt >= 0.0f ? den : (rq * rq) * rcp_f32_raw(den);
                }
            }
        }
    }
    double term = 0.0;
    const double nrm = norm[r];
    const double lam = lammda ? lammda[r] : 1.0;
    const long long pos0 = (((long long)tile * 32 + h * QC) * 4 + rb) * 32 + lane;
#pragma unroll
    for (int q = 0; q < QC; q++) {
        if ((mask >> q) & 1u) {
            const long long pos = pos0 + q * 128;
            const double md = (double)acc[q] * (0.5 * nrm);
            const double c = td.cnt[pos];
            if (md_out) md_out[(long long)r * dm.M + td.orig[pos]] = md;
            if (tw32) {
                double em = __dsub_rn(1.0, __ddiv_rn(c, md));
                tw32[(long long)r * td.slots + pos] = (float)(lam * __dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), dm.alpha));
            }
            if (want_logl) term += c * log(md) - md;
        }
    }
    if (want_logl) {
        __shared__ double sh[8];
        double v = warp_sum(term);
        if (lane == 0) sh[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0)
            partA[(long long)r * gridDim.x + blockIdx.x] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
    }
}

struct BackwardTile32Args {
    const double *xs;                  // [R][S][3][Np] FP64 positions (split into hi/lo per CTA / per tile)
    const float *dc32, *tw32;          // [n_tiles*4096], [R][n_tiles*4096]
    double *Grow, *P;
    const unsigned char *tile_clean;
    int n_kg;
};

__global__ void __launch_bounds__(128, 4)
backward_tile32_kernel(DevModel dm, TileDev td, BackwardTile32Args ba)
{
    extern __shared__ __align__(16) double smem[];        // [STAGES][2][1024] floats: dc, w
    float *sf = reinterpret_cast<float *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // row block outermost, ascending: row block I walks the tiles of columns >= 128 I, so the CTAs with the most
    // work of ALL replicas start first and the short ones fill the end of the launch (longest-first scheduling)
    const int kg = blockIdx.x % ba.n_kg;
    const int n_rep = gridDim.x / (ba.n_kg * td.n_rb);
    const int r = (blockIdx.x / ba.n_kg) % n_rep;
    const int I = blockIdx.x / (ba.n_kg * n_rep);
    const int k = kg * 4 + warp;
    const bool k_ok = k < dm.S;
    const long long rk = (long long)r * dm.S + (k_ok ? k : dm.S - 1);
    const int Np = dm.Np;
    const double *xk = ba.xs + rk * 3LL * Np;
    const int row0 = 128 * I + lane;
    float xh[4][3], xl[4][3], acc[4][3];
    double acc64[4][3];
    auto split = [](double x, float &hi, float &lo) { hi = __double2float_rn(x); lo = __double2float_rn(x - (double)hi); };
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
        const int b0 = row0 + 32 * rb;
        const int b = b0 < dm.N ? b0 : 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            split(xk[c * Np + b], xh[rb][c], xl[rb][c]);
            acc[rb][c] = 0.0f; acc64[rb][c] = 0.0;
        }
    }
    const float alpha = (float)dm.alpha;
    const float *twr = ba.tw32 + (long long)r * td.slots;
    double *Pk = ba.P + (rk * td.n_rb + I) * (long long)dm.N * 3;
    const int t0 = td.rb_begin[I], n_tiles = td.rb_begin[I + 1] - t0;
    const int n_chunks = n_tiles * 4;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sf + EHIC_BT_STAGES * 2048);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < EHIC_BT_STAGES; q++) mbar_init(&full[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        if (c < n_chunks && threadIdx.x == 0) {
            const int st = c % EHIC_BT_STAGES;
            const long long base = (long long)(t0 + (c >> 2)) * 4096 + (c & 3) * 1024;
            float *d0 = sf + st * 2048;
            mbar_expect_tx(&full[st], 2 * 1024 * (unsigned)sizeof(float));
            bulk_g2s(d0, ba.dc32 + base, 1024 * (unsigned)sizeof(float), &full[st]);
            bulk_g2s(d0 + 1024, twr + base, 1024 * (unsigned)sizeof(float), &full[st]);
        }
    };
#pragma unroll
    for (int c = 0; c < EHIC_BT_STAGES - 1; c++) issue(c);

    float ch[3] = {0.f, 0.f, 0.f}, cl[3] = {0.f, 0.f, 0.f};
    unsigned m[4] = {0u, 0u, 0u, 0u};
    int col0 = 0;
    bool clean = false;
    for (int c = 0; c < n_chunks; c++) {
        mbar_wait(&full[c % EHIC_BT_STAGES], (unsigned)((c / EHIC_BT_STAGES) & 1));
        __syncthreads();
        issue(c + EHIC_BT_STAGES - 1);
        const int qg = c & 3;
        if (qg == 0) {
            const int t = t0 + (c >> 2);
            col0 = 32 * td.tile_B[t];
            const int oc = min(col0 + lane, dm.N - 1);
#pragma unroll
            for (int u = 0; u < 3; u++) split(xk[u * Np + oc], ch[u], cl[u]);
            clean = ba.tile_clean[t] != 0;
#pragma unroll
            for (int rb = 0; rb < 4; rb++) m[rb] = clean ? 0xffffffffu : td.mask[((long long)t * 4 + rb) * 32 + lane];
        }
        const float *sd = sf + (c % EHIC_BT_STAGES) * 2048 + lane;
        float rx[8], ry[8], rz[8];
#pragma unroll
        for (int qq = 0; qq < 8; qq++) {
            const int q = qg * 8 + qq;
            const float oxh = __shfl_sync(0xffffffffu, ch[0], q), oyh = __shfl_sync(0xffffffffu, ch[1], q),
                        ozh = __shfl_sync(0xffffffffu, ch[2], q);
            const float oxl = __shfl_sync(0xffffffffu, cl[0], q), oyl = __shfl_sync(0xffffffffu, cl[1], q),
                        ozl = __shfl_sync(0xffffffffu, cl[2], q);
            float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
            for (int rb = 0; rb < 4; rb++) {
                const float dc = sd[(qq * 4 + rb) * 32], w = sd[1024 + (qq * 4 + rb) * 32];
                const float dx = (oxh - xh[rb][0]) + (oxl - xl[rb][0]);
                const float dy = (oyh - xh[rb][1]) + (oyl - xl[rb][1]);
                const float dz = (ozh - xh[rb][2]) + (ozl - xl[rb][2]);
                const float s2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const float rd = rsqrt_f32(s2);
                const float tt = alpha * (dc - s2 * rd);
                const float rq = rsqrt_f32_raw(fmaf(tt, tt, 1.0f));
                float f = w * ((rq * rq) * (rq * rd));
                f = ((m[rb] >> q) & 1u) ? f : 0.0f;
                acc[rb][0] = fmaf(f, dx, acc[rb][0]); acc[rb][1] = fmaf(f, dy, acc[rb][1]); acc[rb][2] = fmaf(f, dz, acc[rb][2]);
                fx = fmaf(-f, dx, fx); fy = fmaf(-f, dy, fy); fz = fmaf(-f, dz, fz);
            }
            rx[qq] = fx; ry[qq] = fy; rz[qq] = fz;
        }
        reduce8_t(rx, lane); reduce8_t(ry, lane); reduce8_t(rz, lane);
        if ((lane & 3) == 0 && k_ok) {
            const int sub = ((lane >> 2) & 1) + 2 * ((lane >> 3) & 1) + 4 * ((lane >> 4) & 1);
            const int col = col0 + qg * 8 + sub;
            if (col < dm.N) { double *o = Pk + (long long)col * 3; o[0] = (double)rx[0]; o[1] = (double)ry[0]; o[2] = (double)rz[0]; }
        }
        if (qg == 3) {                                   // tile done: flush the FP32 row accumulators
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
#pragma unroll
                for (int u = 0; u < 3; u++) { acc64[rb][u] += (double)acc[rb][u]; acc[rb][u] = 0.0f; }
        }
    }
    if (k_ok) {
#pragma unroll
        for (int rb = 0; rb < 4; rb++) {
            const int b = row0 + 32 * rb;
            if (b < dm.N) { double *o = ba.Grow + (rk * dm.N + b) * 3; o[0] = acc64[rb][0]; o[1] = acc64[rb][1]; o[2] = acc64[rb][2]; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// FP32 mode of the LANE path (sparse lists; EHIC_FLAG_FP32, north star: 1e-5 relative).  Same mapping and
// summation structure as forward_lane_kernel / backward_lane_kernel; the pair arithmetic runs on the FP32 pipe
// with the same safeguards as the FP32 tile kernels: the lane copy xg holds (hi, lo) float pairs and differences
// are (hi - hi) + (lo - lo); the contact function of far-apart pairs avoids the 1 + t/q cancellation; per-pair sums
// over the ensemble stay in FP32 (S terms); the force accumulators of a bead are flushed into FP64 every 32 row
// entries.  Contact distances come from float copies of the two list views, weights travel as floats in the
// weight buffer; mock counts, Poisson weights, energies, priors and the leapfrog update stay FP64.
__device__ __forceinline__ void
forward_lane32_body(const DevModel &dm, const LaneDev &ld, const double *__restrict__ xg, const double *__restrict__ norm,
                    double *__restrict__ md_out, float *__restrict__ wbuf32,
                    double *__restrict__ partA, int want_logl, int vblock, int r, int tid, int group, double *sh /* [8] */,
                    int4 *stage /* [256] */)
{
    const int lane = tid & 31, warp = tid >> 5;
    const long long m0 = ((long long)vblock * 8 + warp) * 32;
    const double nrm = norm[r];
    const float alpha = (float)dm.alpha;
    const long long gstride = (long long)dm.Np * 96;
    const float2 *xr = reinterpret_cast<const float2 *>(xg) + (long long)r * ld.nkg * gstride + lane;
    double term = 0.0;
    if (m0 < dm.M) {                                    // warp-uniform
        // the warp's 32 pair records through its shared-memory slot, as in forward_lane_body
        int4 *se = stage + warp * 32;
        __syncwarp();
        se[lane] = __ldg(reinterpret_cast<const int4 *>(ld.pairs32 + min(m0 + lane, dm.M - 1)));
        __syncwarp();
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const long long mb = m0 + b * 8;
            if (mb < dm.M) {                            // warp-uniform
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; e++) acc[e] = 0.f;
                const int i_first = se[b * 8].x;
                for (int kg = 0; kg < ld.nkg; kg++) {
                    const float2 *xk = xr + kg * gstride;
                    const bool kv = kg * 32 + lane < dm.S;
                    const float2 *p0 = xk + (long long)i_first * 96;
                    float2 a0 = p0[0], a1 = p0[32], a2 = p0[64];
                    int ci = i_first;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int4 v = se[b * 8 + e];
                        if (v.x != ci) {                // warp-uniform, rare: the pairs are sorted by (i, j)
                            ci = v.x;
                            const float2 *p = xk + (long long)ci * 96;
                            a0 = p[0]; a1 = p[32]; a2 = p[64];
                        }
                        const float2 *q = xk + (long long)v.y * 96;
                        const float2 q0 = q[0], q1 = q[32], q2 = q[64];
                        const float dx = (a0.x - q0.x) + (a0.y - q0.y);
                        const float dy = (a1.x - q1.x) + (a1.y - q1.y);
                        const float dz = (a2.x - q2.x) + (a2.y - q2.y);
                        const float s2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        const float rd = rsqrt_f32(s2);
                        const float tt = alpha * (__int_as_float(v.z) - s2 * rd);
                        const float rq = rsqrt_f32_raw(fmaf(tt, tt, 1.0f));
                        const float a = fabsf(tt) * rq;      // |t| / q in [0, 1)
                        const float den = 1.0f + a;
                        const float sv = tt >= 0.0f ? den : (rq * rq) * rcp_f32_raw(den);
                        acc[e] += kv ? sv : 0.0f;
                    }
                }
                reduce8_t(acc, lane);                   // lane L holds the total of pair mb + (L >> 2)
                tot[b] = acc[0];
            }
        }
        // lane L finishes pair m0 + 8 (L & 3) + (L >> 2)
        const int bsel = lane & 3;
        const float sum = bsel == 0 ? tot[0] : (bsel == 1 ? tot[1] : (bsel == 2 ? tot[2] : tot[3]));
        const long long m = m0 + 8 * bsel + (lane >> 2);
        if (m < dm.M) {
            const double c = dm.pcnt[m];
            const double md = (double)sum * (0.5 * nrm);
            if (md_out) md_out[(long long)r * dm.M + dm.perm[m]] = md;
            const double em = __dsub_rn(1.0, __ddiv_rn(c, md));
            const float w = (float)__dmul_rn(__dmul_rn(__dmul_rn(0.5, em), nrm), dm.alpha);
            if (wbuf32) {
                float *wr = wbuf32 + (long long)r * dm.wstride * 2;      // the replica's slice of the (double-sized) weight buffer
                wr[ld.slot_a[m]] = w;
                wr[ld.slot_b[m]] = w;
            }
            if (want_logl) term = c * log(md) - md;
        }
    }
    if (want_logl) {
        double v = warp_sum(term);
        if (lane == 0) sh[warp] = v;
        group_sync(group, 256);
        if (tid == 0)
            partA[(long long)r * ((dm.M + 255) / 256) + vblock] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
        group_sync(group, 256);
    }
}

__global__ void __launch_bounds__(256, 3)
forward_lane32_kernel(DevModel dm, LaneDev ld, const double *__restrict__ xg, const double *__restrict__ norm,
                      double *__restrict__ md_out, float *__restrict__ wbuf32,
                      double *__restrict__ partA, int want_logl)
{
    __shared__ double sh[8];
    __shared__ int4 stage[256];
    forward_lane32_body(dm, ld, xg, norm, md_out, wbuf32, partA, want_logl, blockIdx.x, blockIdx.y, threadIdx.x, 0, sh, stage);
}

__global__ void __launch_bounds__(EHIC_LANE_FWD_GROUPS * 256, 1)
forward_lane32_resident_kernel(DevModel dm, LaneDev ld, LaneSched sc, const double *__restrict__ xg, const double *__restrict__ norm,
                               double *__restrict__ md_out, float *__restrict__ wbuf32,
                               double *__restrict__ partA, int want_logl)
{
    __shared__ double sh[EHIC_LANE_FWD_GROUPS][8];
    __shared__ int4 stage[EHIC_LANE_FWD_GROUPS][256];
    const int group = threadIdx.x >> 8, tid = threadIdx.x & 255;
    long long v0, v1;
    lane_resident_range(sc, v0, v1);
    for (long long v = v0 + group; v < v1; v += EHIC_LANE_FWD_GROUPS) {
        const int r = (int)(v / sc.per_replica), vb = (int)(v - (long long)r * sc.per_replica);
        forward_lane32_body(dm, ld, xg, norm, md_out, wbuf32, partA, want_logl, vb, r, tid, group, sh[group], stage[group]);
    }
}

__device__ __forceinline__ void
backward_lane32_body(const DevModel &dm, const LaneDev &ld, const GatherArgs &ga, int n_quads, long long vblock, int tid, int group,
                     int4 *stage_ent /* [EHIC_LANE_WARPS * 32] */, double (*part)[3][32] /* [EHIC_LANE_WARPS] */)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int quad = (int)(vblock % n_quads);
    const int tmp = (int)(vblock / n_quads);
    const int kg = tmp % ld.nkg;
    const int r = tmp / ld.nkg;
    const int4 item = __ldg(ld.items + quad * EHIC_LANE_WARPS + warp);
    const bool bead_ok = item.x >= 0;
    const int b = bead_ok ? item.x : dm.N - 1;
    const int k = kg * 32 + lane;
    const bool k_ok = k < dm.S;
    const long long gstride = (long long)dm.Np * 96;
    const float2 *xk = reinterpret_cast<const float2 *>(ga.xt) + ((long long)r * ld.nkg + kg) * gstride + lane;
    const float2 x0 = xk[(long long)b * 96], x1 = xk[(long long)b * 96 + 32], x2 = xk[(long long)b * 96 + 64];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    if (ga.contact && dm.M > 0 && bead_ok) {            // warp-uniform
        const long long e0 = item.y, e1 = item.z;        // this warp's segment of the bead's CSR row
        const float *wr = reinterpret_cast<const float *>(ga.wbuf) + (long long)r * dm.wstride * 2;
        const float alpha = (float)dm.alpha;
        float f0 = 0.f, f1 = 0.f, f2 = 0.f;
        // metadata in blocks of 32 entries through the warp's shared-memory slot, as in backward_lane_body
        int4 *se = stage_ent + (tid >> 5) * 32;
        int4 nv = make_int4(0, 0, 0, 0);
        if (e0 < e1) {
            const long long e = min(e0 + lane, e1 - 1);
            const int2 t = __ldg(reinterpret_cast<const int2 *>(ld.ent32 + e));
            nv = make_int4(t.x, t.y, __float_as_int(wr[e]), 0);
        }
        for (long long eb = e0; eb < e1; eb += 32) {    // FP32 partial sums of at most 32 entries
            __syncwarp();
            se[lane] = nv;
            __syncwarp();
            if (eb + 32 < e1) {
                const long long e = min(eb + 32 + lane, e1 - 1);
                const int2 t = __ldg(reinterpret_cast<const int2 *>(ld.ent32 + e));
                nv = make_int4(t.x, t.y, __float_as_int(wr[e]), 0);
            }
            const int n = (int)min(32LL, e1 - eb);
#pragma unroll 4
            for (int qe = 0; qe < n; qe++) {
                const int4 v = se[qe];
                const float dc = __int_as_float(v.x);
                const float w = __int_as_float(v.z);
                const float2 *q = xk + (long long)v.y * 96;
                const float2 q0 = q[0], q1 = q[32], q2 = q[64];
                const float dx = (q0.x - x0.x) + (q0.y - x0.y);
                const float dy = (q1.x - x1.x) + (q1.y - x1.y);
                const float dz = (q2.x - x2.x) + (q2.y - x2.y);
                const float s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const float rd = rsqrt_f32(s);
                const float tt = alpha * (dc - s * rd);
                const float rq = rsqrt_f32_raw(fmaf(tt, tt, 1.0f));
                const float f = w * ((rq * rq) * (rq * rd));
                f0 = fmaf(f, dx, f0); f1 = fmaf(f, dy, f1); f2 = fmaf(f, dz, f2);
            }
            acc0 += (double)f0; acc1 += (double)f1; acc2 += (double)f2;
            f0 = f1 = f2 = 0.f;
        }
    }
    // positions for the drift come from the FP64 SoA copy inside the epilogue
    const bool leader = lane_combine_segments(part, group, warp, lane, item.w & 255, item.w >> 8, acc0, acc1, acc2);
    lane_epilogue<true>(dm, ld, ga, vblock, r, kg, k, b, lane, warp, k_ok && bead_ok && leader, gstride, 0.0, 0.0, 0.0, acc0, acc1, acc2);
}

__global__ void __launch_bounds__(EHIC_LANE_WARPS * 32)
backward_lane32_kernel(DevModel dm, LaneDev ld, GatherArgs ga, int n_quads)
{
    __shared__ int4 s_ent[EHIC_LANE_WARPS * 32];
    __shared__ double s_part[EHIC_LANE_WARPS][3][32];
    backward_lane32_body(dm, ld, ga, n_quads, blockIdx.x, threadIdx.x, 0, s_ent, s_part);
}

__global__ void __launch_bounds__(EHIC_LANE_BWD_GROUPS * EHIC_LANE_WARPS * 32, 1)
backward_lane32_resident_kernel(DevModel dm, LaneDev ld, LaneSched sc, GatherArgs ga, int n_quads)
{
    const int group = threadIdx.x / (EHIC_LANE_WARPS * 32), tid = threadIdx.x % (EHIC_LANE_WARPS * 32);
    long long v0, v1;
    lane_resident_range(sc, v0, v1);
    __shared__ int4 s_ent[EHIC_LANE_BWD_GROUPS][EHIC_LANE_WARPS * 32];
    __shared__ double s_part[EHIC_LANE_BWD_GROUPS][EHIC_LANE_WARPS][3][32];
    for (long long v = v0 + group; v < v1; v += EHIC_LANE_BWD_GROUPS) {
        backward_lane32_body(dm, ld, ga, n_quads, v, tid, group, s_ent[group], s_part[group]);
        group_sync(group, EHIC_LANE_WARPS * 32);
    }
}

// Combine + epilogue of the tile path: thread <-> (replica, structure, bead).
// g = (row sum + sum over row blocks of the reaction partials, ascending) + gprior,
// then the same leapfrog epilogue as gather_kernel.
__global__ void __launch_bounds__(256)
combine_kernel(DevModel dm, TileDev td, GatherArgs ga, const double *__restrict__ Grow, const double *__restrict__ P,
               int contact, int KT, int nkt, int f32)
{
    const int k = blockIdx.y, r = blockIdx.z;
    const int b = blockIdx.x * 256 + threadIdx.x;
    const long long rk = (long long)r * dm.S + k;
    double kin = 0.0;
    if (b < dm.N) {
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        const long long at = (rk * dm.N + b) * 3;
        if (contact) {
            g0 = Grow[at]; g1 = Grow[at + 1]; g2 = Grow[at + 2];
            const int sl = b >> 5;
            for (int I = 0; I <= (sl >> 2); I++) {
                if (td.exists[I * td.n_sl + sl] >= 0) {
                    const double *o = P + ((rk * td.n_rb + I) * (long long)dm.N + b) * 3;
                    g0 += o[0]; g1 += o[1]; g2 += o[2];
                }
            }
        }
        // lammda (and beta for the fused nonbonded force) are already inside the tile sums
        if (ga.gprior) { g0 += ga.gprior[at]; g1 += ga.gprior[at + 1]; g2 += ga.gprior[at + 2]; }
        if (ga.grad) { ga.grad[at] = g0; ga.grad[at + 1] = g1; ga.grad[at + 2] = g2; }
        if (ga.p) {
            const double eps = ga.eps[r];
            double p0 = ga.p[at], p1 = ga.p[at + 1], p2 = ga.p[at + 2];
            if (ga.kinetic_mode == 1) kin = p0 * p0 + p1 * p1 + p2 * p2;
            const double ke = ga.kick * eps;
            p0 = fma(-ke, g0, p0); p1 = fma(-ke, g1, p1); p2 = fma(-ke, g2, p2);
            ga.p[at] = p0; ga.p[at + 1] = p1; ga.p[at + 2] = p2;
            if (ga.kinetic_mode == 2) kin = p0 * p0 + p1 * p1 + p2 * p2;
            if (ga.xs_next) {
                const int Np = dm.Np;
                const double *xc = ga.xs_cur + rk * 3LL * Np;
                const double n0 = fma(eps, p0, xc[b]), n1 = fma(eps, p1, xc[Np + b]), n2 = fma(eps, p2, xc[2 * Np + b]);
                double *xn = ga.xs_next + rk * 3LL * Np;
                xn[b] = n0; xn[Np + b] = n1; xn[2 * Np + b] = n2;
                double *un = ga.xt_next + ((((long long)r * nkt + k / KT) * Np + b) * KT + (k % KT)) * 3;
                if (f32) { un[0] = split_hilo(n0); un[1] = split_hilo(n1); un[2] = split_hilo(n2); }
                else { un[0] = n0; un[1] = n1; un[2] = n2; }
            }
        }
    }
    if (ga.partK) {
        __shared__ double sh[8];
        double v = warp_sum(kin);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) t += sh[q];
            ga.partK[rk * gridDim.x + blockIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fixed-order reduction of the per-block / per-warp partials into E[R][4] and K[R].
__global__ void __launch_bounds__(256)
finalize_kernel(DevModel dm, const double *__restrict__ partA, int nA,
                const double *__restrict__ partP, long long nP,
                const double *__restrict__ partK, long long nK,
                const double *__restrict__ partC, int nC,
                double *__restrict__ energies, double *__restrict__ kinetic)
{
    const int r = blockIdx.x;
    __shared__ double sh[5][256];
    double a = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    if (partA)
        for (int q = threadIdx.x; q < nA; q += 256) a += partA[(long long)r * nA + q];
    if (partP)
        for (long long q = threadIdx.x; q < nP; q += 256) {
            const double *o = partP + ((long long)r * nP + q) * 3;
            e0 += o[0]; e1 += o[1]; e2 += o[2];
        }
    if (partK)
        for (long long q = threadIdx.x; q < nK; q += 256) e3 += partK[(long long)r * nK + q];
    if (partC)      // per structure: (nonbonded, backbone, sphere) from the cell-grid / Verlet-list kernels
        for (int q = threadIdx.x; q < nC; q += 256) {
            const double *o = partC + ((long long)r * nC + q) * 3;
            e0 += o[0]; e1 += o[1]; e2 += o[2];
        }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = e0; sh[2][threadIdx.x] = e1;
    sh[3][threadIdx.x] = e2; sh[4][threadIdx.x] = e3;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < 5; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (energies) {
            double *E = energies + (long long)r * 4;
            E[0] = -sh[0][0];                         // -log L
            E[1] = 0.25 * dm.k_nb * sh[1][0];         // each pair visited from both ends: k/2 * 1/2
            E[2] = 0.5 * dm.k_bb * sh[2][0];
            E[3] = 0.5 * dm.sph_k * sh[3][0];
        }
        if (kinetic) kinetic[r] = 0.5 * sh[4][0];
    }
}

// H[r][slot] = lammda*E0 + beta*E1 + E2 + E3 ; H[r][slot+1] = K
__global__ void hamiltonian_kernel(const double *__restrict__ energies, const double *__restrict__ kinetic,
                                   const double *__restrict__ lammda, const double *__restrict__ beta,
                                   double *__restrict__ H, int slot, int R)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double *E = energies + (long long)r * 4;
    double lam = lammda ? lammda[r] : 1.0, bet = beta ? beta[r] : 1.0;
    H[(long long)r * 4 + slot] = lam * E[0] + bet * E[1] + E[2] + E[3];
    H[(long long)r * 4 + slot + 1] = kinetic[r];
}

__global__ void select_kernel(int R, long long n, const int *__restrict__ mask,
                              const double *__restrict__ src, double *__restrict__ dst)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)R * n) return;
    int r = (int)(t / n);
    if (mask[r]) dst[t] = src[t];
}

// ---------------------------------------------------------------------------------------
// Whole-trajectory kernel for SMALL systems (north-star kernel 3): ONE launch = one HMC proposal.
// CTA <-> replica.  The replica's coordinates (all S structures) and the per-slot Poisson weights
// live in shared memory for the whole trajectory, momenta in global memory (each element has one
// owner thread).  Per leapfrog step: forward (thread <-> data point, serial over the ensemble),
// barrier, gradient (thread <-> (structure, bead): contact row + all-pairs nonbonded + backbone +
// sphere, same arithmetic as gather_kernel / prior_kernel in fast mode) fused with the kick,
// barrier, drift, barrier.  Used when S*N*24 B + slots*8 B fits the 227 KB of one SM: hairpin_s,
// proteins, nora2012 15 kb; larger systems take the multi-kernel path (captured in a CUDA graph).
struct TrajSmallArgs {
    double *x, *p;                 // [R][S][N][3] in/out
    const double *norm, *lammda, *beta, *eps;
    double *H;                     // [R][4]
    int n_steps;
    double r_max;
    int split;                     // threads per (structure, bead) item: power of two <= 32
};

__device__ __forceinline__ double block_sum(double v, double *sh /* [32] */)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += sh[q];
    return t;
}

__global__ void __launch_bounds__(512)
trajectory_small_kernel(DevModel dm, TrajSmallArgs ta)
{
    extern __shared__ __align__(16) double smem[];
    const int r = blockIdx.x;
    const int N = dm.N, S = dm.S;
    const int ndof = S * N * 3;
    double *sx = smem;                         // [S][N][3]
    double *sw = sx + ndof;                    // [tot] row-slot weights (lammda already applied)
    double *red = sw + dm.tot;                 // [32]
    double *gx = ta.x + (long long)r * ndof, *gp = ta.p + (long long)r * ndof;
    const double nrm = ta.norm[r];
    const double lam = ta.lammda ? ta.lammda[r] : 1.0, bet = ta.beta ? ta.beta[r] : 1.0;
    const double eps = ta.eps[r];
    const double alpha = dm.alpha;
    const int n_slots = dm.n_slices * 32;

    for (int q = threadIdx.x; q < ndof; q += blockDim.x) sx[q] = gx[q];
    for (long long q = threadIdx.x; q < dm.tot; q += blockDim.x) sw[q] = 0.0;
    __syncthreads();

    for (int step = 0; step <= ta.n_steps; step++) {
        const bool first = step == 0, last = step == ta.n_steps, ends = first || last;
        // ---- forward: md, weights, log L ------------------------------------------------------
        double logl = 0.0;
        for (long long m = threadIdx.x; m < dm.M; m += blockDim.x) {
            const int i = dm.pi[m], j = dm.pj[m];
            const double dc = dm.pdc[m], c = dm.pcnt[m];
            double acc = 0.0;
            for (int k = 0; k < S; k++) {
                const double *a = sx + (k * N + i) * 3, *b = sx + (k * N + j) * 3;
                double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
                double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                double rd = rsqrt_nr(s2);
                double tt = alpha * (dc - s2 * rd);
                double rq = rsqrt_nr(fma(tt, tt, 1.0));
                acc += fma(tt, rq, 1.0);
            }
            const double md = acc * (0.5 * nrm);
            const double em = 1.0 - c / md;
            const double w = lam * (((0.5 * em) * nrm) * alpha);
            sw[dm.slot_a[m]] = w; sw[dm.slot_b[m]] = w;
            if (ends) logl += c * log(md) - md;
        }
        __syncthreads();
        // ---- gradient of every (structure, bead) + kick ------------------------------------------
        double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0, kin = 0.0;
        const double kick = (ends ? 0.5 : 1.0) * eps;
        // each item (structure, bead) is shared by T = split adjacent lanes: lane `sub` takes every
        // T-th row entry / partner; the partial forces are combined with xor shuffles (fixed order)
        const int T = ta.split, sub = threadIdx.x & (T - 1);
        const int groups = blockDim.x / T;
        for (int base = 0; base < S * n_slots; base += groups) {
            const int item = base + threadIdx.x / T;
            const int k = item < S * n_slots ? item / n_slots : 0;
            const int slot = item < S * n_slots ? item - k * n_slots : 0;
            const int b_raw = item < S * n_slots ? dm.row_bead[slot] : -1;
            const bool valid = b_raw >= 0;
            const int b = valid ? b_raw : 0;
            const double *xk = sx + k * N * 3;
            const double xb0 = xk[3 * b], xb1 = xk[3 * b + 1], xb2 = xk[3 * b + 2];
            double g0 = 0.0, g1 = 0.0, g2 = 0.0;
            if (valid) {   // contact row (weights carry lammda)
                const int sl = slot >> 5, lane = slot & 31;
                const long long off = dm.slice_off[sl] + lane;
                const int width = dm.slice_w[sl];
                for (int e = sub; e < width; e += T) {
                    const int o = dm.ent_j[off + 32 * e];
                    const double dc = dm.ent_dc[off + 32 * e], w = sw[off + 32 * e];
                    double dx = xk[3 * o] - xb0, dy = xk[3 * o + 1] - xb1, dz = xk[3 * o + 2] - xb2;
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    double rd = rsqrt_nr(s2);
                    double tt = alpha * (dc - s2 * rd);
                    double rq = rsqrt_nr(fma(tt, tt, 1.0));
                    double f = w * ((rq * rq) * (rq * rd));
                    g0 = fma(f, dx, g0); g1 = fma(f, dy, g1); g2 = fma(f, dz, g2);
                }
            }
            const double rb = dm.radii[b];
            if (valid) {   // nonbonded, partners sub, sub + T, ...
                const double coarse = (rb + ta.r_max) * (rb + ta.r_max);
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int o = sub; o < N; o += T) {
                    double dx = xk[3 * o] - xb0, dy = xk[3 * o + 1] - xb1, dz = xk[3 * o + 2] - xb2;
                    double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (s2 < coarse && o != b) {
                        const double d0 = rb + dm.radii[o];
                        if (s2 < d0 * d0) {
                            double rd = rsqrt_nr(s2);
                            double a = d0 - s2 * rd, aa = a * a;
                            double g = aa * a * rd;
                            a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
                            e_nb += aa * aa;
                        }
                    }
                }
                const double sc = bet * (dm.k_nb * 2.0);
                g0 = fma(sc, a0, g0); g1 = fma(sc, a1, g1); g2 = fma(sc, a2, g2);
            }
            for (int o = 1; o < T; o <<= 1) {      // all lanes of the warp take part: trip count is uniform
                g0 += __shfl_xor_sync(0xffffffffu, g0, o);
                g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                g2 += __shfl_xor_sync(0xffffffffu, g2, o);
            }
            if (!valid || sub != 0) continue;     // lane 0 of the group finishes the bead
            {   // backbone
                double r0 = 0.0, r1 = 0.0, r2 = 0.0;
                if (b > 0 && !isinf(dm.bb_ul[b - 1])) {
                    const double ul = dm.bb_ul[b - 1], ll = dm.bb_ll[b - 1];
                    double d0 = xb0 - xk[3 * b - 3], d1 = xb1 - xk[3 * b - 2], d2 = xb2 - xk[3 * b - 1];
                    double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                    if (s2 > ul * ul) { double d = sqrt(s2), a = d - ul; r0 = a * d0 / d; r1 = a * d1 / d; r2 = a * d2 / d; e_bb += a * a; }
                    else if (s2 < ll * ll) { double d = sqrt(s2), a = ll - d; r0 = -(a * d0 / d); r1 = -(a * d1 / d); r2 = -(a * d2 / d); e_bb += a * a; }
                }
                if (!isinf(dm.bb_ul[b])) {
                    const double ul = dm.bb_ul[b], ll = dm.bb_ll[b];
                    double d0 = xk[3 * b + 3] - xb0, d1 = xk[3 * b + 4] - xb1, d2 = xk[3 * b + 5] - xb2;
                    double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                    if (s2 > ul * ul) { double d = sqrt(s2), a = d - ul; r0 -= a * d0 / d; r1 -= a * d1 / d; r2 -= a * d2 / d; }
                    else if (s2 < ll * ll) { double d = sqrt(s2), a = ll - d; r0 += a * d0 / d; r1 += a * d1 / d; r2 += a * d2 / d; }
                }
                g0 = fma(dm.k_bb, r0, g0); g1 = fma(dm.k_bb, r1, g1); g2 = fma(dm.k_bb, r2, g2);
            }
            if (dm.sphere_active) {
                const double R0 = dm.sph_R;
                double d0 = xb0 - dm.sph_cx, d1 = xb1 - dm.sph_cy, d2 = xb2 - dm.sph_cz;
                double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                if (!(s2 < (R0 * R0 + rb * rb) - (2.0 * R0) * rb)) {
                    double d = sqrt(s2), a = (d + rb) - R0, f = a / d;
                    g0 = fma(dm.sph_k, d0 * f, g0); g1 = fma(dm.sph_k, d1 * f, g1); g2 = fma(dm.sph_k, d2 * f, g2);
                    if (a > 0.0) e_sp += a * a;
                }
            }
            const int at = (k * N + b) * 3;
            double p0 = gp[at], p1 = gp[at + 1], p2 = gp[at + 2];
            if (first) kin += p0 * p0 + p1 * p1 + p2 * p2;
            p0 = fma(-kick, g0, p0); p1 = fma(-kick, g1, p1); p2 = fma(-kick, g2, p2);
            gp[at] = p0; gp[at + 1] = p1; gp[at + 2] = p2;
            if (last) kin += p0 * p0 + p1 * p1 + p2 * p2;
        }
        if (ends) {
            const double tl = block_sum(logl, red), tn = block_sum(e_nb, red), tb = block_sum(e_bb, red),
                         ts = block_sum(e_sp, red), tk = block_sum(kin, red);
            if (threadIdx.x == 0) {
                double *H = ta.H + (long long)r * 4 + (first ? 0 : 2);
                H[0] = lam * (-tl) + bet * (0.25 * dm.k_nb * tn) + 0.5 * dm.k_bb * tb + 0.5 * dm.sph_k * ts;
                H[1] = 0.5 * tk;
            }
        }
        __syncthreads();                       // everybody is done reading sx
        if (!last) {
            for (int q = threadIdx.x; q < ndof; q += blockDim.x) sx[q] = fma(eps, gp[q], sx[q]);
            __syncthreads();
        }
    }
    for (int q = threadIdx.x; q < ndof; q += blockDim.x) gx[q] = sx[q];
}

// ---------------------------------------------------------------------------------------
// Neighbour list of one structure (c/nblist.c:181-303): membership |dx|^2 <= cutoff^2 with
// the dot product associated as mathutils.c:13-15 and no FMA contraction.
__global__ void nblist_count_kernel(const double *__restrict__ x, int N, double cut2, int *__restrict__ counts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int c = 0;
    for (int j = i + 1; j < N; j++) {
        double d0 = __dsub_rn(xi, x[3 * j]), d1 = __dsub_rn(yi, x[3 * j + 1]), d2 = __dsub_rn(zi, x[3 * j + 2]);
        double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        if (!(s > cut2)) c++;
    }
    counts[i] = c;
}

__global__ void nblist_fill_kernel(const double *__restrict__ x, int N, double cut2,
                                   const long long *__restrict__ offsets, int *__restrict__ pairs, long long capacity)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    long long w = offsets[i];
    for (int j = i + 1; j < N; j++) {
        double d0 = __dsub_rn(xi, x[3 * j]), d1 = __dsub_rn(yi, x[3 * j + 1]), d2 = __dsub_rn(zi, x[3 * j + 2]);
        double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        if (!(s > cut2)) {
            if (w < capacity) { pairs[2 * w] = i; pairs[2 * w + 1] = j; }
            w++;
        }
    }
}

}  // namespace ehic
