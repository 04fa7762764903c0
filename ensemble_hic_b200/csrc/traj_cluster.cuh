// traj_cluster.cuh -- whole-trajectory kernel for MEDIUM systems (north-star kernel 3): ONE launch = one HMC proposal,
// one thread-block CLUSTER per replica.
//
// trajectory_small_kernel keeps a replica's whole state in the shared memory of one SM; the reference's production
// shape (nora2012 3 kb: 308 beads x 30 structures, M = 10 145; config_files/nora2012/30structures_3kb.cfg) needs
// 222 KB of coordinates alone and would leave most SMs idle when a GPU holds few replicas (strong scaling of the
// ladder).  Here the S structures of a replica are SPLIT over the C = 2 / 4 / 8 CTAs of a cluster:
//
//   * CTA c holds the coordinates of its Sc = ceil(S / C) structures in shared memory for the whole trajectory,
//     laid out [bead][structure][xyz], and a lane is a STRUCTURE: a warp works on 32 / LK bead pairs at a time for
//     LK >= Sc structures each, so everything about a pair is uniform over LK lanes and partner coordinates are
//     conflict-free shared-memory reads instead of L2 gathers (the lane kernels of the multi-kernel path spend
//     their time in long-scoreboard stalls on exactly those gathers, profiles/r01_nora_lanepath.txt);
//   * the ensemble sum of the forward model (evaluate_FWM_pureC.pxd:29-39) crosses the CTAs through distributed
//     shared memory: every CTA writes its partial sums of a chunk of pairs into the buffer of the pair's OWNER CTA
//     (remote stores), cluster barrier, the owner adds the C partials in rank order (deterministic), forms the Poisson
//     weight (likelihoods_c.pyx:24-25,69) and stores it into the weight table W[M] of ALL CTAs (remote stores).
//     Chunks are double-buffered: one cluster barrier per chunk plus one before the gradient phase;
//   * the gradient phase is local to a CTA: warp <-> bead (longest contact rows first, dealt round robin), gather form over the bead's CSR row (likelihoods_c.pyx:62-73, every pair visited from both beads: no
//     atomics on floating-point data, fixed summation order), weights from W, then the nonbonded term from per-structure
//     Verlet lists (global memory, one 128-byte line per bead; rebuilt inside the kernel when a conservative bound on
//     the displacement since the last build exceeds skin / 2), backbone and sphere terms, kick.  Momenta live in
//     global memory in the same [bead][structure][xyz] order (coalesced), the drift updates the shared coordinates.
//
// Results: same arithmetic as the lane / prior kernels of the multi-kernel path (Newton-refined rsqrt, fast mode);
// only summation orders differ (tests hold 1e-12 against the multi-kernel trajectory and the oracle).
#pragma once
#include <cooperative_groups.h>
#include "kernels.cuh"

namespace ehic {

namespace cg = cooperative_groups;

#ifndef EHIC_TC_THREADS
#define EHIC_TC_THREADS 512
#endif

struct TrajClusterArgs {
    double *x, *p;                 // [R][S][N][3] in/out (public layout)
    const double *norm, *lammda, *beta, *eps;
    double *H;                     // [R][4]
    double *pw;                    // [R][C][N*Sc*3] momenta in kernel order (workspace)
    int n_steps;
    int C, Sc;                     // CTAs per replica, structures per CTA
    int chunk;                     // pairs per forward chunk (multiple of C)
    double r_max, skin;
    unsigned short *vl_idx;        // [R*S][Np][EHIC_VL_CAP]
    int *vl_cnt;                   // [R*S][Np]; > EHIC_VL_CAP = overflow: the bead sweeps all partners
    int *vl_state;                 // [R*S]: reset to 0 ("never built") on exit -- the lists are private to this launch
};

template <int LK>
__device__ __forceinline__ double group_sum(double v)
{   // sum over aligned groups of LK lanes, fixed butterfly order
#pragma unroll
    for (int o = LK >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Halving butterfly over aligned groups of LK lanes: v[t] is every lane's contribution to column t (LK columns);
// afterwards v[0] of the lane with index kl inside its group is the group total of column kl.  LK - 1 shuffle-adds
// instead of LK log2(LK); fixed order (deterministic).
template <int LK>
__device__ __forceinline__ void halving_reduce(double (&v)[LK], int lane)
{
#pragma unroll
    for (int h = LK / 2; h > 0; h >>= 1) {
        const bool hi = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; i++) {
            const double send = hi ? v[i] : v[i + h], keep = hi ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
}

__device__ __forceinline__ double block_sum512(double v, double *sh /* [32] */)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < EHIC_TC_THREADS / 32; q++) t += sh[q];
    return t;
}

template <int LK>
__global__ void __launch_bounds__(EHIC_TC_THREADS, 1)
trajectory_cluster_kernel(DevModel dm, LaneDev ld, TrajClusterArgs ta)
{
    constexpr int PW = 32 / LK;                        // bead pairs / row entries a warp works on at a time
    constexpr int NW = EHIC_TC_THREADS / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = ta.C, crank = (int)cluster.block_rank();
    const int r = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = dm.N, S = dm.S, Np = dm.Np, Sc = ta.Sc, M = (int)dm.M;
    const int k0 = crank * (S / C) + min(crank, S % C);      // balanced split: the first S % C CTAs hold one structure more
    const int nloc = S / C + (crank < S % C ? 1 : 0);        // structures held here (>= 1: the host requires S >= C)
    const int rec = Sc * 3;                            // doubles per bead record (Sc = ceil(S / C))
    const int ndof = N * rec;
    extern __shared__ __align__(16) double smem[];
    double *sx = smem;                                 // [N][Sc][3]
    double *sW = sx + ndof;                            // [M] lammda * Poisson weight of every pair
    double *sA = sW + M;                               // [2][chunk] partial ensemble sums: [source CTA][pair of my sub-slice]
    double *red = sA + 2 * ta.chunk;                   // [32] block reductions
    double *red2 = red + 32;                           // [8][5] cluster reduction of the energy terms (used on rank 0)
    double *s_moved = red2 + 40;                       // [32] per local structure: bound on the displacement since the list build
    unsigned long long *s_maxp = reinterpret_cast<unsigned long long *>(s_moved + 32);   // [32] max |p|^2 of the step (bit pattern)
    double *s_rad = reinterpret_cast<double *>(s_maxp + 32);   // [N] bead radii
    double *s_ul = s_rad + N, *s_ll = s_ul + N;        // [N] bond limits of bond (b, b+1)
    int *s_int = reinterpret_cast<int *>(s_ll + N);    // [0] bead counter, [1] number of stale structures, [2..33] stale list
    int *s_off = s_int + 40;                           // [N+1] CSR row offsets
    int *s_ord = s_off + N + 1;                        // [N] bead handled by work item q (longest rows first)
    const double nrm = ta.norm[r];
    const double lam = ta.lammda ? ta.lammda[r] : 1.0, bet = ta.beta ? ta.beta[r] : 1.0;
    const double eps = ta.eps[r];
    const double alpha = dm.alpha;
    const long long rS = (long long)r * S;
    double *gx = ta.x + rS * N * 3, *gp = ta.p + rS * N * 3;
    double *pw = ta.pw + ((long long)r * C + crank) * ndof;
    const int slot = lane / LK, kl = lane % LK;
    const bool kv = kl < nloc;
    const int kq = kv ? kl : 0;                        // lanes without a structure compute on structure 0, masked out

    // ---- load the state --------------------------------------------------------------------------------------
    for (int q = tid; q < ndof; q += EHIC_TC_THREADS) {
        const int b = q / rec, rem = q - b * rec, k = rem / 3, c = rem - 3 * k;
        double xv = (c == 0) ? (double)b : 0.0, pv = 0.0;          // unused records: distinct, finite coordinates
        if (k < nloc) { const long long g = ((long long)(k0 + k) * N + b) * 3 + c; xv = gx[g]; pv = gp[g]; }
        sx[q] = xv; pw[q] = pv;
    }
    for (int q = tid; q < N; q += EHIC_TC_THREADS) {
        s_rad[q] = dm.radii[q]; s_ul[q] = dm.bb_ul[q]; s_ll[q] = dm.bb_ll[q];
        s_off[q] = (int)ld.row_off[q]; s_ord[q] = ld.order[q];
    }
    if (tid == 0) s_off[N] = (int)ld.row_off[N];
    if (tid < 32) { s_moved[tid] = 0.0; s_maxp[tid] = 0ull; s_int[2 + tid] = tid; }
    if (tid == 0) { s_int[0] = 0; s_int[1] = nloc; }                // first step: every list is built
    __syncthreads();

    // ---- Verlet lists of the stale structures: (structure, bead) items over the whole CTA ----------------------
    auto rebuild = [&]() {
        const int n_stale = s_int[1];
        const double skin = ta.skin;
        for (int item = tid; item < n_stale * N; item += EHIC_TC_THREADS) {
            const int k = s_int[2 + item / N], b = item % N;
            const double *xb = sx + (b * Sc + k) * 3;
            const double xb0 = xb[0], xb1 = xb[1], xb2 = xb[2], rb = s_rad[b];
            unsigned short *row = ta.vl_idx + ((rS + k0 + k) * Np + b) * (long long)EHIC_VL_CAP;
            int n = 0;
            for (int o = 0; o < N; o++) {
                const double *xo = sx + (o * Sc + k) * 3;
                const double dx = xo[0] - xb0, dy = xo[1] - xb1, dz = xo[2] - xb2;
                const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const double lim = rb + s_rad[o] + skin;
                if (!(s2 >= lim * lim) && o != b) {              // !(>=): NaN distances are listed
                    if (n < EHIC_VL_CAP) row[n] = (unsigned short)o;
                    n++;
                }
            }
            ta.vl_cnt[(rS + k0 + k) * Np + b] = n;               // > EHIC_VL_CAP: this bead sweeps all partners
        }
        __syncthreads();
        if (tid < n_stale) s_moved[s_int[2 + tid]] = 0.0;
        __syncthreads();
        if (tid == 0) s_int[1] = 0;
    };
    rebuild();

    for (int step = 0; step <= ta.n_steps; step++) {
        const bool first = step == 0, last = step == ta.n_steps, ends = first || last;
        // ======== forward: ensemble sums across the cluster, Poisson weights into every CTA's table =================
        double logl = 0.0;
        const int n_chunks = (M + ta.chunk - 1) / ta.chunk;
        const int sub = ta.chunk / C;
        for (int q = 0; q < n_chunks; q++) {
            const int m_base = q * ta.chunk, m_end = min(m_base + ta.chunk, M);
            double *bufA = sA + (q & 1) * ta.chunk;
            // blocks of 32 consecutive pairs: ONE coalesced load brings the block's (i, j, dc) records, one per lane
            // (the next block's is requested before this one is worked on); the records are then handed to the LK
            // lanes of a pair by shuffles, LK steps of PW pairs each; the per-lane terms of the LK steps are reduced
            // over the structures with one halving butterfly, after which every lane holds the sum of ONE pair
            const int n_blocks = (m_end - m_base + 31) / 32;
            int4 nxt = make_int4(0, 0, 0, 0);
            if (warp < n_blocks) nxt = __ldg(reinterpret_cast<const int4 *>(ld.pairs + min(m_base + 32 * warp + lane, m_end - 1)));
            for (int blk = warp; blk < n_blocks; blk += NW) {
                const int4 cur = nxt;
                const int m_b = m_base + 32 * blk;
                if (blk + NW < n_blocks) nxt = __ldg(reinterpret_cast<const int4 *>(ld.pairs + min(m_b + 32 * NW + lane, m_end - 1)));
                constexpr int NV = LK <= 16 ? LK : 1;          // LK = 32: the terms are reduced one step at a time
                double v[NV];
                double tot = 0.0;
#pragma unroll
                for (int t = 0; t < LK; t++) {
                    const int src = t * PW + slot;
                    const int pi = __shfl_sync(0xffffffffu, cur.x, src), pj = __shfl_sync(0xffffffffu, cur.y, src);
                    const double dc = __hiloint2double(__shfl_sync(0xffffffffu, cur.w, src), __shfl_sync(0xffffffffu, cur.z, src));
                    const double *ai = sx + (pi * Sc + kq) * 3, *aj = sx + (pj * Sc + kq) * 3;
                    const double dx = ai[0] - aj[0], dy = ai[1] - aj[1], dz = ai[2] - aj[2];
                    const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double rd = rsqrt_nr(s2);
                    const double tt = alpha * (dc - s2 * rd);
                    const double rq = rsqrt_nr(fma(tt, tt, 1.0));
                    const double term = kv ? fma(tt, rq, 1.0) : 0.0;
                    if constexpr (LK <= 16) v[t] = term;
                    else { const double sum = group_sum<LK>(term); tot = (t == kl) ? sum : tot; }
                }
                if constexpr (LK <= 16) { halving_reduce<NV>(v, lane); tot = v[0]; }
                const int m = m_b + kl * PW + slot;           // the pair whose ensemble sum (over this CTA's structures) this lane holds
                if (m < m_end) {
                    const int rel = m - m_base, owner = rel / sub;
                    cluster.map_shared_rank(bufA, owner)[crank * sub + (rel - owner * sub)] = tot;
                }
            }
            cluster.sync();                              // the partials of chunk q have arrived at their owners
            // owner: add the C partials in rank order, Poisson weight -> W of every CTA
            const int own_base = m_base + crank * sub;
            for (int idx = tid; idx < sub; idx += EHIC_TC_THREADS) {
                const int m = own_base + idx;
                if (m < m_end) {
                    double sum = 0.0;
                    for (int c = 0; c < C; c++) sum += bufA[c * sub + idx];
                    const double md = sum * (0.5 * nrm);
                    const double cnt = dm.pcnt[m];
                    const double em = 1.0 - cnt / md;
                    const double w = lam * (((0.5 * em) * nrm) * alpha);
                    for (int c = 0; c < C; c++) cluster.map_shared_rank(sW, c)[m] = w;
                    if (ends) logl += cnt * log(md) - md;
                }
            }
        }
        cluster.sync();                                  // W complete everywhere

        // ======== gradient of every (structure, bead) of this CTA + kick ===============================================
        double e_nb = 0.0, e_bb = 0.0, e_sp = 0.0, kin = 0.0, pmax2 = 0.0;
        const double kick = (ends ? 0.5 : 1.0) * eps;
        const double sc_nb = bet * (dm.k_nb * 2.0);
        // beads are dealt to the warps round robin in longest-row-first order: a STATIC assignment, so that the per-thread
        // energy partials (and with them H) are summed in the same order in every run -- a shared work counter balanced
        // the warps slightly better but made the last bits of H depend on the scheduling
        for (int it = warp; it < N; it += NW) {
            const int b = s_ord[it];
            const double *xbp = sx + (b * Sc + kq) * 3;
            const double xb0 = xbp[0], xb1 = xbp[1], xb2 = xbp[2];
            const double rb = s_rad[b];
            // Verlet list of (structure kl, bead b): requested now, consumed after the contact row
            const long long vat = (rS + k0 + kq) * Np + b;
            const int vcnt = kv ? ta.vl_cnt[vat] : 0;
            const unsigned short *vrow = ta.vl_idx + vat * (long long)EHIC_VL_CAP;
            const uint4 vfirst = *reinterpret_cast<const uint4 *>(vrow + 8 * slot);      // entries [8 slot, 8 slot + 8)
            double g0 = 0.0, g1 = 0.0, g2 = 0.0;
            {   // contact row of bead b (both directions of every pair are in the CSR view); weights carry lammda.
                // Blocks of 32 entries: one coalesced load (entry + its weight per lane), next block requested ahead,
                // entries handed to the LK lanes of a row slot by shuffles, two steps in flight.
                const int e0 = s_off[b], e1 = s_off[b + 1];
                const int n_blocks = (e1 - e0 + 31) / 32;
                int4 nxt = make_int4(0, 0, 0, 0);
                if (n_blocks > 0) nxt = __ldg(reinterpret_cast<const int4 *>(ld.ent + min(e0 + lane, e1 - 1)));
                for (int blk = 0; blk < n_blocks; blk++) {
                    const int4 cur = nxt;
                    const int e_b = e0 + 32 * blk;
                    if (blk + 1 < n_blocks) nxt = __ldg(reinterpret_cast<const int4 *>(ld.ent + min(e_b + 32 + lane, e1 - 1)));
                    const double wl = (e_b + lane < e1) ? sW[cur.w] : 0.0;      // entries past the end of the row: weight 0
                    const int n_steps = min(LK, (e1 - e_b + PW - 1) / PW);
                    for (int t = 0; t < n_steps; t += 2) {
                        const int sa = t * PW + slot, sb = min(sa + PW, 31);
                        const bool vb = t + 1 < n_steps;
                        const int ja = __shfl_sync(0xffffffffu, cur.z, sa), jb = __shfl_sync(0xffffffffu, cur.z, sb);
                        const double dca = __hiloint2double(__shfl_sync(0xffffffffu, cur.y, sa), __shfl_sync(0xffffffffu, cur.x, sa));
                        const double dcb = __hiloint2double(__shfl_sync(0xffffffffu, cur.y, sb), __shfl_sync(0xffffffffu, cur.x, sb));
                        const double wa = __shfl_sync(0xffffffffu, wl, sa);
                        double wb = __shfl_sync(0xffffffffu, wl, sb);
                        wb = vb ? wb : 0.0;
                        const double *xa = sx + (ja * Sc + kq) * 3, *xc = sx + (jb * Sc + kq) * 3;
                        const double adx = xa[0] - xb0, ady = xa[1] - xb1, adz = xa[2] - xb2;
                        const double bdx = xc[0] - xb0, bdy = xc[1] - xb1, bdz = xc[2] - xb2;
                        const double as2 = fma(adz, adz, fma(ady, ady, adx * adx));
                        const double bs2 = fma(bdz, bdz, fma(bdy, bdy, bdx * bdx));
                        const double ard = rsqrt_nr(as2), brd = rsqrt_nr(bs2);
                        const double att = alpha * (dca - as2 * ard), btt = alpha * (dcb - bs2 * brd);
                        const double arq = rsqrt_nr(fma(att, att, 1.0)), brq = rsqrt_nr(fma(btt, btt, 1.0));
                        const double fa = wa * ((arq * arq) * (arq * ard)), fb = wb * ((brq * brq) * (brq * brd));
                        g0 = fma(fa, adx, g0); g1 = fma(fa, ady, g1); g2 = fma(fa, adz, g2);
                        g0 = fma(fb, bdx, g0); g1 = fma(fb, bdy, g1); g2 = fma(fb, bdz, g2);
                    }
                }
            }
            if (kv) {   // nonbonded: slot s takes list entries [8 s, 8 s + 8), then (long lists) 8 PW + s, 8 PW + s + PW, ...
                const bool all = vcnt > EHIC_VL_CAP;          // overflow: sweep every partner
                const int n = all ? N : vcnt;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                auto visit = [&](int o) {
                    const double *xo = sx + (o * Sc + kl) * 3;
                    const double dx = xo[0] - xb0, dy = xo[1] - xb1, dz = xo[2] - xb2;
                    const double s2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double d0 = rb + s_rad[o];
                    if (s2 < d0 * d0 && o != b) {
                        const double rd = rsqrt_nr(s2);
                        const double a = d0 - s2 * rd, aa = a * a;
                        const double g = aa * a * rd;
                        a0 = fma(g, dx, a0); a1 = fma(g, dy, a1); a2 = fma(g, dz, a2);
                        e_nb += aa * aa;
                    }
                };
                if (all) {
                    for (int o = slot; o < N; o += PW) visit(o);
                } else {
                    const unsigned ws[4] = {vfirst.x, vfirst.y, vfirst.z, vfirst.w};
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (8 * slot + u < n) visit((int)((ws[u >> 1] >> (16 * (u & 1))) & 0xffffu));
                    for (int s = 8 * PW + slot; s < n; s += PW) visit((int)vrow[s]);
                }
                g0 = fma(sc_nb, a0, g0); g1 = fma(sc_nb, a1, g1); g2 = fma(sc_nb, a2, g2);
            }
            __syncwarp();
#pragma unroll
            for (int o = LK; o < 32; o <<= 1) {              // combine the PW partial forces of every structure
                g0 += __shfl_xor_sync(0xffffffffu, g0, o);
                g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                g2 += __shfl_xor_sync(0xffffffffu, g2, o);
            }
            if (slot != 0 || !kv) continue;                  // the first LK lanes finish the bead, one structure each
            {   // backbone: bonds (b-1, b) and (b, b+1)  (backbone_prior_c.pyx:26-42)
                double r0 = 0.0, r1 = 0.0, r2 = 0.0;
                if (b > 0 && !isinf(s_ul[b - 1])) {
                    const double ul = s_ul[b - 1], ll = s_ll[b - 1];
                    const double *xo = sx + ((b - 1) * Sc + kl) * 3;
                    const double d0 = xb0 - xo[0], d1 = xb1 - xo[1], d2 = xb2 - xo[2];
                    const double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                    if (s2 > ul * ul) { const double d = sqrt(s2), a = d - ul; r0 = a * d0 / d; r1 = a * d1 / d; r2 = a * d2 / d; e_bb += a * a; }
                    else if (s2 < ll * ll) { const double d = sqrt(s2), a = ll - d; r0 = -(a * d0 / d); r1 = -(a * d1 / d); r2 = -(a * d2 / d); e_bb += a * a; }
                }
                if (!isinf(s_ul[b])) {
                    const double ul = s_ul[b], ll = s_ll[b];
                    const double *xo = sx + ((b + 1) * Sc + kl) * 3;
                    const double d0 = xo[0] - xb0, d1 = xo[1] - xb1, d2 = xo[2] - xb2;
                    const double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                    if (s2 > ul * ul) { const double d = sqrt(s2), a = d - ul; r0 -= a * d0 / d; r1 -= a * d1 / d; r2 -= a * d2 / d; }
                    else if (s2 < ll * ll) { const double d = sqrt(s2), a = ll - d; r0 += a * d0 / d; r1 += a * d1 / d; r2 += a * d2 / d; }
                }
                g0 = fma(dm.k_bb, r0, g0); g1 = fma(dm.k_bb, r1, g1); g2 = fma(dm.k_bb, r2, g2);
            }
            if (dm.sphere_active) {   // sphere_prior_c.pyx:25-36
                const double R0 = dm.sph_R;
                const double d0 = xb0 - dm.sph_cx, d1 = xb1 - dm.sph_cy, d2 = xb2 - dm.sph_cz;
                const double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                if (!(s2 < (R0 * R0 + rb * rb) - (2.0 * R0) * rb)) {
                    const double d = sqrt(s2), a = (d + rb) - R0, f = a / d;
                    g0 = fma(dm.sph_k, d0 * f, g0); g1 = fma(dm.sph_k, d1 * f, g1); g2 = fma(dm.sph_k, d2 * f, g2);
                    if (a > 0.0) e_sp += a * a;
                }
            }
            double *pp = pw + (b * Sc + kl) * 3;
            double p0 = pp[0], p1 = pp[1], p2 = pp[2];
            if (first) kin += p0 * p0 + p1 * p1 + p2 * p2;
            p0 = fma(-kick, g0, p0); p1 = fma(-kick, g1, p1); p2 = fma(-kick, g2, p2);
            pp[0] = p0; pp[1] = p1; pp[2] = p2;
            const double pp2 = p0 * p0 + p1 * p1 + p2 * p2;
            if (last) kin += pp2;
            pmax2 = fmax(pmax2, (pp2 == pp2) ? pp2 : 1e300);   // NaN momenta: rebuild (the sweep then sees them)
        }
        if (slot == 0 && kv) atomicMax(&s_maxp[kl], (unsigned long long)__double_as_longlong(pmax2));   // max: order-free

        if (ends) {   // Hamiltonian: block sums, then across the cluster through rank 0's shared memory
            const double tl = block_sum512(logl, red), tn = block_sum512(e_nb, red), tb = block_sum512(e_bb, red),
                         ts = block_sum512(e_sp, red), tk = block_sum512(kin, red);
            if (tid == 0) {
                double *dst = cluster.map_shared_rank(red2, 0) + crank * 5;
                dst[0] = tl; dst[1] = tn; dst[2] = tb; dst[3] = ts; dst[4] = tk;
            }
            cluster.sync();
            if (crank == 0 && tid == 0) {
                double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                for (int c = 0; c < C; c++)
#pragma unroll
                    for (int u = 0; u < 5; u++) a[u] += red2[c * 5 + u];
                double *H = ta.H + (long long)r * 4 + (first ? 0 : 2);
                H[0] = lam * (-a[0]) + bet * (0.25 * dm.k_nb * a[1]) + 0.5 * dm.k_bb * a[2] + 0.5 * dm.sph_k * a[3];
                H[1] = 0.5 * a[4];
            }
        }
        __syncthreads();                                  // every momentum of this CTA is kicked, nobody reads sx any more
        if (!last) {
            if (tid < nloc) {                             // displacement bound of the coming drift: eps * max |p|
                const double moved = s_moved[tid] + eps * sqrt(__longlong_as_double((long long)s_maxp[tid]));
                s_moved[tid] = moved;
                s_maxp[tid] = 0ull;
                if (!(moved <= 0.5 * ta.skin)) s_int[2 + atomicAdd(&s_int[1], 1)] = tid;    // integer slots: order-free result
            }
            for (int q = tid; q < ndof; q += EHIC_TC_THREADS) sx[q] = fma(eps, pw[q], sx[q]);
            __syncthreads();
            if (s_int[1] > 0) rebuild();                  // CTA-uniform
        }
    }

    // ---- store the end point -----------------------------------------------------------------------------------
    for (int q = tid; q < ndof; q += EHIC_TC_THREADS) {
        const int b = q / rec, rem = q - b * rec, k = rem / 3, c = rem - 3 * k;
        if (k < nloc) { const long long g = ((long long)(k0 + k) * N + b) * 3 + c; gx[g] = sx[q]; gp[g] = pw[q]; }
    }
    if (tid < nloc && ta.vl_state) ta.vl_state[rS + k0 + tid] = 0;
    cluster.sync();                                       // no CTA exits while a peer may still address its shared memory
}

// dynamic shared memory of one CTA
static inline size_t traj_cluster_smem(int N, int Sc, long long M, int chunk)
{
    return sizeof(double) * ((size_t)N * Sc * 3 + (size_t)M + 2 * (size_t)chunk + 32 + 40 + 32 + 3 * (size_t)N) +
           sizeof(unsigned long long) * 32 + sizeof(int) * (40 + 2 * (size_t)N + 2);
}

}  // namespace ehic
