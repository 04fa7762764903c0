// ehic_b200.cu -- C ABI (include/ehic_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is deliberately no CPU implementation behind these entry points: without a
// compute-capability-10 device every compute call fails with EHIC_ERR_NODEVICE.
#include "../../include/ehic_b200.h"
#include "kernels.cuh"
#include "traj_cluster.cuh"
#include "layout.hpp"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

using namespace ehic;

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string &msg) { g_err = msg; return code; }

struct ehic_model;
struct KernelTimer {            // RAII: records a start event now and a stop event at scope exit
    ehic_model *m; cudaStream_t st; int idx;
    KernelTimer(ehic_model *m_, int cls, cudaStream_t st_);
    ~KernelTimer();
};

#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(e_ == cudaErrorMemoryAllocation ? EHIC_ERR_NOMEM : EHIC_ERR_CUDA,  \
                        std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

enum { KC_PACK = 0, KC_FORWARD, KC_PRIOR, KC_GATHER, KC_FINALIZE, KC_OTHER, KC_COUNT };

struct ehic_model {
    int device = 0;
    int flags = 0;
    int max_replicas = 0;
    DevModel dm{};
    ContactLayout layout;           // host copy (sizes, perm)
    bool traj_small_ready = false;
    bool use_tiles = false;         // dense list + fast mode: symmetric tile path
    bool nb_fused = false;          // tiles cover all bead pairs: nonbonded force rides in the backward tile kernel
    const unsigned char *tile_clean = nullptr;
    TileDev td{};
    bool f32 = false;               // EHIC_FLAG_FP32: FP32 pair arithmetic in the tile kernels (dense lists) / lane kernels (all other lists)
    const float *dc32 = nullptr;    // [n_tiles*4096] contact distances as floats
    float *tw32 = nullptr;          // [R][n_tiles*4096] lammda * weight as floats
    bool use_lanes = false;         // sparse list + fast mode + enough structures to fill warps: lane path
    LaneDev ld{};
    int n_quads = 0;                // lane path: backward CTAs per (replica, structure group)
    int n_quads_split = 0;          // the same with long rows cut into segments (few replicas per GPU)
    const int4 *lane_items_split = nullptr;
    bool in_split = false;          // enqueueing one of two concurrent half batches: no replica-resident launches (they fill every SM)
    int last_lane_quads = 0;        // quads per (replica, structure group) of the last backward launch (kinetic partials)
    int lane_split_mode = -1;       // EHIC_LANE_SPLIT: 0 / 1 forced, -1 by launch size
    std::vector<long long> lane_item_work;  // host: CSR entries (+ epilogue) per backward work item
    const long long *lane_cum = nullptr;   // [nkg * n_quads + 1] cumulative CSR entries of a replica's backward virtual blocks
    bool lane_res_fwd = false, lane_res_bwd = false;   // replica-resident launches (a replica's lane copy fits one SM's L1)
    int n_sm = 0;
    double tile_fill = 0.0;
    int nA_tile = 0, nKc = 0;       // forward-tile blocks per replica; combine blocks per structure
    double *tw = nullptr, *Grow = nullptr, *Pbuf = nullptr;
    std::vector<void *> owned;      // device allocations
    // workspace
    double *xs[2] = {nullptr, nullptr};   // SoA working copies (forward + prior kernels)
    double *xt[2] = {nullptr, nullptr};   // tile-AoS working copies (gather kernel)
    double *wbuf = nullptr, *gprior = nullptr, *bbox = nullptr, *partC = nullptr;
    int *cell_list = nullptr;
    bool use_cells = false;         // fast mode: nonbonded term through the per-structure cell grid
    bool in_host_pipeline = false;  // ehic_evaluate_host is enqueueing chunks on the two compute streams
    // tuning knobs from the environment, read ONCE when the model is created (never on the hot path):
    int opt_split = -1;             // EHIC_SPLIT: parts a large device batch is evaluated in (-1 = automatic)
    bool opt_traj_multi = false;    // EHIC_TRAJ=multi: never use the one-launch trajectory kernels
 bool opt_traj_no_small = false; // EHIC_TRAJ=cluster: skip the one-CTA kernel (prefer the cluster kernel)
    int opt_cluster = -1;           // EHIC_TRAJ_CLUSTER: 2/4/8 = use the cluster kernel with this cluster size; with EHIC_TRAJ=cluster and no size: automatic size
    // cluster trajectory kernel (traj_cluster.cuh)
    bool tc_possible = false;       // static eligibility: lane layout uploaded, Verlet buffers present
    double *tc_pw = nullptr;        // [max_replicas][(S + 7) * N * 3] momenta in kernel order
    struct TcPlan { bool ok = false; int R = 0, C = 0, Sc = 0, LK = 0, chunk = 0, conc = 0; size_t smem = 0; };
    std::vector<TcPlan> tc_plans;   // per batch size R
    bool tc_attr[4] = {false, false, false, false};
    bool use_verlet = false;        // HMC trajectories: nonbonded term from Verlet lists reused across leapfrog steps
    bool vl_fused = true;           // one verlet_step_kernel per leapfrog step (N <= 512); else check / build / sweep / force kernels
    VerletDev vl{};
    double *partA = nullptr, *partP = nullptr, *partK = nullptr;
    double *energies = nullptr, *kinetic = nullptr;
    int nA = 0;                     // forward blocks per replica
    int nPx = 0;                    // prior blocks per structure
    int prior_threads = 256;
    bool last_prior_used_cells = false;
    int cur_xt = 0;                 // which xt buffer matches the xs buffer handed to launch_forward
    int kt = 1, nkt = 1, nktg = 1;  // structures per warp, tiles per replica, tile groups (gw warps) per slice
    int gw = 4;                     // warps per gather CTA
    double r_max = 0.0;
    // CUDA graphs of whole HMC trajectories, keyed by the call signature; dropped when a parameter changes
    struct TrajGraph { std::vector<const void *> key; cudaGraphExec_t exec; long long n_kernels; };
    std::vector<TrajGraph> graphs;
    // optional per-kernel timing (ehic_profile): events bracket every launch on its stream
    bool profiling = false;
    struct Timed { int cls; cudaEvent_t a, b; };
    std::vector<Timed> timed;
    // staging for the host entry points
    double *st_x = nullptr, *st_g = nullptr, *st_md = nullptr, *st_small = nullptr;
    cudaStream_t st_stream = nullptr, st_stream2 = nullptr, st_in = nullptr, st_out = nullptr;
    static constexpr int kMaxChunks = 8;
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_done[kMaxChunks] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // trajectory: prior kernels run beside the forward kernel
    // single evaluations: compute lane 0 = the caller's stream / st_stream, lane 1 = st_stream2; each lane has its
    // own side stream for the prior kernels
    cudaStream_t st_side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork_l[2] = {nullptr, nullptr}, ev_join_l[2] = {nullptr, nullptr}, ev_half = nullptr, ev_half_done = nullptr;
    // per-replica strides (elements) of the workspace arrays, for evaluating a chunk in replica slots [s, s+n)
    struct WsStrides { size_t xs = 0, xt = 0, wbuf = 0, tw = 0, gprior = 0, bbox = 0, partC = 0, partA = 0, partP = 0,
                       partK = 0, Grow = 0, Pbuf = 0, cell = 0; } wss;
};

// RAII: marks the model as "enqueueing the parts of a split call" for the duration of a scope (also on the
// early-return error paths)
struct PipelineScope {
    ehic_model *m; bool prev;
    explicit PipelineScope(ehic_model *m_) : m(m_), prev(m_->in_host_pipeline) { m->in_host_pipeline = true; }
    ~PipelineScope() { m->in_host_pipeline = prev; }
};

static void drop_graphs(ehic_model *m)
{
    for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
}

KernelTimer::KernelTimer(ehic_model *m_, int cls, cudaStream_t st_) : m(m_), st(st_), idx(-1)
{
    if (!m->profiling) return;
    ehic_model::Timed t; t.cls = cls;
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
    cudaEventRecord(t.a, st);
    m->timed.push_back(t);
    idx = (int)m->timed.size() - 1;
}
KernelTimer::~KernelTimer() { if (idx >= 0) cudaEventRecord(m->timed[idx].b, st); }

static size_t forward_tile_smem(int kt)
{
    return sizeof(double) * (4096 + (size_t)EHIC_FT_STAGES * (128 + 32) * kt * 3) + 64;   // + mbarriers
}

template <typename T>
static int upload(ehic_model *m, const std::vector<T> &v, const T **out)
{
    T *d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    m->owned.push_back(d);
    if (!v.empty()) CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return EHIC_OK;
}

static int dalloc(ehic_model *m, double **out, size_t n)
{
    double *d = nullptr;
    CU(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(double)));
    m->owned.push_back(d);
    *out = d;
    return EHIC_OK;
}

extern "C" {

int ehic_version(void) { return 100; }
const char *ehic_last_error(void) { return g_err.c_str(); }
int64_t ehic_launch_count(void) { return g_launches.load(); }

int ehic_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int check_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(EHIC_ERR_NODEVICE, "no CUDA device visible: libehic_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(EHIC_ERR_INVALID, "device index out of range");
    int major = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return fail(EHIC_ERR_NODEVICE, "device is not compute capability 10.x (kernels are built for sm_100a only)");
    return EHIC_OK;
}

// warps per gather CTA: all warps of a CTA share one slice's row stream, so more warps mean
// less L2 traffic; pick the count that wastes the fewest warp slots on the last tile group.
static int pick_gw(int nkt)
{
    const char *env = getenv("EHIC_GW");
    if (env) { int v = atoi(env); if (v >= 1 && v <= EHIC_GATHER_MAX_WARPS) return v; }
    if (nkt <= 4) return nkt < 1 ? 1 : nkt;
    int best = 4; double best_eff = 0.0;
    for (int w = 4; w <= 8; w++) {
        double eff = (double)nkt / (double)(((nkt + w - 1) / w) * w);
        if (eff >= best_eff) { best_eff = eff; best = w; }
    }
    return best;
}

static int pick_kt(int S)
{
    const char *env = getenv("EHIC_KT");
    if (env) { int v = atoi(env); if (v == 1 || v == 2 || v == 4) return v; }
    if (S >= 4) return 4;
    if (S >= 2) return 2;
    return 1;
}

int ehic_model_create(const ehic_model_desc_t *desc, int device, ehic_model_t **out)
{
    if (!desc || !out) return fail(EHIC_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->n_beads <= 0 || desc->n_structures <= 0) return fail(EHIC_ERR_INVALID, "n_beads and n_structures must be positive");
    if (desc->n_data < 0 || (desc->n_data > 0 && (!desc->data_points || !desc->contact_distances)))
        return fail(EHIC_ERR_INVALID, "data_points / contact_distances missing");
    if (!desc->bead_radii) return fail(EHIC_ERR_INVALID, "bead_radii missing");
    if (desc->max_replicas <= 0) return fail(EHIC_ERR_INVALID, "max_replicas must be positive");
    if ((desc->flags & EHIC_FLAG_FP32) && (desc->flags & EHIC_FLAG_EXACT))
        return fail(EHIC_ERR_INVALID, "EHIC_FLAG_FP32 and EHIC_FLAG_EXACT exclude each other");
    int rc = check_device(device);
    if (rc) return rc;
    CU(cudaSetDevice(device));

    ehic_model *m = new (std::nothrow) ehic_model();
    if (!m) return fail(EHIC_ERR_NOMEM, "host allocation failed");
    m->device = device; m->flags = desc->flags; m->max_replicas = desc->max_replicas;
    { const char *e = getenv("EHIC_SPLIT"); m->opt_split = e ? atoi(e) : -1; }
    { const char *e = getenv("EHIC_TRAJ"); m->opt_traj_multi = e && !strcmp(e, "multi"); m->opt_traj_no_small = e && !strcmp(e, "cluster"); }
    { const char *e = getenv("EHIC_TRAJ_CLUSTER"); m->opt_cluster = e ? atoi(e) : -1; }
    const int N = desc->n_beads, S = desc->n_structures;
    const bool exact = (desc->flags & EHIC_FLAG_EXACT) != 0;
    std::string err = build_layout(N, desc->n_data, desc->data_points, desc->contact_distances, exact, m->layout);
    if (!err.empty()) { delete m; return fail(EHIC_ERR_INVALID, err); }
    ContactLayout &L = m->layout;

    // bonds: [N] arrays indexed by bond (b, b+1)
    std::vector<double> ul(N, std::numeric_limits<double>::infinity()), ll(N, 0.0);
    std::vector<char> is_end(N, 0);
    if (desc->mol_ranges && desc->n_mol_ranges >= 2) {
        if (desc->mol_ranges[0] != 0 || desc->mol_ranges[desc->n_mol_ranges - 1] != N) {
            delete m; return fail(EHIC_ERR_INVALID, "mol_ranges must start at 0 and end at n_beads");
        }
        for (int q = 1; q < desc->n_mol_ranges; q++) {
            if (desc->mol_ranges[q] <= desc->mol_ranges[q - 1]) { delete m; return fail(EHIC_ERR_INVALID, "mol_ranges must be increasing"); }
            is_end[desc->mol_ranges[q] - 1] = 1;
        }
    } else {
        is_end[N - 1] = 1;
    }
    for (int b = 0; b + 1 < N; b++) {
        if (is_end[b]) continue;
        ul[b] = desc->bb_upper ? desc->bb_upper[b] : desc->bead_radii[b] + desc->bead_radii[b + 1];
        ll[b] = desc->bb_lower ? desc->bb_lower[b] : 0.0;
    }
    std::vector<double> radii(desc->bead_radii, desc->bead_radii + N);

    DevModel &dm = m->dm;
    dm.N = N; dm.S = S; dm.Np = (N + 31) / 32 * 32; dm.n_slices = L.n_slices;
    dm.M = L.M; dm.tot = L.tot; dm.wstride = L.tot + ContactLayout::kTailPad;
    static_assert(ContactLayout::kTailPad == EHIC_CHUNK * 32, "tail pad must cover one pipeline chunk");
    for (int b = 0; b < N; b++) m->r_max = std::max(m->r_max, desc->bead_radii[b]);
    dm.alpha = desc->alpha; dm.k_nb = desc->nonbonded_k; dm.k_bb = desc->backbone_k;
    dm.sphere_active = desc->sphere_active;
    dm.sph_R = desc->sphere_radius; dm.sph_k = desc->sphere_k;
    dm.sph_cx = desc->sphere_center[0]; dm.sph_cy = desc->sphere_center[1]; dm.sph_cz = desc->sphere_center[2];

    // dense lists in fast mode take the symmetric tile path (EHIC_PATH=rows|tiles overrides)
    TileLayout TL;
    {
        const char *path = getenv("EHIC_PATH");
        const char *mf = getenv("EHIC_TILE_MIN_FILL");
        double min_fill = mf ? atof(mf) : 0.5;
        bool want = !exact && !(path && (!strcmp(path, "rows") || !strcmp(path, "lanes")));
        if (path && !strcmp(path, "tiles")) min_fill = 0.0;
        if (want && desc->n_data > 0) build_tiles(L, desc->data_points, desc->contact_distances, min_fill, TL);
        m->use_tiles = TL.eligible;
        m->tile_fill = TL.fill;
        // FP32 mode: dense lists run the FP32 tile kernels, every other list the FP32 lane kernels (below)
        if ((desc->flags & EHIC_FLAG_FP32) && desc->n_data > 0) {
            if (!TL.eligible && path && !strcmp(path, "rows")) {
                delete m;
                return fail(EHIC_ERR_INVALID, "EHIC_FLAG_FP32 has no row-path kernels: unset EHIC_PATH=rows (dense lists take the tile "
                                              "path, all others the lane path)");
            }
            m->f32 = true;
        }
    }
    std::vector<long long> perm(L.perm.begin(), L.perm.end()), soff(L.slice_off.begin(), L.slice_off.end());
#define UP(vec, field)                                              \
    do { rc = upload(m, vec, &dm.field); if (rc) { ehic_model_destroy(m); return rc; } } while (0)
    UP(L.pi, pi); UP(L.pj, pj); UP(L.slot_a, slot_a); UP(L.slot_b, slot_b);
    UP(L.pdc, pdc); UP(L.pcnt, pcnt); UP(perm, perm);
    UP(soff, slice_off); UP(L.slice_w, slice_w); UP(L.ent_j, ent_j); UP(L.ent_dc, ent_dc); UP(L.row_bead, row_bead);
    UP(radii, radii); UP(ul, bb_ul); UP(ll, bb_ll);
#undef UP
    if (m->use_tiles) {
        TileDev &td = m->td;
        td.n_rb = TL.n_rb; td.n_tiles = TL.n_tiles; td.n_sub = (int)TL.sub_tile.size(); td.n_sl = L.n_slices;
        td.slots = (long long)TL.n_tiles * 4096;
        std::vector<long long> orig(TL.orig.begin(), TL.orig.end());
#define UPT(vec, field)                                             \
    do { rc = upload(m, vec, &td.field); if (rc) { ehic_model_destroy(m); return rc; } } while (0)
        UPT(TL.tile_I, tile_I); UPT(TL.tile_B, tile_B); UPT(TL.rb_begin, rb_begin);
        UPT(TL.sub_tile, sub_tile); UPT(TL.sub_rb, sub_rb); UPT(TL.exists, exists);
        UPT(TL.mask, mask); UPT(TL.dc, dc); UPT(TL.cnt, cnt); UPT(orig, orig);
#undef UPT
        rc = upload(m, TL.clean, &m->tile_clean); if (rc) { ehic_model_destroy(m); return rc; }
        if (m->f32) {
            std::vector<float> dcf(TL.dc.begin(), TL.dc.end());
            rc = upload(m, dcf, &m->dc32); if (rc) { ehic_model_destroy(m); return rc; }
        }
        // fusing the nonbonded force into the backward tile kernel is off by default: with slice bounding-box
        // pruning the separate prior kernel is cheaper than the extra registers the fused variant needs
        { const char *nf = getenv("EHIC_NB_FUSED"); m->nb_fused = !m->f32 && TL.covers_all_pairs && nf && !strcmp(nf, "1"); }
    }
    // sparse lists in fast mode take the lane path (lane <-> structure) when the ensemble fills at least
    // 60 % of the lanes of its structure groups; EHIC_PATH=lanes|rows overrides
    if (!exact && desc->n_data > 0) {
        const char *path = getenv("EHIC_PATH");
        const int nkg = (S + 31) / 32;
        bool want = (double)S >= 0.6 * 32.0 * nkg;
        if (path && !strcmp(path, "lanes")) want = true;
        if (path && !strcmp(path, "rows")) want = false;
        if (m->f32) want = true;                               // the FP32 mode of sparse lists lives on the lane path
        if (m->use_tiles) want = false;
        // the cluster trajectory kernel walks the same CSR / pair views: its weight table (8 M bytes) must fit beside
        // the coordinates in shared memory, and Verlet lists index beads with 16 bits
        // Opt-in (EHIC_TRAJ=cluster or EHIC_TRAJ_CLUSTER=2/4/8): on the reference's production shape (nora2012 3 kb) the
        // captured multi-kernel graph measured faster at every batch size, so it stays the default.
        const bool tc_want = (size_t)desc->n_data * 8 <= 160 * 1024 && N >= 32 && N < EHIC_CELL_MAX && S >= 2 &&
                             (m->opt_traj_no_small || m->opt_cluster > 0);
        if (want || tc_want) {
            LaneLayout LL;
            err = build_lanes(L, LL);
            if (!err.empty()) { ehic_model_destroy(m); return fail(EHIC_ERR_INVALID, err); }
            std::vector<LanePair> pairs((size_t)L.M);
            for (int64_t s = 0; s < L.M; s++) { pairs[s].i = L.pi[s]; pairs[s].j = L.pj[s]; pairs[s].dc = L.pdc[s]; }
            std::vector<LaneEnt> ents(LL.ent_j.size());
            for (size_t e = 0; e < ents.size(); e++) { ents[e].dc = LL.ent_dc[e]; ents[e].j = LL.ent_j[e]; ents[e].pad = 0; }
            for (int64_t s = 0; s < L.M; s++) { ents[LL.slot_a[s]].pad = (int)s; ents[LL.slot_b[s]].pad = (int)s; }   // pair of the entry
            std::vector<long long> roff(LL.row_off.begin(), LL.row_off.end());
            LaneDev &ld = m->ld;
            ld.nkg = nkg;
#define UPL(vec, field)                                             \
    do { rc = upload(m, vec, &ld.field); if (rc) { ehic_model_destroy(m); return rc; } } while (0)
            // backward work items: longest rows first when the rows are ragged (the 4 warps of a CTA then walk rows of
            // similar length and the short rows fill the end of the launch); natural order otherwise, which keeps
            // the partner rows that neighbouring beads share (banded lists) close in time
            std::vector<int32_t> order(N);
            std::iota(order.begin(), order.end(), 0);
            {
                int64_t mx = 0;
                for (int b = 0; b < N; b++) mx = std::max<int64_t>(mx, L.row_len[b]);
                const char *lo = getenv("EHIC_LANE_ORDER");
                const bool ragged = (double)mx * N > 1.5 * 2.0 * (double)L.M;       // max > 1.5 x mean row length
                if (lo ? !strcmp(lo, "sorted") : ragged)
                    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return L.row_len[a] > L.row_len[b]; });
            }
            UPL(pairs, pairs); UPL(LL.slot_a, slot_a); UPL(LL.slot_b, slot_b); UPL(roff, row_off); UPL(ents, ent); UPL(order, order);
            if (m->f32 && want) {                                  // float copies of the contact distances in both views
                std::vector<LanePair32> p32(pairs.size());
                for (size_t q = 0; q < pairs.size(); q++) { p32[q].i = pairs[q].i; p32[q].j = pairs[q].j; p32[q].dc = (float)pairs[q].dc; p32[q].pad = 0; }
                std::vector<LaneEnt32> e32(ents.size());
                for (size_t q = 0; q < ents.size(); q++) { e32[q].dc = (float)ents[q].dc; e32[q].j = ents[q].j; }
                UPL(p32, pairs32); UPL(e32, ent32);
            }
#undef UPL
            m->use_lanes = want;
            m->tc_possible = tc_want;
            // backward work items: one warp each, four to a CTA; two lists.  The plain one gives every bead one warp
            // (neighbouring rows then share partner lines in L1).  In the SPLIT one rows are cut into 2 or 4 segments that
            // sit in the same CTA and are combined in shared memory: when a GPU holds few replicas of a medium system --
            // the ladder strong-scaled over 8 GPUs -- the launch is as long as the serial walk of the longest row, and
            // splitting shortens it (nora2012 backward pass, 19 replicas: 0.066 -> 0.051 ms, 38: 0.095 -> 0.088); with many
            // replicas the extra CTAs and barriers cost 12 %, so launch_gather picks the list by the size of the launch.
            // EHIC_LANE_SPLIT=0/1 forces one of them.
            {
                const char *sp = getenv("EHIC_LANE_SPLIT");
                m->lane_split_mode = sp ? (atoi(sp) != 0 ? 1 : 0) : -1;
                for (int pass = 0; pass < 2; pass++) {
                    const bool split = pass == 1;
                    std::vector<int4> items;
                    const int4 none = make_int4(-1, 0, 0, 0 | (1 << 8));
                    for (int it = 0; it < N; it++) {
                        const int b = order[it];
                        const int len = (int)L.row_len[b];
                        const int nseg = !split ? 1 : (len > 96 ? 4 : (len > 48 ? 2 : 1));
                        while ((int)(items.size() % EHIC_LANE_WARPS) + nseg > EHIC_LANE_WARPS) {
                            int4 e = none; e.w = (int)(items.size() % EHIC_LANE_WARPS) | (1 << 8); items.push_back(e);
                        }
                        const int lead = (int)(items.size() % EHIC_LANE_WARPS);
                        for (int sgm = 0; sgm < nseg; sgm++)
                            items.push_back(make_int4(b, (int)(roff[b] + (long long)len * sgm / nseg),
                                                      (int)(roff[b] + (long long)len * (sgm + 1) / nseg), lead | (nseg << 8)));
                    }
                    while (items.size() % EHIC_LANE_WARPS) { int4 e = none; e.w = (int)(items.size() % EHIC_LANE_WARPS) | (1 << 8); items.push_back(e); }
                    if (!split) {
                        m->n_quads = (int)(items.size() / EHIC_LANE_WARPS);
                        rc = upload(m, items, &m->ld.items); if (rc) { ehic_model_destroy(m); return rc; }
                        m->lane_item_work.assign(items.size(), 0);
                        for (size_t q = 0; q < items.size(); q++) m->lane_item_work[q] = (items[q].z - items[q].y) + (items[q].x >= 0 ? 8 : 0);
                    } else {
                        m->n_quads_split = (int)(items.size() / EHIC_LANE_WARPS);
                        rc = upload(m, items, &m->lane_items_split); if (rc) { ehic_model_destroy(m); return rc; }
                    }
                }
            }
            if (want) {
                // replica-resident launches (kernels.cuh, lane_resident_range): chosen when the lane copy a pass walks --
                // one structure group in the backward pass, all of them in the forward pass -- fits one SM's L1
                // (EHIC_LANE_RESIDENT=0 / 1 / 2: off / forward / forward and backward; EHIC_LANE_L1_KB: the capacity assumed, default 248)
                std::vector<long long> cum((size_t)nkg * m->n_quads + 1, 0);
                for (int kgq = 0; kgq < nkg; kgq++)
                    for (int q = 0; q < m->n_quads; q++) {
                        long long w = 0;
                        for (int u = 0; u < EHIC_LANE_WARPS; u++) w += m->lane_item_work[(size_t)q * EHIC_LANE_WARPS + u];   // entries + the bead's epilogue
                        cum[(size_t)kgq * m->n_quads + q + 1] = cum[(size_t)kgq * m->n_quads + q] + w;
                    }
                rc = upload(m, cum, &m->lane_cum); if (rc) { ehic_model_destroy(m); return rc; }
                cudaDeviceGetAttribute(&m->n_sm, cudaDevAttrMultiProcessorCount, device);
                const char *lr = getenv("EHIC_LANE_RESIDENT"), *lk = getenv("EHIC_LANE_L1_KB");
                const size_t cap = (size_t)(lk ? atoi(lk) : 248) * 1024, one = (size_t)dm.Np * 768;
                // measured on nora2012 (302 replicas, one B200): forward 0.350 -> 0.314 ms per step (93 % L1 hits, L2 sectors / 4);
                // the backward pass gets 87 % L1 hits too but runs 0.556 -> 0.699 ms (8 warps per scheduler instead of 9, static
                // ranges instead of the hardware's dynamic CTA dispatch), so only the forward pass takes it by default
                m->lane_res_bwd = lr ? atoi(lr) == 2 : false;
                m->lane_res_fwd = lr ? atoi(lr) != 0 : one * nkg <= cap;
                for (auto f : {(const void *)forward_lane_resident_kernel, (const void *)forward_lane32_resident_kernel,
                               (const void *)backward_lane_resident_kernel, (const void *)backward_lane32_resident_kernel})
                    cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 0);      // all of it as L1
            }
        }
    }
    // free the big host vectors we no longer need
    std::vector<int32_t>().swap(L.ent_j); std::vector<double>().swap(L.ent_dc);
    std::vector<int32_t>().swap(L.slot_a); std::vector<int32_t>().swap(L.slot_b);
    std::vector<double>().swap(L.pdc); std::vector<double>().swap(L.pcnt);

    m->kt = m->use_lanes ? 32 : pick_kt(S);                  // lane path: xt holds xg[R][nkg][Np][3][32]
    m->nkt = (S + m->kt - 1) / m->kt;
    m->gw = pick_gw(m->nkt);
    m->nktg = (m->nkt + m->gw - 1) / m->gw;
    m->nA = (int)((L.M + 255) / 256);
    {   // prior CTA size: the one that wastes the fewest thread slots on the last block of a structure
        int best = 256; long long best_slots = 1LL << 60;
        for (int t : {256, 128, 64}) {
            long long slots = (long long)((N + t - 1) / t) * t;
            if (slots < best_slots) { best_slots = slots; best = t; }
        }
        m->prior_threads = best;
        m->nPx = (N + best - 1) / best;
    }
    const size_t Rm = (size_t)desc->max_replicas;
    const size_t xs_n = Rm * S * 3 * (size_t)dm.Np;
#define DA(ptr, n)                                                   \
    do { rc = dalloc(m, &(ptr), (n)); if (rc) { ehic_model_destroy(m); return rc; } } while (0)
    DA(m->xs[0], xs_n); DA(m->xs[1], xs_n);
    const size_t xt_n = Rm * m->nkt * (size_t)dm.Np * m->kt * 3;
    DA(m->xt[0], xt_n); DA(m->xt[1], xt_n);
    for (int q = 0; q < 2; q++)     // records of padding structures (k >= S) stay zero forever
        if (cudaMemset(m->xt[q], 0, xt_n * sizeof(double)) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "cudaMemset failed"); }
    DA(m->wbuf, Rm * (size_t)dm.wstride);
    DA(m->gprior, Rm * S * (size_t)N * 3);
    DA(m->bbox, Rm * S * (size_t)L.n_slices * 6);
    DA(m->partC, Rm * S * 3);
    {
        const char *nbm = getenv("EHIC_NB");
        m->use_cells = !exact && (N >= 4096 || (nbm && !strcmp(nbm, "cells") && N >= 32)) && !(nbm && !strcmp(nbm, "pairs"));
        if ((desc->flags & EHIC_FLAG_NB_CELLS) && !exact) m->use_cells = true;
        if (m->use_cells) {
            int *cl = nullptr;
            if (cudaMalloc(&cl, sizeof(int) * Rm * S * (size_t)N) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_NOMEM, "cudaMalloc failed"); }
            m->owned.push_back(cl); m->cell_list = cl;
        }
    }
    {   // Verlet lists (trajectories only): fast mode, all-pairs regime, bounded memory
        const char *ve = getenv("EHIC_VERLET");
        const size_t idx_bytes = sizeof(unsigned short) * Rm * S * (size_t)EHIC_VL_CAP * dm.Np;
        m->use_verlet = !exact && !m->use_cells && N >= 32 && N < EHIC_CELL_MAX && desc->nonbonded_k != 0.0 &&
                        !(ve && !strcmp(ve, "0")) &&
                        idx_bytes <= ((size_t)8 << 30);
        if (m->use_verlet) {
            // One verlet_step_kernel per leapfrog step for structures of up to 512 beads, where the CTA's own all-pairs rebuild
            // from shared memory is cheap; longer chains keep the check / build / force kernels, whose rebuild is pruned by
            // the slice boxes and spread over many CTAs (config D at step sizes that make lists go stale every few steps:
            // prior part 0.40 ms per leapfrog step with the kernel chain, 1.38 ms fused).  EHIC_VERLET_FUSED=0/1 overrides.
            { const char *vf = getenv("EHIC_VERLET_FUSED"); m->vl_fused = vf ? strcmp(vf, "0") != 0 : N <= 512; }
            // skin: the fused kernel pays ~6 normal steps for a rebuild, so it wants rarer rebuilds and accepts longer lists
            // (nora2012 at its adapted step size 2.5e-3: 1.266 ms per step at 0.25 r_max, 1.129 at 0.7, 1.117 at 1.0); the
            // kernel chain was tuned at 0.25 r_max in round 1
            const char *sk = getenv("EHIC_VERLET_SKIN");
            const double skin_default = m->vl_fused ? 0.7 : 0.25;
            m->vl.skin = (sk ? atof(sk) : skin_default) * m->r_max;
            if (!(m->vl.skin > 0.0)) m->vl.skin = skin_default * m->r_max;
            int *ip = nullptr;
            const size_t n_int = Rm * S * (size_t)(2 + dm.Np + 32) / 32 * 32 + idx_bytes / sizeof(int);
            if (cudaMalloc(&ip, sizeof(int) * n_int) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_NOMEM, "cudaMalloc failed"); }
            m->owned.push_back(ip);
            if (cudaMemset(ip, 0, sizeof(int) * Rm * S * (size_t)(2 + dm.Np)) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "cudaMemset failed"); }
            m->vl.state = ip; m->vl.stale = ip + Rm * S; m->vl.cnt = ip + 2 * Rm * S; m->vl.idx = reinterpret_cast<unsigned short *>(ip + Rm * S * (size_t)(2 + dm.Np + 32) / 32 * 32);     // 128-byte aligned
            DA(m->vl.xref, xs_n);
            {
                unsigned short *pp = nullptr;
                if (cudaMalloc(&pp, sizeof(unsigned short) * Rm * S * (size_t)dm.Np) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_NOMEM, "cudaMalloc failed"); }
                m->owned.push_back(pp); m->vl.perm = pp;
            }
            const int vsm = (int)((sizeof(double) * 7 + sizeof(int)) * (size_t)dm.Np);
            cudaError_t ev = cudaFuncSetAttribute(verlet_step_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, vsm);
            if (ev == cudaSuccess) ev = cudaFuncSetAttribute(verlet_step_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, vsm);
            if (ev == cudaSuccess) ev = cudaFuncSetAttribute(verlet_step_kernel<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, vsm);
            if (ev == cudaSuccess) ev = cudaFuncSetAttribute(verlet_step_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, vsm);
            if (ev == cudaSuccess) ev = cudaFuncSetAttribute(verlet_step_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, vsm);
            if (ev != cudaSuccess) { cudaGetLastError(); m->vl_fused = false; }
        }
    }
    DA(m->partA, Rm * (size_t)std::max(m->nA, 1));
    DA(m->partP, Rm * (size_t)S * m->nPx * 3);
    DA(m->partK, Rm * std::max((size_t)L.n_slices * m->nktg * m->gw, (size_t)std::max(m->n_quads, m->n_quads_split) * EHIC_LANE_WARPS * m->nkt));
    DA(m->energies, Rm * 4); DA(m->kinetic, Rm);
    {
        auto &w = m->wss;
        w.xs = (size_t)S * 3 * dm.Np; w.xt = (size_t)m->nkt * dm.Np * m->kt * 3; w.wbuf = (size_t)dm.wstride;
        w.gprior = (size_t)S * N * 3; w.bbox = (size_t)S * L.n_slices * 6; w.partC = (size_t)S * 3; w.cell = (size_t)S * N;
        w.partA = (size_t)std::max(m->nA, 1); w.partP = (size_t)S * m->nPx * 3;
        w.partK = std::max((size_t)L.n_slices * m->nktg * m->gw, (size_t)std::max(m->n_quads, m->n_quads_split) * EHIC_LANE_WARPS * m->nkt);
    }
    if (m->use_tiles) {
        m->nA_tile = m->td.n_tiles;                                   // one forward CTA (partial) per tile
        m->nKc = (N + 255) / 256;
        DA(m->tw, Rm * (size_t)m->td.slots);
        DA(m->Grow, Rm * S * (size_t)N * 3);
        DA(m->Pbuf, Rm * S * (size_t)m->td.n_rb * N * 3);
        double *pa2 = nullptr, *pk2 = nullptr;                         // partials sized for the tile kernels
        DA(pa2, Rm * (size_t)std::max(std::max(m->nA_tile, m->nA), 1)); m->partA = pa2;
        DA(pk2, Rm * (size_t)std::max<size_t>((size_t)S * m->nKc, (size_t)L.n_slices * m->nktg * m->gw)); m->partK = pk2;
        m->wss.partA = (size_t)std::max(std::max(m->nA_tile, m->nA), 1);
        m->wss.partK = std::max<size_t>((size_t)S * m->nKc, (size_t)L.n_slices * m->nktg * m->gw);
        m->wss.tw = (size_t)m->td.slots; m->wss.Grow = (size_t)S * N * 3; m->wss.Pbuf = (size_t)S * m->td.n_rb * N * 3;
        if (cudaMemset(m->tw, 0, Rm * (size_t)m->td.slots * sizeof(double)) != cudaSuccess) {
            ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "cudaMemset failed");
        }
        if (m->f32) {
            const size_t nb32 = sizeof(float) * Rm * (size_t)m->td.slots;
            if (cudaMalloc(&m->tw32, nb32) != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_NOMEM, "cudaMalloc failed"); }
            m->owned.push_back(m->tw32);
            cudaError_t e3 = cudaMemset(m->tw32, 0, nb32);
            const int fsm32 = (int)forward_tile32_smem_bytes(m->kt);
            if (e3 == cudaSuccess) {
                switch (m->kt) {
                case 4: e3 = cudaFuncSetAttribute(forward_tile32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm32); break;
                case 2: e3 = cudaFuncSetAttribute(forward_tile32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm32); break;
                default: e3 = cudaFuncSetAttribute(forward_tile32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm32); break;
                }
            }
            if (e3 != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "FP32 mode setup failed"); }
        }
        // both tile kernels use more than the 48 KB of static shared memory
        cudaError_t e1 = cudaFuncSetAttribute(backward_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)((EHIC_BT_STAGES * 2048 * sizeof(double) + 64)));
        if (e1 == cudaSuccess)
            e1 = cudaFuncSetAttribute(backward_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((EHIC_BT_STAGES * 2048 * sizeof(double) + 64)));
        cudaError_t e2 = cudaSuccess;
        const int fsm = (int)forward_tile_smem(m->kt);
        switch (m->kt) {
        case 4: e2 = cudaFuncSetAttribute(forward_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm); break;
        case 2: e2 = cudaFuncSetAttribute(forward_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm); break;
        default: e2 = cudaFuncSetAttribute(forward_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm); break;
        }
        if (e1 != cudaSuccess || e2 != cudaSuccess) { ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "cudaFuncSetAttribute failed"); }
    }
#undef DA
    // weights of padding slots are never read (ent_j < 0), but keep the buffer defined
    if (cudaMemset(m->wbuf, 0, Rm * (size_t)dm.wstride * sizeof(double)) != cudaSuccess) {
        ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "cudaMemset failed");
    }
    bool ok = cudaStreamCreateWithFlags(&m->st_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&m->st_stream2, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&m->st_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&m->st_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int q = 0; ok && q < ehic_model::kMaxChunks; q++)
        ok = cudaEventCreateWithFlags(&m->ev_in[q], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->ev_done[q], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->ev_half, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->ev_half_done, cudaEventDisableTiming) == cudaSuccess;
    for (int q = 0; ok && q < 2; q++)
        ok = cudaStreamCreateWithFlags(&m->st_side[q], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->ev_fork_l[q], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->ev_join_l[q], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { ehic_model_destroy(m); return fail(EHIC_ERR_CUDA, "stream / event creation failed"); }
    *out = m;
    return EHIC_OK;
}

int ehic_model_destroy(ehic_model_t *m)
{
    if (!m) return EHIC_OK;
    cudaSetDevice(m->device);
    drop_graphs(m);
    for (void *p : m->owned) cudaFree(p);
    cudaFree(m->st_x); cudaFree(m->st_g); cudaFree(m->st_md); cudaFree(m->st_small);
    if (m->st_stream) cudaStreamDestroy(m->st_stream);
    if (m->st_stream2) cudaStreamDestroy(m->st_stream2);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->ev_half) cudaEventDestroy(m->ev_half);
    if (m->ev_half_done) cudaEventDestroy(m->ev_half_done);
    for (int q = 0; q < 2; q++) {
        if (m->st_side[q]) cudaStreamDestroy(m->st_side[q]);
        if (m->ev_fork_l[q]) cudaEventDestroy(m->ev_fork_l[q]);
        if (m->ev_join_l[q]) cudaEventDestroy(m->ev_join_l[q]);
    }
    if (m->st_in) cudaStreamDestroy(m->st_in);
    if (m->st_out) cudaStreamDestroy(m->st_out);
    for (int q = 0; q < ehic_model::kMaxChunks; q++) { if (m->ev_in[q]) cudaEventDestroy(m->ev_in[q]); if (m->ev_done[q]) cudaEventDestroy(m->ev_done[q]); }
    delete m;
    return EHIC_OK;
}

int ehic_model_set_param(ehic_model_t *m, int which, double v)
{
    if (!m) return fail(EHIC_ERR_INVALID, "null model");
    double cur = 0.0;
    if (ehic_model_get_param(m, which, &cur) == EHIC_OK && cur == v) return EHIC_OK;
    drop_graphs(m);       // kernel arguments are baked into captured graphs
    switch (which) {
    case EHIC_PARAM_ALPHA: m->dm.alpha = v; break;
    case EHIC_PARAM_NONBONDED_K: m->dm.k_nb = v; break;
    case EHIC_PARAM_BACKBONE_K: m->dm.k_bb = v; break;
    case EHIC_PARAM_SPHERE_RADIUS: m->dm.sph_R = v; break;
    case EHIC_PARAM_SPHERE_K: m->dm.sph_k = v; break;
    case EHIC_PARAM_SPHERE_CX: m->dm.sph_cx = v; break;
    case EHIC_PARAM_SPHERE_CY: m->dm.sph_cy = v; break;
    case EHIC_PARAM_SPHERE_CZ: m->dm.sph_cz = v; break;
    default: return fail(EHIC_ERR_INVALID, "unknown parameter id");
    }
    return EHIC_OK;
}

int ehic_model_get_param(const ehic_model_t *m, int which, double *v)
{
    if (!m || !v) return fail(EHIC_ERR_INVALID, "null argument");
    switch (which) {
    case EHIC_PARAM_ALPHA: *v = m->dm.alpha; break;
    case EHIC_PARAM_NONBONDED_K: *v = m->dm.k_nb; break;
    case EHIC_PARAM_BACKBONE_K: *v = m->dm.k_bb; break;
    case EHIC_PARAM_SPHERE_RADIUS: *v = m->dm.sph_R; break;
    case EHIC_PARAM_SPHERE_K: *v = m->dm.sph_k; break;
    case EHIC_PARAM_SPHERE_CX: *v = m->dm.sph_cx; break;
    case EHIC_PARAM_SPHERE_CY: *v = m->dm.sph_cy; break;
    case EHIC_PARAM_SPHERE_CZ: *v = m->dm.sph_cz; break;
    default: return fail(EHIC_ERR_INVALID, "unknown parameter id");
    }
    return EHIC_OK;
}

int ehic_model_info(const ehic_model_t *m, int64_t out[8])
{
    if (!m || !out) return fail(EHIC_ERR_INVALID, "null argument");
    out[0] = m->dm.N; out[1] = m->dm.S; out[2] = m->dm.M; out[3] = m->max_replicas;
    out[4] = m->flags; out[5] = m->dm.tot; out[6] = m->device; out[7] = m->kt * 100 + m->gw + (m->use_tiles ? 10000 : 0) + (m->use_lanes ? 20000 : 0);
    return EHIC_OK;
}

}  // extern "C"

// ---- internal launch helpers ------------------------------------------------------------

static int check_call(ehic_model *m, int R)
{
    if (!m) return fail(EHIC_ERR_INVALID, "null model");
    if (R <= 0 || R > m->max_replicas) return fail(EHIC_ERR_INVALID, "R must be in [1, max_replicas]");
    CU(cudaSetDevice(m->device));
    return EHIC_OK;
}

static int launch_pack(ehic_model *m, int R, const double *d_x, double *xs, double *xt, cudaStream_t st)
{
    long long rows = (long long)R * m->dm.S, total = rows * m->dm.Np;
    KernelTimer kt_(m, KC_PACK, st);
    if (m->use_lanes && xt) {
        dim3 grid((unsigned)(m->dm.Np / 32), (unsigned)m->ld.nkg, (unsigned)R);
        pack_lane_kernel<<<grid, 256, 0, st>>>(d_x, xs, xt, m->dm.N, m->dm.Np, m->dm.S, m->ld.nkg, m->f32 ? 1 : 0);
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_x, xs, xt, m->dm.N, m->dm.Np, m->dm.S, m->kt, m->nkt, rows, m->use_lanes ? 1 : (m->f32 ? 2 : 0));
    LAUNCHED();
    CU(cudaGetLastError());
    return EHIC_OK;
}

static int launch_forward(ehic_model *m, int R, const double *xs, const double *d_norm, double *d_md,
                          bool want_w, bool want_logl, cudaStream_t st, const double *d_lammda = nullptr)
{
    if (m->dm.M == 0) {
        if (want_logl) CU(cudaMemsetAsync(m->partA, 0, sizeof(double) * (size_t)R * std::max(m->nA, 1), st));
        return EHIC_OK;
    }
    KernelTimer kt_(m, KC_FORWARD, st);
    if (m->use_tiles) {
        dim3 gridt((unsigned)m->td.n_tiles, (unsigned)R);
        if (m->f32) {
            float *twp32 = want_w ? m->tw32 : nullptr;
            const size_t sm32 = forward_tile32_smem_bytes(m->kt);
            switch (m->kt) {
            case 4: forward_tile32_kernel<4><<<gridt, 256, sm32, st>>>(m->dm, m->td, m->dc32, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp32, m->partA, want_logl, m->nkt); break;
            case 2: forward_tile32_kernel<2><<<gridt, 256, sm32, st>>>(m->dm, m->td, m->dc32, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp32, m->partA, want_logl, m->nkt); break;
            default: forward_tile32_kernel<1><<<gridt, 256, sm32, st>>>(m->dm, m->td, m->dc32, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp32, m->partA, want_logl, m->nkt); break;
            }
            LAUNCHED();
            CU(cudaGetLastError());
            return EHIC_OK;
        }
        double *twp = want_w ? m->tw : nullptr;
        const size_t sm = forward_tile_smem(m->kt);
        switch (m->kt) {
        case 4: forward_tile_kernel<4><<<gridt, 256, sm, st>>>(m->dm, m->td, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp, m->partA, want_logl, m->nkt); break;
        case 2: forward_tile_kernel<2><<<gridt, 256, sm, st>>>(m->dm, m->td, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp, m->partA, want_logl, m->nkt); break;
        default: forward_tile_kernel<1><<<gridt, 256, sm, st>>>(m->dm, m->td, m->xt[m->cur_xt], d_norm, d_lammda, d_md, twp, m->partA, want_logl, m->nkt); break;
        }
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    dim3 grid((unsigned)m->nA, (unsigned)R);
    if (m->use_lanes) {
        const long long nv = (long long)R * m->nA;
        if (m->lane_res_fwd && !m->in_split && nv >= 2LL * m->n_sm) {              // a replica's pairs on one SM: its lane copy stays in L1
            LaneSched sc{EHIC_LANE_FWD_GROUPS, nv, m->nA, nullptr};
            if (m->f32)
                forward_lane32_resident_kernel<<<m->n_sm, EHIC_LANE_FWD_GROUPS * 256, 0, st>>>(
                    m->dm, m->ld, sc, m->xt[m->cur_xt], d_norm, d_md, want_w ? reinterpret_cast<float *>(m->wbuf) : nullptr, m->partA, want_logl);
            else
                forward_lane_resident_kernel<<<m->n_sm, EHIC_LANE_FWD_GROUPS * 256, 0, st>>>(
                    m->dm, m->ld, sc, m->xt[m->cur_xt], d_norm, d_md, want_w ? m->wbuf : nullptr, m->partA, want_logl);
        } else if (m->f32)
            forward_lane32_kernel<<<grid, 256, 0, st>>>(m->dm, m->ld, m->xt[m->cur_xt], d_norm, d_md,
                                                         want_w ? reinterpret_cast<float *>(m->wbuf) : nullptr, m->partA, want_logl);
        else
            forward_lane_kernel<<<grid, 256, 0, st>>>(m->dm, m->ld, m->xt[m->cur_xt], d_norm, d_md, want_w ? m->wbuf : nullptr,
                                                       m->partA, want_logl);
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    static const bool fwd_soa = []() { const char *e = getenv("EHIC_FORWARD"); return e && !strcmp(e, "soa"); }();
    if (m->flags & EHIC_FLAG_EXACT)
        forward_kernel<true><<<grid, 256, 0, st>>>(m->dm, xs, d_norm, d_md, want_w ? m->wbuf : nullptr, m->partA, want_logl);
    else if (fwd_soa)
        forward_kernel<false><<<grid, 256, 0, st>>>(m->dm, xs, d_norm, d_md, want_w ? m->wbuf : nullptr, m->partA, want_logl);
    else {
        double *wb = want_w ? m->wbuf : nullptr;
        switch (m->kt) {
        case 4: forward_rec_kernel<4><<<grid, 256, 0, st>>>(m->dm, m->xt[m->cur_xt], d_norm, d_md, wb, m->partA, want_logl, m->nkt); break;
        case 2: forward_rec_kernel<2><<<grid, 256, 0, st>>>(m->dm, m->xt[m->cur_xt], d_norm, d_md, wb, m->partA, want_logl, m->nkt); break;
        default: forward_rec_kernel<1><<<grid, 256, 0, st>>>(m->dm, m->xt[m->cur_xt], d_norm, d_md, wb, m->partA, want_logl, m->nkt); break;
        }
    }
    LAUNCHED();
    CU(cudaGetLastError());
    return EHIC_OK;
}

static int launch_prior(ehic_model *m, int R, int terms, const double *xs, const double *d_beta,
                        bool want_grad, bool want_energy, cudaStream_t st, bool verlet = false)
{
    verlet = verlet && m->use_verlet && (terms & EHIC_TERM_NONBONDED);
    dim3 grid((unsigned)m->nPx, (unsigned)m->dm.S, (unsigned)R);
    double *gp = want_grad ? m->gprior : nullptr, *pp = want_energy ? m->partP : nullptr;
    KernelTimer kt_(m, KC_PRIOR, st);
    const bool ex = (m->flags & EHIC_FLAG_EXACT) != 0;
    // compact structures go through the cell grid, extended ones through the box-pruned all-pairs sweep;
    // the two kernels take the same per-structure decision from the slice boxes
    const bool cells = m->use_cells && (terms & EHIC_TERM_NONBONDED);
    m->last_prior_used_cells = cells || verlet;          // partC carries part of the nonbonded energy
    // Verlet trajectories of small systems (< 1024 beads: few slices, compact structures) do without the slice
    // boxes -- list rebuilds are rare and the box kernel costs more per step than its pruning saves per rebuild
    // (nora2012: 0.41 -> 0.37 ms per step); for long chains the pruned rebuild wins (config D: RE sweeps 1.765
    // with boxes, 1.738 without)
    if (verlet && m->vl_fused && !cells) {
        // one launch per step: displacement check, conditional list rebuild, list (or overflow all-pairs) force, backbone,
        // sphere and the energy sums of every structure (kernels.cuh: verlet_step_kernel)
        const unsigned n_struct = (unsigned)((long long)R * m->dm.S);
        const size_t vsm = (sizeof(double) * 7 + sizeof(int)) * (size_t)m->dm.Np;
        double *pc = want_energy ? m->partC : nullptr;
        if (want_energy) CU(cudaMemsetAsync(m->partP, 0, sizeof(double) * 3 * (size_t)R * m->dm.S * m->nPx, st));   // nothing from prior_kernel
        const int N = m->dm.N;
        if (N <= 128) verlet_step_kernel<128><<<n_struct, 128, vsm, st>>>(m->dm, m->vl, xs, d_beta, gp, pc, terms);
        else if (N <= 256) verlet_step_kernel<256><<<n_struct, 256, vsm, st>>>(m->dm, m->vl, xs, d_beta, gp, pc, terms);
        else if (N <= 320) verlet_step_kernel<320><<<n_struct, 320, vsm, st>>>(m->dm, m->vl, xs, d_beta, gp, pc, terms);
        else if (N <= 384) verlet_step_kernel<384><<<n_struct, 384, vsm, st>>>(m->dm, m->vl, xs, d_beta, gp, pc, terms);
        else verlet_step_kernel<512><<<n_struct, 512, vsm, st>>>(m->dm, m->vl, xs, d_beta, gp, pc, terms);
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    static const int vl_boxes = []() { const char *e = getenv("EHIC_VERLET_BOXES"); return e ? atoi(e) : -1; }();
    const bool no_boxes = verlet && (vl_boxes >= 0 ? vl_boxes == 0 : m->dm.N < 1024);
    const double *boxes = no_boxes ? nullptr : m->bbox;
    if ((terms & EHIC_TERM_NONBONDED) && !no_boxes) {
        dim3 gb((unsigned)((long long)R * m->dm.S), (unsigned)m->dm.n_slices);
        bbox_kernel<<<gb, 32, 0, st>>>(xs, m->bbox, m->dm.N, m->dm.Np, m->dm.n_slices);
        LAUNCHED();
    }
    const int *vls = verlet ? m->vl.state : nullptr;
    if (verlet) {
        const unsigned n_struct = (unsigned)((long long)R * m->dm.S);
        verlet_check_kernel<<<n_struct, 256, 0, st>>>(m->dm, m->vl, xs);
        LAUNCHED();
        switch (m->prior_threads) {
        case 64: verlet_build_kernel<64><<<grid, 64, 0, st>>>(m->dm, m->vl, xs, m->r_max, boxes); break;
        case 128: verlet_build_kernel<128><<<grid, 128, 0, st>>>(m->dm, m->vl, xs, m->r_max, boxes); break;
        default: verlet_build_kernel<256><<<grid, 256, 0, st>>>(m->dm, m->vl, xs, m->r_max, boxes); break;
        }
        LAUNCHED();
    }
#define PK(T)                                                                                         \
    do {                                                                                              \
        if (ex) prior_kernel<true, T><<<grid, T, 0, st>>>(m->dm, xs, d_beta, gp, pp, terms, m->r_max, m->bbox, 0);   \
        else prior_kernel<false, T><<<grid, T, 0, st>>>(m->dm, xs, d_beta, gp, pp, terms, m->r_max, boxes, cells ? 1 : 0, vls);     \
    } while (0)
    switch (m->prior_threads) {
    case 64: PK(64); break;
    case 128: PK(128); break;
    default: PK(256); break;
    }
#undef PK
    LAUNCHED();
    if (verlet) {
        const int vth = std::min(512, std::max(64, (m->dm.N + 31) / 32 * 32));
        verlet_force_kernel<<<(unsigned)((long long)R * m->dm.S), vth, 0, st>>>(m->dm, m->vl, xs, d_beta, gp,
                                                                                 want_energy ? m->partC : nullptr, terms);
        LAUNCHED();
    }
    if (cells) {
        dim3 gc((unsigned)m->dm.S, (unsigned)R);
        cell_nonbonded_kernel<<<gc, EHIC_CELL_THREADS, 0, st>>>(m->dm, xs, d_beta, gp, want_energy ? m->partC : nullptr,
                                                               m->r_max, m->cell_list, m->bbox);
        LAUNCHED();
    }
    CU(cudaGetLastError());
    return EHIC_OK;
}

template <bool EXACT>
static void launch_gather_t(ehic_model *m, const GatherArgs &ga, int R, cudaStream_t st)
{
    unsigned blocks = (unsigned)((long long)R * m->dm.n_slices * m->nktg);
    const int threads = m->gw * 32;
    switch (m->kt) {
    case 4: gather_kernel<EXACT, 4><<<blocks, threads, 0, st>>>(m->dm, ga, m->nkt, m->nktg); break;
    case 2: gather_kernel<EXACT, 2><<<blocks, threads, 0, st>>>(m->dm, ga, m->nkt, m->nktg); break;
    default: gather_kernel<EXACT, 1><<<blocks, threads, 0, st>>>(m->dm, ga, m->nkt, m->nktg); break;
    }
    LAUNCHED();
}

static int launch_gather(ehic_model *m, const GatherArgs &ga, int R, cudaStream_t st)
{
    if (m->use_tiles) {
        if (ga.contact && m->f32) {
            KernelTimer kt_(m, KC_GATHER, st);
            BackwardTile32Args b32{};
            b32.xs = ga.xs_cur; b32.dc32 = m->dc32; b32.tw32 = m->tw32; b32.Grow = m->Grow; b32.P = m->Pbuf;
            b32.tile_clean = m->tile_clean; b32.n_kg = (m->dm.S + 3) / 4;
            const unsigned blocks = (unsigned)((long long)R * m->td.n_rb * b32.n_kg);
            backward_tile32_kernel<<<blocks, 128, EHIC_BT_STAGES * 2048 * sizeof(float) + 64, st>>>(m->dm, m->td, b32);
            LAUNCHED();
            CU(cudaGetLastError());
        } else if (ga.contact) {
            KernelTimer kt_(m, KC_GATHER, st);
            BackwardTileArgs ba{};
            ba.xs = ga.xs_cur; ba.tw = m->tw; ba.beta = ga.beta; ba.Grow = m->Grow; ba.P = m->Pbuf;
            ba.tile_clean = m->tile_clean; ba.n_kg = (m->dm.S + 3) / 4; ba.nb = ga.nb_fused; ba.r_max = m->r_max;
            ba.bbox = m->bbox;
            if (ga.nb_fused) {
                dim3 gb((unsigned)((long long)R * m->dm.S), (unsigned)m->dm.n_slices);
                bbox_kernel<<<gb, 32, 0, st>>>(ga.xs_cur, m->bbox, m->dm.N, m->dm.Np, m->dm.n_slices);
                LAUNCHED();
            }
            const unsigned blocks = (unsigned)((long long)R * m->td.n_rb * ba.n_kg);
            const size_t sm = (EHIC_BT_STAGES * 2048 * sizeof(double) + 64);
            if (ga.nb_fused) backward_tile_kernel<true><<<blocks, 128, sm, st>>>(m->dm, m->td, ba);
            else backward_tile_kernel<false><<<blocks, 128, sm, st>>>(m->dm, m->td, ba);
            LAUNCHED();
            CU(cudaGetLastError());
        }
        KernelTimer kt2_(m, KC_OTHER, st);
        dim3 grid((unsigned)m->nKc, (unsigned)m->dm.S, (unsigned)R);
        combine_kernel<<<grid, 256, 0, st>>>(m->dm, m->td, ga, m->Grow, m->Pbuf, ga.contact, m->kt, m->nkt, m->f32 ? 1 : 0);
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    KernelTimer kt_(m, KC_GATHER, st);
    if (m->use_lanes) {
        // few replicas of a medium system: the split work list (long rows over 2 or 4 warps of a CTA)
        const bool split = m->lane_split_mode >= 0 ? m->lane_split_mode == 1
                                                   : (2 * m->dm.M <= 200000 && (long long)R * m->ld.nkg * m->n_quads < 3LL * 8 * m->n_sm);
        LaneDev ldv = m->ld;
        const int nq = split ? m->n_quads_split : m->n_quads;
        if (split) ldv.items = m->lane_items_split;
        const unsigned blocks = (unsigned)((long long)R * m->ld.nkg * nq);
        if (!split && m->lane_res_bwd && (long long)blocks >= 2LL * EHIC_LANE_BWD_GROUPS * m->n_sm) {
            LaneSched sc{EHIC_LANE_BWD_GROUPS, (long long)blocks, m->ld.nkg * m->n_quads, m->lane_cum};
            const int th = EHIC_LANE_BWD_GROUPS * EHIC_LANE_WARPS * 32;
            if (m->f32) backward_lane32_resident_kernel<<<m->n_sm, th, 0, st>>>(m->dm, m->ld, sc, ga, m->n_quads);
            else backward_lane_resident_kernel<<<m->n_sm, th, 0, st>>>(m->dm, m->ld, sc, ga, m->n_quads);
        } else if (m->f32) backward_lane32_kernel<<<blocks, EHIC_LANE_WARPS * 32, 0, st>>>(m->dm, ldv, ga, nq);
        else backward_lane_kernel<<<blocks, EHIC_LANE_WARPS * 32, 0, st>>>(m->dm, ldv, ga, nq);
        m->last_lane_quads = nq;
        LAUNCHED();
        CU(cudaGetLastError());
        return EHIC_OK;
    }
    if (m->flags & EHIC_FLAG_EXACT) launch_gather_t<true>(m, ga, R, st);
    else launch_gather_t<false>(m, ga, R, st);
    CU(cudaGetLastError());
    return EHIC_OK;
}

static int launch_finalize(ehic_model *m, int R, bool haveA, bool haveP, bool haveK,
                           double *d_energies, double *d_kin, cudaStream_t st)
{
    KernelTimer kt_(m, KC_FINALIZE, st);
    finalize_kernel<<<R, 256, 0, st>>>(m->dm, haveA ? m->partA : nullptr, m->use_tiles ? m->nA_tile : m->nA,
                                       haveP ? m->partP : nullptr, (long long)m->dm.S * m->nPx,
                                       haveK ? m->partK : nullptr,
                                       m->use_tiles ? (long long)m->dm.S * m->nKc
                                       : m->use_lanes ? (long long)m->ld.nkg * (m->last_lane_quads ? m->last_lane_quads : m->n_quads) * EHIC_LANE_WARPS
                                                      : (long long)m->dm.n_slices * m->nktg * m->gw,
                                       (haveP && m->last_prior_used_cells) ? m->partC : nullptr, m->dm.S,
                                       d_energies, d_kin);
    LAUNCHED();
    CU(cudaGetLastError());
    return EHIC_OK;
}

// RAII: rebases the per-replica workspace pointers of the model to replica slot `s` while a chunk is being
// enqueued (kernel arguments are captured at launch), so that two chunks can be in flight at once.
struct SlotShift {
    ehic_model *m;
    double *xs0, *xt0, *wbuf, *tw, *gprior, *bbox, *partC, *partA, *partP, *partK, *Grow, *Pbuf;
    float *tw32;
    int *cell;
    double *xs1, *xt1, *energies, *kinetic;
    VerletDev vl;
    SlotShift(ehic_model *m_, int s) : m(m_)
    {
        xs1 = m->xs[1]; xt1 = m->xt[1]; energies = m->energies; kinetic = m->kinetic; vl = m->vl;
        {
            const size_t z = (size_t)s;
            m->xs[1] += z * m->wss.xs; m->xt[1] += z * m->wss.xt; m->energies += z * 4; m->kinetic += z;
            if (m->use_verlet) {
                const size_t sS = z * (size_t)m->dm.S;
                m->vl.state += sS; m->vl.stale += sS; m->vl.cnt += sS * m->dm.Np;
                m->vl.idx += sS * (size_t)m->dm.Np * EHIC_VL_CAP; m->vl.xref += z * m->wss.xs; m->vl.perm += sS * m->dm.Np;
            }
        }
        xs0 = m->xs[0]; xt0 = m->xt[0]; wbuf = m->wbuf; tw = m->tw; gprior = m->gprior; bbox = m->bbox; partC = m->partC;
        partA = m->partA; partP = m->partP; partK = m->partK; Grow = m->Grow; Pbuf = m->Pbuf; tw32 = m->tw32; cell = m->cell_list;
        const auto &w = m->wss;
        const size_t z = (size_t)s;
        m->xs[0] += z * w.xs; m->xt[0] += z * w.xt; m->wbuf += z * w.wbuf; m->gprior += z * w.gprior; m->bbox += z * w.bbox;
        m->partC += z * w.partC; m->partA += z * w.partA; m->partP += z * w.partP; m->partK += z * w.partK;
        if (m->tw) m->tw += z * w.tw;
        if (m->tw32) m->tw32 += z * w.tw;
        if (m->Grow) m->Grow += z * w.Grow;
        if (m->Pbuf) m->Pbuf += z * w.Pbuf;
        if (m->cell_list) m->cell_list += z * w.cell;
    }
    ~SlotShift()
    {
        m->xs[0] = xs0; m->xt[0] = xt0; m->wbuf = wbuf; m->tw = tw; m->gprior = gprior; m->bbox = bbox; m->partC = partC;
        m->partA = partA; m->partP = partP; m->partK = partK; m->Grow = Grow; m->Pbuf = Pbuf; m->tw32 = tw32; m->cell_list = cell;
        m->xs[1] = xs1; m->xt[1] = xt1; m->energies = energies; m->kinetic = kinetic; m->vl = vl;
    }
};

extern "C" {

// ---- hot path, device buffers -----------------------------------------------------------

int ehic_mock_data(ehic_model_t *m, int R, const double *d_x, const double *d_norm, double *d_md, void *stream)
{
    int rc = check_call(m, R);
    if (rc) return rc;
    if (!d_x || !d_norm || !d_md) return fail(EHIC_ERR_INVALID, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_pack(m, R, d_x, m->xs[0], (m->flags & EHIC_FLAG_EXACT) ? nullptr : m->xt[0], st);
    if (rc) return rc;
    m->cur_xt = 0;
    return launch_forward(m, R, m->xs[0], d_norm, d_md, false, false, st);
}

int ehic_evaluate(ehic_model_t *m, int R, int terms, const double *d_x, const double *d_norm,
                  const double *d_lammda, const double *d_beta,
                  double *d_grad, double *d_energies, double *d_md, void *stream)
{
    int rc = check_call(m, R);
    if (rc) return rc;
    if (!d_x) return fail(EHIC_ERR_INVALID, "null coordinates");
    {   // Large batches run as two halves side by side (second half on st_stream2, in its own workspace slots):
        // the CTAs of one half fill the partial last waves of the other's kernels.  Measured at config D:
        // 1128 -> 1141 grad evals/s; four parts: 1134.  Only when a half is still a large launch.
        cudaStream_t st0 = (cudaStream_t)stream;
        const bool heavy = m->use_tiles && (double)m->dm.S * (double)m->dm.M * (R / 2) >= 2e8;     // see enqueue_trajectory
        const int parts = m->opt_split >= 0 ? m->opt_split : (heavy ? 2 : 1);
        if (parts >= 2 && R >= 2 * parts && !m->in_host_pipeline && !m->profiling && st0 != m->st_stream2) {
            const size_t per = (size_t)m->dm.S * m->dm.N * 3;
            PipelineScope scope(m);
            CU(cudaEventRecord(m->ev_half, st0));
            CU(cudaStreamWaitEvent(m->st_stream2, m->ev_half, 0));
            for (int c = 0; c < parts && rc == EHIC_OK; c++) {
                const int r0 = (int)((long long)R * c / parts), r1 = (int)((long long)R * (c + 1) / parts);
                SlotShift shift(m, r0);
                rc = ehic_evaluate(m, r1 - r0, terms, d_x + r0 * per, d_norm ? d_norm + r0 : nullptr, d_lammda ? d_lammda + r0 : nullptr,
                                   d_beta ? d_beta + r0 : nullptr, d_grad ? d_grad + r0 * per : nullptr,
                                   d_energies ? d_energies + 4 * (size_t)r0 : nullptr, d_md ? d_md + (size_t)r0 * m->dm.M : nullptr,
                                   (c & 1) ? m->st_stream2 : st0);
            }
            if (rc) return rc;
            CU(cudaEventRecord(m->ev_half_done, m->st_stream2));
            CU(cudaStreamWaitEvent(st0, m->ev_half_done, 0));
            return EHIC_OK;
        }
    }
    terms &= EHIC_TERM_ALL;
    if (!m->dm.sphere_active) terms &= ~EHIC_TERM_SPHERE;
    if (m->dm.M == 0) terms &= ~EHIC_TERM_LIKELIHOOD;         // no data points: the likelihood term is vacuous
    const bool lik = (terms & EHIC_TERM_LIKELIHOOD) != 0;
    const int prior_terms = terms & ~EHIC_TERM_LIKELIHOOD;
    if ((lik || d_md) && !d_norm) return fail(EHIC_ERR_INVALID, "norm required for the likelihood term");
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_pack(m, R, d_x, m->xs[0], (d_grad || !(m->flags & EHIC_FLAG_EXACT)) ? m->xt[0] : nullptr, st);
    if (rc) return rc;
    m->cur_xt = 0;
    const bool need_forward = (lik && (d_grad || d_energies)) || d_md;
    // pure gradient calls on the tile path evaluate the nonbonded force inside the backward tile kernel
    const bool fuse_nb = m->use_tiles && m->nb_fused && lik && (terms & EHIC_TERM_NONBONDED) && d_grad && !d_energies;
    const int kernel_prior_terms = fuse_nb ? (prior_terms & ~EHIC_TERM_NONBONDED) : prior_terms;
    const bool need_prior = kernel_prior_terms && (d_grad || d_energies);
    // the prior kernels only need the packed positions: beside the forward kernel, on a side stream
    static const bool fork_env = []() { const char *e = getenv("EHIC_FORK"); return !(e && !strcmp(e, "0")); }();
    // (not under per-kernel profiling, whose event pairs assume one stream; not when the call is one of several
    // parts already running side by side -- measured: no gain there, and the host pipeline loses 2 %)
    const bool fork = fork_env && need_forward && need_prior && !m->profiling && !m->in_host_pipeline;
    const int lane_id = st == m->st_stream2 ? 1 : 0;
    cudaStream_t side = fork ? m->st_side[lane_id] : st;
    if (fork) {
        CU(cudaEventRecord(m->ev_fork_l[lane_id], st));
        CU(cudaStreamWaitEvent(side, m->ev_fork_l[lane_id], 0));
    }
    if (need_forward) {
        rc = launch_forward(m, R, m->xs[0], d_norm, d_md, lik && d_grad, lik && d_energies, st, d_lammda);
        if (rc) return rc;
    }
    if (need_prior) {
        rc = launch_prior(m, R, kernel_prior_terms, m->xs[0], d_beta, d_grad != nullptr, d_energies != nullptr, side);
        if (rc) return rc;
        if (fork) {
            CU(cudaEventRecord(m->ev_join_l[lane_id], side));
            CU(cudaStreamWaitEvent(st, m->ev_join_l[lane_id], 0));
        }
    }
    if (d_grad) {
        GatherArgs ga{};
        ga.xt = m->xt[0]; ga.xs_cur = m->xs[0]; ga.wbuf = m->wbuf; ga.lammda = d_lammda;
        ga.gprior = need_prior ? m->gprior : nullptr;
        ga.grad = d_grad; ga.contact = lik ? 1 : 0;
        ga.nb_fused = fuse_nb ? 1 : 0; ga.beta = d_beta;
        rc = launch_gather(m, ga, R, st);
        if (rc) return rc;
    }
    if (d_energies) {
        rc = launch_finalize(m, R, lik && need_forward, need_prior, false, d_energies, nullptr, st);
        if (rc) return rc;
    }
    return EHIC_OK;
}

static int enqueue_trajectory(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                              const double *d_norm, const double *d_lammda, const double *d_beta,
                              const double *d_eps, double *d_H, cudaStream_t st);

// shared memory of the one-CTA-per-replica trajectory kernel: coordinates of all structures + row-slot weights
static size_t traj_small_smem(const ehic_model *m)
{
    return sizeof(double) * ((size_t)m->dm.S * m->dm.N * 3 + (size_t)m->dm.tot + 32);
}
static bool traj_small_ok(const ehic_model *m)
{
    return traj_small_smem(m) <= 200 * 1024 && !(m->flags & EHIC_FLAG_EXACT) && !m->f32 && !m->opt_traj_multi && !m->opt_traj_no_small &&
           !m->profiling;
}

}  // extern "C"

// ---- cluster trajectory kernel: choice of the cluster size ---------------------------------------------------------
template <int LK>
static cudaError_t tc_query(ehic_model *m, int C, size_t smem, int *n_clusters)
{
    const int slot = LK == 4 ? 0 : LK == 8 ? 1 : LK == 16 ? 2 : 3;
    if (!m->tc_attr[slot]) {
        cudaError_t e = cudaFuncSetAttribute(trajectory_cluster_kernel<LK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        m->tc_attr[slot] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(C * 64)); cfg.blockDim = dim3(EHIC_TC_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(n_clusters, trajectory_cluster_kernel<LK>, &cfg);
}

// One cluster of C CTAs per replica, the S structures split Sc = ceil(S / C) per CTA.  C is the feasible size (shared
// memory: coordinates of Sc structures + weight table + two chunk buffers) with the smallest estimated trajectory
// time: waves of co-resident clusters x per-CTA work ~ 1 / C, with a small penalty per doubling for the barriers.
static const ehic_model::TcPlan &plan_cluster(ehic_model *m, int R)
{
    for (const auto &p : m->tc_plans) if (p.R == R) return p;
    ehic_model::TcPlan best; best.R = R;
    const bool eligible = m->tc_possible && m->use_verlet && !(m->flags & EHIC_FLAG_EXACT) && !m->f32 && !m->opt_traj_multi && m->opt_cluster != 0;
    if (eligible) {
        const int N = m->dm.N, S = m->dm.S; const long long M = m->dm.M;
        double best_cost = 1e300;
        for (int C : {2, 4, 8}) {
            if (m->opt_cluster > 0 && C != m->opt_cluster) continue;
            if (C > S) continue;
            const int Sc = (S + C - 1) / C;
            if (Sc > 32) continue;                                            // (S >= C: every CTA holds at least one structure)
            int LK = 4; while (LK < Sc) LK *= 2;
            const size_t fixed = traj_cluster_smem(N, Sc, M, 0), cap = 227 * 1024;
            if (fixed + 16 * (size_t)C * 64 > cap) continue;
            long long chunk = (long long)((cap - fixed) / 16) / C * C;         // two buffers of `chunk` doubles
            const long long whole = (M + C - 1) / C * C;
            if (chunk > whole) chunk = whole;
            const int n_chunks = (int)((M + chunk - 1) / chunk);
            if (n_chunks > 32) continue;
            const size_t smem = traj_cluster_smem(N, Sc, M, (int)chunk);
            int conc = 0;
            cudaError_t e = LK == 4 ? tc_query<4>(m, C, smem, &conc) : LK == 8 ? tc_query<8>(m, C, smem, &conc)
                          : LK == 16 ? tc_query<16>(m, C, smem, &conc) : tc_query<32>(m, C, smem, &conc);
            if (e != cudaSuccess || conc < 1) { cudaGetLastError(); continue; }
            const int waves = (R + conc - 1) / conc;
            const double lane_eff = (double)(LK * C) / (double)S;              // idle lanes: LK > Sc, last CTA short
            const double cost = waves * lane_eff / C * (1.0 + 0.03 * std::log2((double)C) + 0.005 * n_chunks);
            if (cost < best_cost) {
                best_cost = cost;
                best.ok = true; best.C = C; best.Sc = Sc; best.LK = LK; best.chunk = (int)chunk; best.conc = conc; best.smem = smem;
            }
        }
    }
    if (m->tc_plans.size() >= 16) m->tc_plans.erase(m->tc_plans.begin());
    m->tc_plans.push_back(best);
    return m->tc_plans.back();
}

static int launch_cluster_trajectory(ehic_model *m, const ehic_model::TcPlan &pl, int R, int n_steps, double *d_x, double *d_p,
                                     const double *d_norm, const double *d_lammda, const double *d_beta,
                                     const double *d_eps, double *d_H, cudaStream_t st)
{
    if (!m->tc_pw) {
        const size_t n = (size_t)m->max_replicas * (size_t)(m->dm.S + 7) * m->dm.N * 3;
        CU(cudaMalloc(&m->tc_pw, sizeof(double) * n));
        m->owned.push_back(m->tc_pw);
    }
    TrajClusterArgs ta{};
    ta.x = d_x; ta.p = d_p; ta.norm = d_norm; ta.lammda = d_lammda; ta.beta = d_beta; ta.eps = d_eps; ta.H = d_H;
    ta.pw = m->tc_pw; ta.n_steps = n_steps; ta.C = pl.C; ta.Sc = pl.Sc; ta.chunk = pl.chunk;
    ta.r_max = m->r_max; ta.skin = m->vl.skin;
    ta.vl_idx = m->vl.idx; ta.vl_cnt = m->vl.cnt; ta.vl_state = m->vl.state;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(R * pl.C)); cfg.blockDim = dim3(EHIC_TC_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)pl.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    switch (pl.LK) {
    case 4: CU(cudaLaunchKernelEx(&cfg, trajectory_cluster_kernel<4>, m->dm, m->ld, ta)); break;
    case 8: CU(cudaLaunchKernelEx(&cfg, trajectory_cluster_kernel<8>, m->dm, m->ld, ta)); break;
    case 16: CU(cudaLaunchKernelEx(&cfg, trajectory_cluster_kernel<16>, m->dm, m->ld, ta)); break;
    default: CU(cudaLaunchKernelEx(&cfg, trajectory_cluster_kernel<32>, m->dm, m->ld, ta)); break;
    }
    LAUNCHED();
    return EHIC_OK;
}

extern "C" {

int ehic_trajectory_plan(const ehic_model_t *cm, int R, int n_steps, int64_t out[4])
{
    if (!cm || !out) return fail(EHIC_ERR_INVALID, "null argument");
    ehic_model *m = const_cast<ehic_model *>(cm);        // the plan cache is not part of the model's logical state
    int rc = check_call(m, R);
    if (rc) return rc;
    (void)n_steps;
    out[0] = 0; out[1] = 1; out[2] = 0; out[3] = 0;
    if (traj_small_ok(m)) { out[0] = 1; out[2] = (int64_t)traj_small_smem(m); out[3] = 1; return EHIC_OK; }
    if (m->profiling) return EHIC_OK;
    const ehic_model::TcPlan &pl = plan_cluster(m, R);
    if (pl.ok) { out[0] = 2; out[1] = pl.C; out[2] = (int64_t)pl.smem; out[3] = 1; }
    return EHIC_OK;
}

int ehic_hmc_trajectory(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                        const double *d_norm, const double *d_lammda, const double *d_beta,
                        const double *d_eps, double *d_H, void *stream)
{
    int rc = check_call(m, R);
    if (rc) return rc;
    if (!d_x || !d_p || !d_norm || !d_eps || !d_H) return fail(EHIC_ERR_INVALID, "null buffer");
    if (n_steps < 1) return fail(EHIC_ERR_INVALID, "n_steps must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    // small systems: the whole trajectory in ONE kernel launch, state resident in shared memory
    {
        const size_t sm = traj_small_smem(m);
        if (traj_small_ok(m)) {
            if (!m->traj_small_ready) {
                CU(cudaFuncSetAttribute(trajectory_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                m->traj_small_ready = true;
            }
            TrajSmallArgs ta{};
            ta.x = d_x; ta.p = d_p; ta.norm = d_norm; ta.lammda = d_lammda; ta.beta = d_beta; ta.eps = d_eps;
            ta.H = d_H; ta.n_steps = n_steps; ta.r_max = m->r_max;
            const int items = m->dm.S * m->dm.n_slices * 32;
            const int threads = items >= 512 || m->dm.M >= 512 ? 512 : (items > 128 || m->dm.M > 128 ? 256 : 128);
            int split = 1;
            while (split < 32 && items * split * 2 <= threads) split *= 2;
            ta.split = split;
            trajectory_small_kernel<<<R, threads, sm, st>>>(m->dm, ta);
            LAUNCHED();
            CU(cudaGetLastError());
            return EHIC_OK;
        }
    }
    // medium systems: one launch, one thread-block cluster per replica (traj_cluster.cuh)
    if (!m->profiling) {
        const ehic_model::TcPlan &pl = plan_cluster(m, R);
        if (pl.ok) return launch_cluster_trajectory(m, pl, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    }
    static const bool use_graphs = []() { const char *e = getenv("EHIC_GRAPH"); return !(e && !strcmp(e, "0")); }();
    if (!use_graphs || m->profiling)
        return enqueue_trajectory(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    // One graph launch per proposal: the 3 (L+1) + few kernels of a trajectory are captured once per
    // call signature and replayed, which removes the per-kernel launch latency that dominates the
    // small configurations (hairpin_s, proteins).
    std::vector<const void *> key = {(const void *)(intptr_t)R, (const void *)(intptr_t)n_steps, d_x, d_p, d_norm,
                                     d_lammda, d_beta, d_eps, d_H};
    for (auto &g : m->graphs)
        if (g.key == key) {
            CU(cudaGraphLaunch(g.exec, st));
            g_launches.fetch_add(g.n_kernels, std::memory_order_relaxed);
            return EHIC_OK;
        }
    // capture on the model's own stream (the legacy default stream cannot be captured)
    cudaStream_t cs = m->st_stream;
    const long long before = g_launches.load();
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return enqueue_trajectory(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    }
    rc = enqueue_trajectory(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, cs);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    const long long n_kernels = g_launches.load() - before;
    g_launches.store(before);
    if (rc != EHIC_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc != EHIC_OK) return rc;
        return enqueue_trajectory(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return enqueue_trajectory(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    }
    if (m->graphs.size() >= 8) { cudaGraphExecDestroy(m->graphs.front().exec); m->graphs.erase(m->graphs.begin()); }
    m->graphs.push_back({key, exec, n_kernels});
    CU(cudaGraphLaunch(exec, st));
    g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
    return EHIC_OK;
}

}  // extern "C"

static int enqueue_trajectory_part(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                                   const double *d_norm, const double *d_lammda, const double *d_beta,
                                   const double *d_eps, double *d_H, cudaStream_t st);

// Large batches: the two halves of the replicas as two independent chains on two streams (parallel branches of
// the captured graph), each in its own workspace slots -- the CTAs of one half fill the partial last waves of
// the other's kernels, as in ehic_evaluate.
static int enqueue_trajectory(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                              const double *d_norm, const double *d_lammda, const double *d_beta,
                              const double *d_eps, double *d_H, cudaStream_t st)
{
    // automatic on the dense tile path (FP64-bound kernels whose last waves are partial) ...
    const bool heavy = m->use_tiles && (double)m->dm.S * (double)m->dm.M * (R / 2) >= 2e8;
    // ... and on the lane path when the GPU holds FEW replicas (the ladder strong-scaled over several GPUs): every kernel
    // of a step is then two or three partial waves, and two half batches running out of phase fill each other's tails
    // (nora2012, 38 replicas: 192 -> 159 us per leapfrog step; 76: 332 -> 304; 151: 596 -> 571; 302: no difference)
    const bool light = m->use_lanes && (long long)R * m->ld.nkg * m->n_quads < 16LL * 8 * std::max(m->n_sm, 1);
    const int parts = m->opt_split >= 0 ? m->opt_split : ((heavy || light) ? 2 : 1);
    if (parts < 2 || R < 4 || m->profiling || st == m->st_stream2)
        return enqueue_trajectory_part(m, R, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    struct SplitScope { ehic_model *m; SplitScope(ehic_model *m_) : m(m_) { m->in_split = true; } ~SplitScope() { m->in_split = false; } } split_scope(m);
    const int h = R / 2;
    const size_t per = (size_t)m->dm.S * m->dm.N * 3;
    CU(cudaEventRecord(m->ev_half, st));
    CU(cudaStreamWaitEvent(m->st_stream2, m->ev_half, 0));
    int rc = enqueue_trajectory_part(m, h, n_steps, d_x, d_p, d_norm, d_lammda, d_beta, d_eps, d_H, st);
    if (rc == EHIC_OK) {
        SlotShift shift(m, h);
        rc = enqueue_trajectory_part(m, R - h, n_steps, d_x + h * per, d_p + h * per, d_norm + h,
                                     d_lammda ? d_lammda + h : nullptr, d_beta ? d_beta + h : nullptr, d_eps + h,
                                     d_H + 4 * (size_t)h, m->st_stream2);
    }
    // join even after an error, so that a stream capture in progress can be ended cleanly
    cudaEventRecord(m->ev_half_done, m->st_stream2);
    cudaStreamWaitEvent(st, m->ev_half_done, 0);
    return rc;
}

static int enqueue_trajectory_part(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                                   const double *d_norm, const double *d_lammda, const double *d_beta,
                                   const double *d_eps, double *d_H, cudaStream_t st)
{
    int rc = EHIC_OK;
    int prior_terms = EHIC_TERM_ALL & ~EHIC_TERM_LIKELIHOOD;
    if (!m->dm.sphere_active) prior_terms &= ~EHIC_TERM_SPHERE;
    rc = launch_pack(m, R, d_x, m->xs[0], m->xt[0], st);
    if (rc) return rc;
    if (m->use_verlet) {      // structures whose Verlet list overflowed in an earlier trajectory get a new attempt
        const long long n = (long long)R * m->dm.S;
        verlet_retry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->vl, n);
        LAUNCHED();
    }
    // The prior kernels (nonbonded lists, backbone, sphere) only need the positions: they run on a side stream
    // beside the forward kernel and join before the backward / gather kernel (parallel branches of the graph).
    static const bool fork_env = []() { const char *e = getenv("EHIC_FORK"); return !(e && !strcmp(e, "0")); }();
    const bool fork = fork_env && !m->profiling;
    const int lane_id = st == m->st_stream2 ? 1 : 0;         // second half of a split trajectory
    cudaStream_t side = fork ? m->st_side[lane_id] : st;
    cudaEvent_t ev_f = m->ev_fork_l[lane_id], ev_j = m->ev_join_l[lane_id];
    int cur = 0;
    for (int s = 0; s <= n_steps; s++) {
        const bool first = (s == 0), last = (s == n_steps);
        const bool ends = first || last;
        m->cur_xt = cur;
        const bool fuse_nb = m->use_tiles && m->nb_fused && !ends;
        const int kpt = fuse_nb ? (prior_terms & ~EHIC_TERM_NONBONDED) : prior_terms;
        if (fork && kpt) {
            CU(cudaEventRecord(ev_f, st));
            CU(cudaStreamWaitEvent(side, ev_f, 0));
        }
        rc = launch_forward(m, R, m->xs[cur], d_norm, nullptr, true, ends, st, d_lammda);
        if (rc) return rc;
        if (kpt) {
            rc = launch_prior(m, R, kpt, m->xs[cur], d_beta, true, ends, side, true);
            if (rc) return rc;
            if (fork) {
                CU(cudaEventRecord(ev_j, side));
                CU(cudaStreamWaitEvent(st, ev_j, 0));
            }
        }
        GatherArgs ga{};
        ga.xt = m->xt[cur]; ga.xs_cur = m->xs[cur]; ga.wbuf = m->wbuf; ga.lammda = d_lammda; ga.gprior = kpt ? m->gprior : nullptr;
        ga.nb_fused = fuse_nb ? 1 : 0; ga.beta = d_beta;
        ga.grad = nullptr; ga.partK = ends ? m->partK : nullptr; ga.contact = 1;
        ga.p = d_p; ga.eps = d_eps; ga.kick = ends ? 0.5 : 1.0;
        ga.xs_next = last ? nullptr : m->xs[cur ^ 1];
        ga.xt_next = last ? nullptr : m->xt[cur ^ 1];
        ga.kinetic_mode = first ? 1 : (last ? 2 : 0);
        rc = launch_gather(m, ga, R, st);
        if (rc) return rc;
        if (ends) {
            rc = launch_finalize(m, R, true, true, true, m->energies, m->kinetic, st);
            if (rc) return rc;
            hamiltonian_kernel<<<(R + 127) / 128, 128, 0, st>>>(m->energies, m->kinetic, d_lammda, d_beta, d_H, first ? 0 : 2, R);
            LAUNCHED();
            CU(cudaGetLastError());
        }
        if (!last) cur ^= 1;
    }
    long long rows = (long long)R * m->dm.S, total = rows * m->dm.N;
    unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m->xs[cur], d_x, m->dm.N, m->dm.Np, rows);
    LAUNCHED();
    CU(cudaGetLastError());
    return EHIC_OK;
}

extern "C" {

int ehic_select(int R, int64_t n, const int32_t *d_mask, const double *d_src, double *d_dst, void *stream)
{
    if (R <= 0 || n <= 0 || !d_mask || !d_src || !d_dst) return fail(EHIC_ERR_INVALID, "bad argument");
    long long total = (long long)R * n;
    select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, n, d_mask, d_src, d_dst);
    LAUNCHED();
    CU(cudaGetLastError());
    return EHIC_OK;
}

// ---- neighbour list -----------------------------------------------------------------------

int ehic_nblist_host(int n_beads, const double *h_coords, double cutoff,
                     int32_t *h_pairs, int64_t capacity, int64_t *n_pairs, int device)
{
    if (n_beads <= 0 || !h_coords || !n_pairs || capacity < 0 || (capacity > 0 && !h_pairs))
        return fail(EHIC_ERR_INVALID, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    CU(cudaSetDevice(device));
    const int N = n_beads;
    double *d_x = nullptr; int *d_cnt = nullptr; long long *d_off = nullptr; int *d_pairs = nullptr;
    int status = EHIC_OK;
    std::vector<int> cnt(N);
    std::vector<long long> off(N + 1, 0);
    const double cut2 = cutoff * cutoff;      // nblist.c:198 cellsize2 = cellsize * cellsize
    do {
#define CUB(call) if ((call) != cudaSuccess) { status = fail(EHIC_ERR_CUDA, #call " failed"); break; }
        CUB(cudaMalloc(&d_x, sizeof(double) * 3 * N));
        CUB(cudaMalloc(&d_cnt, sizeof(int) * N));
        CUB(cudaMalloc(&d_off, sizeof(long long) * N));
        CUB(cudaMemcpy(d_x, h_coords, sizeof(double) * 3 * N, cudaMemcpyHostToDevice));
        nblist_count_kernel<<<(N + 127) / 128, 128>>>(d_x, N, cut2, d_cnt);
        LAUNCHED();
        CUB(cudaMemcpy(cnt.data(), d_cnt, sizeof(int) * N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; i++) off[i + 1] = off[i] + cnt[i];
        *n_pairs = off[N];
        long long cap = std::min<long long>(capacity, off[N]);
        if (cap > 0) {
            CUB(cudaMalloc(&d_pairs, sizeof(int) * 2 * cap));
            CUB(cudaMemcpy(d_off, off.data(), sizeof(long long) * N, cudaMemcpyHostToDevice));
            nblist_fill_kernel<<<(N + 127) / 128, 128>>>(d_x, N, cut2, d_off, d_pairs, cap);
            LAUNCHED();
            CUB(cudaMemcpy(h_pairs, d_pairs, sizeof(int) * 2 * cap, cudaMemcpyDeviceToHost));
        }
        if (off[N] > capacity && capacity > 0) status = fail(EHIC_ERR_CAPACITY, "pair buffer too small");
#undef CUB
    } while (0);
    cudaFree(d_x); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_pairs);
    return status;
}

// ---- host-buffer entry points ---------------------------------------------------------------

static int ensure_staging(ehic_model *m)
{
    if (m->st_x) return EHIC_OK;
    const size_t Rm = (size_t)m->max_replicas;
    const size_t nx = Rm * m->dm.S * (size_t)m->dm.N * 3;
    CU(cudaMalloc(&m->st_x, sizeof(double) * nx));
    CU(cudaMalloc(&m->st_g, sizeof(double) * nx));
    CU(cudaMalloc(&m->st_md, sizeof(double) * std::max<size_t>(Rm * (size_t)m->dm.M, 1)));
    CU(cudaMalloc(&m->st_small, sizeof(double) * Rm * 8));
    return EHIC_OK;
}

int ehic_mock_data_host(ehic_model_t *m, int R, const double *h_x, const double *h_norm, double *h_md)
{
    return ehic_evaluate_host(m, R, 0, h_x, h_norm, nullptr, nullptr, nullptr, nullptr, h_md);
}

int ehic_evaluate_host(ehic_model_t *m, int R, int terms, const double *h_x, const double *h_norm,
                       const double *h_lammda, const double *h_beta,
                       double *h_grad, double *h_energies, double *h_md)
{
    int rc = check_call(m, R);
    if (rc) return rc;
    if (!h_x) return fail(EHIC_ERR_INVALID, "null coordinates");
    rc = ensure_staging(m);
    if (rc) return rc;
    // Software pipeline over chunks of replicas: all input copies are queued on one stream, the kernels of
    // consecutive chunks alternate between two compute streams (each chunk works in its own replica slots of
    // the workspace, so chunk c+1 fills the partial last wave of chunk c), and output copies follow on a
    // third stream.  Exposed: the first chunk's H2D and the last chunk's D2H.  Over PCIe the two 38 MB
    // copies of config D would otherwise add ~40 % to the step.
    cudaStream_t s_in = m->st_in, s_out = m->st_out;
    const size_t per = (size_t)m->dm.S * (size_t)m->dm.N * 3;
    const size_t M = (size_t)m->dm.M;
    double *d_norm = m->st_small, *d_lam = m->st_small + m->max_replicas, *d_bet = m->st_small + 2 * (size_t)m->max_replicas;
    double *d_E = m->st_small + 3 * (size_t)m->max_replicas;   // 4 * Rmax
    if (h_norm) CU(cudaMemcpyAsync(d_norm, h_norm, sizeof(double) * R, cudaMemcpyHostToDevice, s_in));
    if (h_lammda) CU(cudaMemcpyAsync(d_lam, h_lammda, sizeof(double) * R, cudaMemcpyHostToDevice, s_in));
    if (h_beta) CU(cudaMemcpyAsync(d_bet, h_beta, sizeof(double) * R, cudaMemcpyHostToDevice, s_in));
    static const int chunk_cap = []() { const char *e = getenv("EHIC_HOST_CHUNKS"); int v = e ? atoi(e) : 0;
                                        return (v >= 1 && v <= ehic_model::kMaxChunks) ? v : 4; }();      // 4 measured best at config D (8 replicas)
    const int n_chunks = std::min(R, chunk_cap);
    for (int c = 0; c < n_chunks; c++) {
        const int r0 = (int)((long long)R * c / n_chunks), r1 = (int)((long long)R * (c + 1) / n_chunks);
        CU(cudaMemcpyAsync(m->st_x + r0 * per, h_x + r0 * per, sizeof(double) * per * (r1 - r0), cudaMemcpyHostToDevice, s_in));
        CU(cudaEventRecord(m->ev_in[c], s_in));
    }
    for (int c = 0; c < n_chunks; c++) {
        const int r0 = (int)((long long)R * c / n_chunks), r1 = (int)((long long)R * (c + 1) / n_chunks);
        cudaStream_t st = (c & 1) ? m->st_stream2 : m->st_stream;
        CU(cudaStreamWaitEvent(st, m->ev_in[c], 0));
        {
            SlotShift shift(m, r0);
            PipelineScope scope(m);
            rc = ehic_evaluate(m, r1 - r0, terms, m->st_x + r0 * per, h_norm ? d_norm + r0 : nullptr,
                               h_lammda ? d_lam + r0 : nullptr, h_beta ? d_bet + r0 : nullptr,
                               h_grad ? m->st_g + r0 * per : nullptr, h_energies ? d_E + 4 * (size_t)r0 : nullptr,
                               h_md ? m->st_md + r0 * M : nullptr, st);
        }
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CU(cudaEventRecord(m->ev_done[c], st));
        CU(cudaStreamWaitEvent(s_out, m->ev_done[c], 0));
        if (h_grad) CU(cudaMemcpyAsync(h_grad + r0 * per, m->st_g + r0 * per, sizeof(double) * per * (r1 - r0), cudaMemcpyDeviceToHost, s_out));
        if (h_md && M > 0) CU(cudaMemcpyAsync(h_md + r0 * M, m->st_md + r0 * M, sizeof(double) * M * (r1 - r0), cudaMemcpyDeviceToHost, s_out));
    }
    if (h_energies) CU(cudaMemcpyAsync(h_energies, d_E, sizeof(double) * 4 * R, cudaMemcpyDeviceToHost, s_out));
    CU(cudaStreamSynchronize(s_out));
    CU(cudaStreamSynchronize(m->st_stream));
    CU(cudaStreamSynchronize(m->st_stream2));
    return EHIC_OK;
}

// ---- per-kernel timing ------------------------------------------------------------------------

int ehic_profile(ehic_model_t *m, int enable)
{
    if (!m) return fail(EHIC_ERR_INVALID, "null model");
    m->profiling = enable != 0;
    return EHIC_OK;
}

int ehic_profile_read(ehic_model_t *m, double ms_out[8], int64_t count_out[8])
{
    if (!m || !ms_out || !count_out) return fail(EHIC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(m->device));
    CU(cudaDeviceSynchronize());
    for (int q = 0; q < 8; q++) { ms_out[q] = 0.0; count_out[q] = 0; }
    for (auto &t : m->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) { ms_out[t.cls] += ms; count_out[t.cls]++; }
        cudaEventDestroy(t.a); cudaEventDestroy(t.b);
    }
    m->timed.clear();
    return EHIC_OK;
}

// ---- layout debug (host only) ----------------------------------------------------------------

int ehic_layout_debug(int32_t n_beads, int64_t n_data, const int64_t *data_points, int exact_order,
                      int64_t *h_sorted, int64_t *n_slots, int32_t *h_row_len)
{
    if (n_data > 0 && !data_points) return fail(EHIC_ERR_INVALID, "null data_points");
    ContactLayout L;
    std::string err = build_layout(n_beads, n_data, data_points, nullptr, exact_order != 0, L, false);
    if (!err.empty()) return fail(EHIC_ERR_INVALID, err);
    if (h_sorted) for (int64_t s = 0; s < n_data; s++) h_sorted[s] = L.perm[s];
    if (n_slots) *n_slots = L.tot;
    if (h_row_len) for (int32_t b = 0; b < n_beads; b++) h_row_len[b] = L.row_len[b];
    // self-check of the row view: every sorted pair must be found at its two slots
    for (int64_t s = 0; s < n_data; s++) {
        if (L.ent_j[L.slot_a[s]] != L.pj[s] || L.ent_j[L.slot_b[s]] != L.pi[s])
            return fail(EHIC_ERR_INVALID, "internal: row view inconsistent");
    }
    if (!exact_order) {     // same self-check for the lane view (CSR) of the sparse path
        LaneLayout C;
        err = build_lanes(L, C);
        if (!err.empty()) return fail(EHIC_ERR_INVALID, err);
        if (C.row_off[n_beads] != 2 * n_data) return fail(EHIC_ERR_INVALID, "internal: lane view has the wrong size");
        for (int64_t s = 0; s < n_data; s++) {
            const int64_t a = C.slot_a[s], b = C.slot_b[s];
            const bool ok = a >= C.row_off[L.pi[s]] && a < C.row_off[L.pi[s] + 1] && b >= C.row_off[L.pj[s]] &&
                            b < C.row_off[L.pj[s] + 1] && C.ent_j[a] == L.pj[s] && C.ent_j[b] == L.pi[s];
            if (!ok) return fail(EHIC_ERR_INVALID, "internal: lane view inconsistent");
        }
    }
    return EHIC_OK;
}

}  // extern "C"
