// layout.hpp -- host-side construction of the contact-list layout walked by the kernels.
//
// The reference walks data_points[M][3] row by row for every structure
// (evaluate_FWM_pureC.pxd:29-39, likelihoods_c.pyx:62-73) and scatter-adds into both beads
// of each pair.  The GPU path instead uses two static views of the same list, built once
// per model:
//
//  (1) the PAIR view: rows sorted by (i, j) so that one warp of the forward kernel reads a
//      run of pairs sharing bead i with consecutive beads j (coalesced coordinate loads);
//      `perm` maps a sorted position back to the row of data_points.
//  (2) the ROW view (sliced ELL, slice height 32 = one warp): for every bead b the list of
//      its partners o, so that the gradient is a gather  g_b += f_m * (x_o - x_b)  with no
//      atomics and a fixed summation order.  Slice s holds beads 32s..32s+31; entry e of
//      lane l lives at slice_off[s] + 32*e + l (coalesced across the warp).  Row order is
//      ascending partner index (fast mode) or data-point order (exact mode, which makes the
//      per-bead summation order identical to likelihoods_c.pyx:62-73).
//      slot_a[m] / slot_b[m] are the ROW-view positions of sorted pair m in row i and row j:
//      the forward kernel scatters the per-pair weight there.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <numeric>
#include <string>
#include <vector>

namespace ehic {

struct ContactLayout {
    int32_t N = 0;
    int64_t M = 0;
    // pair view (length M, sorted by (i, j))
    std::vector<int32_t> pi, pj;
    std::vector<double> pdc, pcnt;
    std::vector<int64_t> perm;
    std::vector<int32_t> slot_a, slot_b;
    // row view
    int32_t n_slices = 0;
    std::vector<int64_t> slice_off;   // n_slices + 1, in entries
    std::vector<int32_t> slice_w;     // n_slices
    std::vector<int32_t> row_len;     // N
    std::vector<int32_t> row_bead;    // n_slices * 32: bead handled by each lane slot (-1 = none).  Identity
                                      // for near-uniform rows; sorted by row length when that saves padding.
    std::vector<int32_t> ent_j;       // tot + kTailPad; padding slots hold a harmless partner (weight 0)
    std::vector<double> ent_dc;       // tot + kTailPad
    int64_t tot = 0;                  // slots that belong to slices (multiple of 32)
    static constexpr int64_t kTailPad = 16 * 32;   // so a 16-entry chunk copy never leaves the array
};

// (3) TILE view (dense lists only): every unordered pair {lo < hi} once, in the tile
//     (row block I = lo / 128, column slice B = hi / 32); a tile is 128 row beads x 32 column
//     beads = 4096 slots, slot = ((tile*32 + q)*4 + rb)*32 + lane with q = hi % 32,
//     rb = (lo / 32) % 4, lane = lo % 32.  The symmetric backward kernel walks the tiles of a row
//     block and accumulates action (row bead, registers) and reaction (column bead, warp
//     shuffles); the forward kernel walks the non-empty 32 x 32 sub-tiles (tile, rb).
struct TileLayout {
    bool eligible = false;          // dense enough, no duplicate pairs
    std::string why;                // reason when not eligible
    int32_t n_rb = 0;               // row blocks (128 beads)
    int32_t n_tiles = 0;
    std::vector<int32_t> tile_I, tile_B;      // n_tiles, sorted by (I, B)
    std::vector<int32_t> rb_begin;            // n_rb + 1: tile range of each row block
    std::vector<int32_t> sub_tile, sub_rb;    // non-empty 32x32 sub-tiles
    std::vector<uint32_t> mask;               // [n_tiles][4][32]: bit q = slot (q, rb, lane) present
    std::vector<double> dc, cnt;              // [n_tiles*4096]
    std::vector<int64_t> orig;                // [n_tiles*4096] row of data_points, -1 = absent
    std::vector<int32_t> exists;              // [n_rb][n_slices]: tile index or -1
    std::vector<unsigned char> clean;         // [n_tiles]: above the diagonal blocks, all slots present, no clamped beads
    bool covers_all_pairs = false;            // every tile (I, B >= 4 I) exists: the nonbonded force can ride along
    double fill = 0.0;
};

inline void build_tiles(const ContactLayout &L, const int64_t *data, const double *cds, double min_fill, TileLayout &T)
{
    T = TileLayout();
    const int32_t N = L.N; const int64_t M = L.M;
    if (M == 0 || N < 64) { T.why = "list too small"; return; }
    const int32_t n_sl = (N + 31) / 32;
    T.n_rb = (n_sl + 3) / 4;
    std::vector<int32_t> id((size_t)T.n_rb * n_sl, -1);
    std::vector<int64_t> per_tile((size_t)T.n_rb * n_sl, 0);
    for (int64_t m = 0; m < M; m++) {
        int64_t i = data[3 * m], j = data[3 * m + 1];
        int64_t lo = std::min(i, j), hi = std::max(i, j);
        per_tile[(lo / 128) * n_sl + hi / 32]++;
    }
    for (int32_t I = 0; I < T.n_rb; I++) {
        T.rb_begin.push_back(T.n_tiles);
        for (int32_t B = 0; B < n_sl; B++)
            if (per_tile[(size_t)I * n_sl + B] > 0) {
                id[(size_t)I * n_sl + B] = T.n_tiles++;
                T.tile_I.push_back(I); T.tile_B.push_back(B);
            }
    }
    T.rb_begin.push_back(T.n_tiles);
    T.fill = (double)M / ((double)T.n_tiles * 4096.0);
    if (T.fill < min_fill) { T.why = "tile fill below threshold"; return; }
    if ((int64_t)T.n_tiles * 4096 >= ((int64_t)1 << 31)) { T.why = "too many tile slots"; return; }
    const size_t slots = (size_t)T.n_tiles * 4096;
    T.mask.assign((size_t)T.n_tiles * 128, 0u);
    T.dc.assign(slots, 0.0); T.cnt.assign(slots, 0.0); T.orig.assign(slots, -1);
    for (int64_t m = 0; m < M; m++) {
        int64_t i = data[3 * m], j = data[3 * m + 1];
        int64_t lo = std::min(i, j), hi = std::max(i, j);
        int32_t t = id[(lo / 128) * n_sl + hi / 32];
        int32_t q = (int32_t)(hi % 32), rb = (int32_t)((lo / 32) % 4), lane = (int32_t)(lo % 32);
        size_t pos = (((size_t)t * 32 + q) * 4 + rb) * 32 + lane;
        if (T.orig[pos] >= 0) { T.why = "duplicate data points for one bead pair"; return; }
        T.orig[pos] = m; T.dc[pos] = cds[m]; T.cnt[pos] = (double)data[3 * m + 2];
        T.mask[((size_t)t * 4 + rb) * 32 + lane] |= (1u << q);
    }
    for (int32_t t = 0; t < T.n_tiles; t++)
        for (int32_t rb = 0; rb < 4; rb++) {
            uint32_t any = 0;
            for (int32_t lane = 0; lane < 32; lane++) any |= T.mask[((size_t)t * 4 + rb) * 32 + lane];
            if (any) { T.sub_tile.push_back(t); T.sub_rb.push_back(rb); }
        }
    T.exists = id;
    T.clean.assign(T.n_tiles, 0);
    for (int32_t t = 0; t < T.n_tiles; t++) {
        const int32_t I = T.tile_I[t], B = T.tile_B[t];
        bool ok = B >= 4 * I + 4 && 32 * B + 31 < N && (4 * I + 3) * 32 + 31 < N;
        for (int32_t q = 0; ok && q < 128; q++) ok = T.mask[(size_t)t * 128 + q] == 0xffffffffu;
        T.clean[t] = ok ? 1 : 0;
    }
    T.covers_all_pairs = true;
    for (int32_t I = 0; I < T.n_rb; I++)
        for (int32_t B = 4 * I; B < n_sl; B++)
            if (id[(size_t)I * n_sl + B] < 0) T.covers_all_pairs = false;
    T.eligible = true;
}

// Returns "" on success, otherwise an error message.
inline std::string build_layout(int32_t N, int64_t M, const int64_t *data, const double *cds,
                                bool exact_order, ContactLayout &L, bool want_values = true)
{
    L = ContactLayout();
    L.N = N; L.M = M;
    if (N <= 0) return "n_beads must be positive";
    if (M < 0) return "n_data must be non-negative";
    for (int64_t m = 0; m < M; m++) {
        int64_t i = data[3 * m], j = data[3 * m + 1];
        if (i < 0 || i >= N || j < 0 || j >= N)
            return "data point " + std::to_string(m) + " has a bead index outside [0, n_beads)";
        if (i == j)
            return "data point " + std::to_string(m) + " pairs a bead with itself (distance 0)";
    }
    // (1) pair view: stable sort of the rows by (i, j)
    L.perm.resize(M);
    std::iota(L.perm.begin(), L.perm.end(), (int64_t)0);
    std::stable_sort(L.perm.begin(), L.perm.end(), [&](int64_t a, int64_t b) {
        int64_t ka = data[3 * a] * (int64_t)N + data[3 * a + 1];
        int64_t kb = data[3 * b] * (int64_t)N + data[3 * b + 1];
        return ka < kb;
    });
    L.pi.resize(M); L.pj.resize(M);
    if (want_values) { L.pdc.resize(M); L.pcnt.resize(M); }
    for (int64_t s = 0; s < M; s++) {
        int64_t m = L.perm[s];
        L.pi[s] = (int32_t)data[3 * m];
        L.pj[s] = (int32_t)data[3 * m + 1];
        if (want_values) { L.pdc[s] = cds[m]; L.pcnt[s] = (double)data[3 * m + 2]; }
    }
    // (2) row view
    L.row_len.assign(N, 0);
    for (int64_t s = 0; s < M; s++) { L.row_len[L.pi[s]]++; L.row_len[L.pj[s]]++; }
    L.n_slices = (N + 31) / 32;
    // lane slot -> bead: identity, unless sorting the beads by row length (SELL-C-sigma with
    // sigma = N) saves more than 10 % of the padded slots (ragged lists such as nora2012)
    L.row_bead.assign((size_t)L.n_slices * 32, -1);
    for (int32_t b = 0; b < N; b++) L.row_bead[b] = b;
    auto padded = [&](const std::vector<int32_t> &rb) {
        int64_t tot = 0;
        for (int32_t sl = 0; sl < L.n_slices; sl++) {
            int32_t w = 0;
            for (int32_t l = 0; l < 32; l++) { int32_t b = rb[(size_t)sl * 32 + l]; if (b >= 0) w = std::max(w, L.row_len[b]); }
            tot += (int64_t)32 * w;
        }
        return tot;
    };
    {
        std::vector<int32_t> sorted(L.row_bead);
        std::stable_sort(sorted.begin(), sorted.begin() + N, [&](int32_t a, int32_t b) { return L.row_len[a] > L.row_len[b]; });
        if ((double)padded(sorted) < 0.9 * (double)padded(L.row_bead)) L.row_bead.swap(sorted);
    }
    std::vector<int32_t> slot_of_bead(N);
    for (size_t q = 0; q < L.row_bead.size(); q++) if (L.row_bead[q] >= 0) slot_of_bead[L.row_bead[q]] = (int32_t)q;
    L.slice_off.assign(L.n_slices + 1, 0);
    L.slice_w.assign(L.n_slices, 0);
    for (int32_t sl = 0; sl < L.n_slices; sl++) {
        int32_t w = 0;
        for (int32_t l = 0; l < 32; l++) { int32_t b = L.row_bead[(size_t)sl * 32 + l]; if (b >= 0) w = std::max(w, L.row_len[b]); }
        L.slice_w[sl] = w;
        L.slice_off[sl + 1] = L.slice_off[sl] + (int64_t)32 * w;
    }
    L.tot = L.slice_off[L.n_slices];
    if (L.tot >= ((int64_t)1 << 31)) return "contact list too large for 32-bit row slots";
    // Padding slots (rows shorter than the slice width, lanes past the last bead, tail) point at
    // a real neighbouring bead and carry weight 0 / contact distance 0: the kernel needs no
    // branch, and a zero weight contributes exactly 0 to the gradient.
    L.ent_j.assign(L.tot + ContactLayout::kTailPad, 0);
    for (int32_t sl = 0; sl < L.n_slices; sl++)
        for (int32_t e = 0; e < L.slice_w[sl]; e++)
            for (int32_t lane = 0; lane < 32; lane++) {
                int32_t b = L.row_bead[(size_t)sl * 32 + lane];
                int32_t pad = (b < 0) ? 0 : (b + 1 < N ? b + 1 : (b > 0 ? b - 1 : 0));
                L.ent_j[L.slice_off[sl] + (int64_t)32 * e + lane] = pad;
            }
    if (want_values) L.ent_dc.assign(L.tot + ContactLayout::kTailPad, 0.0);
    L.slot_a.resize(M); L.slot_b.resize(M);
    std::vector<int32_t> fill(N, 0);
    auto place = [&](int32_t b, int32_t o, int64_t s, bool first) {
        const int32_t sb = slot_of_bead[b];
        int64_t pos = L.slice_off[sb / 32] + (int64_t)32 * fill[b] + (sb % 32);
        fill[b]++;
        L.ent_j[pos] = o;
        if (want_values) L.ent_dc[pos] = L.pdc[s];
        (first ? L.slot_a : L.slot_b)[s] = (int32_t)pos;
    };
    if (exact_order) {
        // rows in data-point order: walk the ORIGINAL list; inv maps original row -> sorted position
        std::vector<int64_t> inv(M);
        for (int64_t s = 0; s < M; s++) inv[L.perm[s]] = s;
        for (int64_t m = 0; m < M; m++) {
            int64_t s = inv[m];
            place(L.pi[s], L.pj[s], s, true);
            place(L.pj[s], L.pi[s], s, false);
        }
    } else {
        // walking the (i, j)-sorted list yields, for i < j lists, partners in ascending order
        for (int64_t s = 0; s < M; s++) {
            place(L.pi[s], L.pj[s], s, true);
            place(L.pj[s], L.pi[s], s, false);
        }
    }
    return "";
}

// (4) LANE view (sparse lists, fast arithmetic): the kernels put the 32 structures of a structure
//     group on the 32 lanes of a warp, so everything about a pair is warp-uniform and the list is
//     plain CSR: row_off[b] .. row_off[b+1] are bead b's entries (partner, contact distance), both
//     directions of every pair present; slot_a[m] / slot_b[m] are the CSR positions of sorted pair m
//     in row i and in row j (where the forward kernel drops the pair's weight).
struct LaneLayout {
    std::vector<int64_t> row_off;          // N + 1
    std::vector<int32_t> ent_j;            // 2 M
    std::vector<double> ent_dc;            // 2 M
    std::vector<int32_t> slot_a, slot_b;   // M
};

inline std::string build_lanes(const ContactLayout &L, LaneLayout &C)
{
    C = LaneLayout();
    const int32_t N = L.N; const int64_t M = L.M;
    if (2 * M >= ((int64_t)1 << 31)) return "contact list too large for 32-bit row slots";
    C.row_off.assign((size_t)N + 1, 0);
    for (int32_t b = 0; b < N; b++) C.row_off[b + 1] = C.row_off[b] + L.row_len[b];
    C.ent_j.resize((size_t)2 * M); C.ent_dc.resize((size_t)2 * M);
    C.slot_a.resize(M); C.slot_b.resize(M);
    std::vector<int64_t> fill(C.row_off.begin(), C.row_off.end() - 1);
    const bool have_dc = !L.pdc.empty();
    for (int64_t s = 0; s < M; s++) {
        const int32_t i = L.pi[s], j = L.pj[s];
        const double dc = have_dc ? L.pdc[s] : 0.0;
        int64_t a = fill[i]++, b = fill[j]++;
        C.ent_j[a] = j; C.ent_dc[a] = dc; C.slot_a[s] = (int32_t)a;
        C.ent_j[b] = i; C.ent_dc[b] = dc; C.slot_b[s] = (int32_t)b;
    }
    return "";
}

}  // namespace ehic
