/*
 * ehic_b200.h -- C ABI of libehic_b200.so: the B200 (sm_100a) implementation of
 * ensemble_hic's posterior hot path.
 *
 * The reference has no C ABI / plugin registry for this path: its native
 * entry points are Python-callable Cython / CPython-2 C-API functions.  Each function
 * below names the reference interface it replaces (paths relative to the reference
 * repository).  All functions are exception-free and return an EHIC_* status; the text
 * of the last error of the calling thread is available through ehic_last_error().
 *
 * Conventions
 *   - "d_" pointers are DEVICE pointers owned by the caller; "h_" pointers are HOST
 *     pointers owned by the caller.  Nothing is retained after a call returns except
 *     what ehic_model_create() copies into the model.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device
 *     entry points only enqueue work; host entry points synchronise before returning.
 *   - Coordinates use the reference's flat layout x.reshape(S, N, 3), C order,
 *     structure-major (forward_models.py:78), with a leading replica axis:
 *     x[R][S][N][3], FP64.  Gradients use the same layout.
 *   - A "replica" is one tempered copy of the posterior: (x, norm, lammda, beta).
 *     Every hot-path call is batched over the R replicas resident on this GPU.
 *   - Threading: a model is NOT re-entrant.  Its workspace, staging buffers, streams and captured
 *     trajectory graphs belong to one caller at a time: calls on the same ehic_model_t must be
 *     serialised by the caller (the reference is single-threaded per replica process).
 *     Different models may be used from different host threads concurrently.
 *   - Tuning knobs (EHIC_* environment variables) are read when a model is created,
 *     never on the hot path.
 */
#ifndef EHIC_B200_H
#define EHIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHIC_OK            0
#define EHIC_ERR_INVALID   1   /* bad argument (message in ehic_last_error) */
#define EHIC_ERR_CUDA      2   /* CUDA runtime error */
#define EHIC_ERR_NOMEM     3   /* host or device allocation failed */
#define EHIC_ERR_NODEVICE  4   /* no usable sm_100 device: there is NO CPU fallback */
#define EHIC_ERR_CAPACITY  5   /* caller-provided output buffer too small */

/* terms of the posterior (bit mask) */
#define EHIC_TERM_LIKELIHOOD 1  /* likelihoods.py:41-61 + likelihoods_c.pyx:27-75 */
#define EHIC_TERM_NONBONDED  2  /* nonbonded_prior.py:103-121 + c/forcefield.c, c/prolsq.c */
#define EHIC_TERM_BACKBONE   4  /* backbone_prior.py:112-151 + backbone_prior_c.pyx:11-44 */
#define EHIC_TERM_SPHERE     8  /* sphere_prior.py:85-97 + sphere_prior_c.pyx:12-38 */
#define EHIC_TERM_ALL        15

/* model flags */
#define EHIC_FLAG_EXACT      1  /* IEEE sqrt/div, reference operation order and association,
                                   contact rows in data-point order: the Cython entry points
                                   are reproduced bit for bit (slower; the default "fast" mode
                                   uses Newton-refined rsqrt and is within ~1e-13 relative) */
#define EHIC_FLAG_FP32       2  /* optional FP32 pair arithmetic of the contact term (1e-5 relative), FP64 I/O
                                   and priors; dense contact lists run the FP32 tile kernels, all other lists
                                   the FP32 lane kernels (whole trajectories then run as the multi-kernel graph:
                                   the one-launch trajectory kernels are FP64 only); excludes EHIC_FLAG_EXACT */
#define EHIC_FLAG_NB_CELLS   4  /* nonbonded term through a cell grid instead of tiled all-pairs */

/* indices into the per-replica energy vector written by ehic_evaluate */
#define EHIC_E_LIKELIHOOD 0     /* -log L = sum(md) - sum(c log md)   (untempered) */
#define EHIC_E_NONBONDED  1     /* sum_k E_nb(x_k)                    (untempered) */
#define EHIC_E_BACKBONE   2     /* -log p_backbone */
#define EHIC_E_SPHERE     3     /* -log p_sphere */
#define EHIC_N_ENERGIES   4

/* runtime-settable scalar parameters (the reference's Parameter objects: 'k_bb',
 * 'sphere_radius', 'sphere_k', forcefield.force_constant, smooth_steepness) */
#define EHIC_PARAM_ALPHA          0
#define EHIC_PARAM_NONBONDED_K    1
#define EHIC_PARAM_BACKBONE_K     2
#define EHIC_PARAM_SPHERE_RADIUS  3
#define EHIC_PARAM_SPHERE_K       4
#define EHIC_PARAM_SPHERE_CX      5
#define EHIC_PARAM_SPHERE_CY      6
#define EHIC_PARAM_SPHERE_CZ      7

typedef struct ehic_model ehic_model_t;   /* opaque */

/* Everything setup_functions.make_posterior (setup_functions.py:64-98) feeds into the
 * native kernels.  All pointers are HOST pointers, copied by ehic_model_create. */
typedef struct ehic_model_desc {
    int32_t n_beads;                 /* N  [general] n_beads */
    int32_t n_structures;            /* S  [general] n_structures */
    int64_t n_data;                  /* M  rows of data_points (may be 0) */
    const int64_t *data_points;      /* [M][3] (i, j, count), the reference's Py_ssize_t array
                                        (evaluate_FWM_pureC.pxd:15) as produced by parse_data */
    const double *contact_distances; /* [M]  setup_functions.py:615 */
    double alpha;                    /* smooth_steepness, [forward_model] alpha */
    const double *bead_radii;        /* [N]  setup_functions.py:473-478 */
    double nonbonded_k;              /* [nonbonded_prior] force_constant */
    double backbone_k;               /* [backbone_prior] force_constant ('k_bb') */
    const double *bb_lower;          /* [N-1] lower limit of bond (b, b+1); entries that straddle a
                                        molecule boundary are ignored (NULL = all 0) */
    const double *bb_upper;          /* [N-1] upper limit of bond (b, b+1) (NULL = r_b + r_{b+1},
                                        setup_functions.py:426-428) */
    const int32_t *mol_ranges;       /* [n_mol_ranges] bead offsets of the molecules, first 0, last N
                                        (backbone_prior.py _mol_ranges); NULL = one molecule */
    int32_t n_mol_ranges;
    int32_t sphere_active;           /* [sphere_prior] active == 'True' */
    double sphere_radius, sphere_k;  /* sphere_prior.py:52-56 */
    double sphere_center[3];
    int32_t max_replicas;            /* replicas this model can evaluate per call (workspace size) */
    int32_t flags;                   /* EHIC_FLAG_* */
} ehic_model_desc_t;

/* ---- library ------------------------------------------------------------------------ */
int         ehic_version(void);
const char *ehic_last_error(void);
/* number of CUDA devices with compute capability 10.x; 0 when there is none */
int         ehic_device_count(void);

/* ---- model lifetime (replaces the per-call argument marshalling of
 *      forward_models_c.pyx:14-32 / likelihoods_c.pyx:27-57, and the NBList/PROLSQ
 *      object setup of forcefields.py:166-212) ------------------------------------------ */
int ehic_model_create(const ehic_model_desc_t *desc, int device, ehic_model_t **out);
int ehic_model_destroy(ehic_model_t *m);
int ehic_model_set_param(ehic_model_t *m, int which, double value);
int ehic_model_get_param(const ehic_model_t *m, int which, double *value);
/* dimensions: out[0]=N, out[1]=S, out[2]=M, out[3]=max_replicas, out[4]=flags,
 * out[5]=contact-row slots (padded), out[6]=device */
int ehic_model_info(const ehic_model_t *m, int64_t out[8]);

/* ---- hot path, device buffers ---------------------------------------------------------
 * ehic_mock_data: forward_models_c.ensemble_contacts_evaluate (forward_models_c.pyx:14-32,
 * evaluate_FWM_pureC.pxd:11-41) for R replicas.  d_md[R][M] is in data_points row order. */
int ehic_mock_data(ehic_model_t *m, int R, const double *d_x, const double *d_norm,
                   double *d_md, void *stream);

/* ehic_evaluate: one fused evaluation of the selected posterior terms for R replicas.
 *   d_grad     [R][S][N][3] or NULL: gradient of -log posterior w.r.t. structures
 *              = lammda * calculate_gradient (likelihoods.py:48-61, likelihoods_c.pyx:27-75)
 *              + beta * dE_nb/dx            (nonbonded_prior.py:111-121, c/prolsq.c:24-38)
 *              + backbone_prior_gradient    (backbone_prior_c.pyx:11-44)
 *              + sphere_prior_gradient      (sphere_prior_c.pyx:12-38), restricted to `terms`
 *   d_energies [R][EHIC_N_ENERGIES] or NULL: the UNTEMPERED terms (see EHIC_E_*); the caller
 *              forms log_prob = -lammda*E[0] - beta*E[1] - E[2] - E[3] (+ Gamma prior)
 *   d_md       [R][M] or NULL: mock data by-product
 *   d_lammda, d_beta [R] (NULL = 1.0): tempering parameters (run_simulation.py:89-90) */
int ehic_evaluate(ehic_model_t *m, int R, int terms, const double *d_x, const double *d_norm,
                  const double *d_lammda, const double *d_beta,
                  double *d_grad, double *d_energies, double *d_md, void *stream);

/* ehic_hmc_trajectory: one HMC proposal per replica (binf HMCSampler / csb FastLeapFrog
 * as driven by setup_functions.py:239-254): half kick, (L-1) x (drift, kick), drift,
 * half kick = L+1 gradient evaluations, positions and momenta resident on the device.
 *   d_x, d_p   [R][S][N][3] in/out: overwritten with the END point of the trajectory
 *   d_eps      [R] leapfrog timesteps
 *   d_H        [R][4] out: (E0, K0, E1, K1), E = -log posterior(structures | norm) at the
 *              replica's (lammda, beta) WITHOUT the norm prior, K = |p|^2 / 2
 * Accept/reject is the caller's (it owns the RNG stream); see ehic_select. */
int ehic_hmc_trajectory(ehic_model_t *m, int R, int n_steps, double *d_x, double *d_p,
                        const double *d_norm, const double *d_lammda, const double *d_beta,
                        const double *d_eps, double *d_H, void *stream);

/* How ehic_hmc_trajectory runs a trajectory of this model (north-star kernel 3: "one launch produces one HMC
 * proposal"): out[0] = 0 multi-kernel step sequence replayed as one CUDA graph, 1 one kernel launch with one CTA
 * per replica (state in shared memory), 2 one kernel launch with a thread-block cluster per replica (state split
 * over the cluster's shared memories); out[1] = CTAs per replica; out[2] = dynamic shared memory per CTA in
 * bytes; out[3] = kernel launches per trajectory (0 = depends on the step sequence). */
int ehic_trajectory_plan(const ehic_model_t *m, int R, int n_steps, int64_t out[4]);

/* d_dst[r] = d_src[r] for replicas with d_mask[r] != 0 (row length n doubles) */
int ehic_select(int R, int64_t n, const int32_t *d_mask, const double *d_src, double *d_dst, void *stream);

/* ---- neighbour list (c/nblist.c:181-303 nblist_update; nblist.py:51-83) -----------------
 * PARITY SURFACE of the NBList drop-in (one structure per call, O(N^2) membership sweep, allocates
 * and copies inside the call): it exists so that list membership can be compared bit for bit with
 * the reference's.  The production nonbonded path never materialises this list -- ehic_evaluate /
 * ehic_hmc_trajectory use the pruned all-pairs, cell-grid and Verlet-list kernels instead.
 * Half list of one structure: all i<j with |x_i - x_j|^2 <= cutoff^2, dot product
 * associated as mathutils.c:13-15, rows sorted (canonical order).
 * h_pairs[capacity][2]; *n_pairs receives the total count (returned by NBList.update). */
int ehic_nblist_host(int n_beads, const double *h_coords, double cutoff,
                     int32_t *h_pairs, int64_t capacity, int64_t *n_pairs, int device);

/* ---- hot path, host buffers (what the Python drop-in functions call; copies inside) ---- */
int ehic_mock_data_host(ehic_model_t *m, int R, const double *h_x, const double *h_norm, double *h_md);
int ehic_evaluate_host(ehic_model_t *m, int R, int terms, const double *h_x, const double *h_norm,
                       const double *h_lammda, const double *h_beta,
                       double *h_grad, double *h_energies, double *h_md);

/* ---- contact-list layout (host only, no GPU needed): the canonical order the kernels use.
 * h_sorted[M] receives, for each position of the (i, j)-sorted list, the row index into
 * data_points ("contact lists bit-exact": data_points[h_sorted] is the list the GPU walks). */
int ehic_layout_debug(int32_t n_beads, int64_t n_data, const int64_t *data_points, int exact_order,
                      int64_t *h_sorted, int64_t *n_slots, int32_t *h_row_len /* [N] */);

/* ---- per-kernel timing: with profiling enabled every kernel launch of this model is bracketed
 * by CUDA events on the stream it is launched on; ehic_profile_read synchronises the device and
 * returns (and resets) the summed milliseconds and launch counts per kernel class:
 * 0 pack, 1 forward, 2 prior, 3 gather, 4 finalize, 5 other. */
int ehic_profile(ehic_model_t *m, int enable);
int ehic_profile_read(ehic_model_t *m, double ms_out[8], int64_t count_out[8]);

/* ---- measurement helpers ---------------------------------------------------------------
 * number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t ehic_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EHIC_B200_H */
