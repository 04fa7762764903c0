/*
 * ehic_oracle.c -- CPU restatement of ensemble_hic's posterior hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ensemble_hic_b200/) may
 * link, load or call this file; it is the checker used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Every function restates, in plain C (FP64, no FMA contraction: build with
 * -ffp-contract=off), the algorithm of the reference file:line it cites (paths
 * relative to /root/reference).  Pinning:
 *   - the Cython-backed functions (fwm, likelihood gradient, O(N^2) forcefield,
 *     backbone, sphere) are checked bit-for-bit against the reference's own
 *     compiled .pyx (oracle/_ref, built by oracle/build_ref.py) and against the
 *     reference's known-answer tests (tests/cases/forward_models.py,
 *     tests/cases/backbone_prior.py) in tests/test_oracle.py;
 *   - the neighbour list + PROLSQ part restates ensemble_hic/c/{nblist,forcefield,
 *     prolsq,mathutils}.c.  Those sources are a Python-2 C-API extension that does not
 *     build as an extension of CPython 3, but oracle/build_ref.build_nbl compiles them
 *     UNMODIFIED as plain C against the stand-in headers in oracle/shim/ (oracle/_ref/
 *     libref_nbl.so, interface oracle/ref_nbl_driver.c); this restatement equals them bit
 *     for bit -- same contacts in the same list order, same energy and gradient bits --
 *     and reproduces the golden vectors generated from them (tests/golden/nbl_reference.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* evaluate_FWM_pureC.pxd:11-41 -- smooth-contact ensemble forward model.     */
/* x[S][N][3]; data[M][3] = (i, j, count) int64; res[M] must be zeroed by the */
/* caller (forward_models_c.pyx:27).  distances/sqrtdenoms are [S][M] scratch.*/
EXPORT int oracle_fwm(int64_t S, int64_t N, int64_t M, const double *x,
                      double norm, const double *cds, double alpha,
                      const int64_t *data, double *distances,
                      double *sqrtdenoms, double *res)
{
    for (int64_t k = 0; k < S; k++) {
        for (int64_t m = 0; m < M; m++) {
            int64_t i = data[3 * m], j = data[3 * m + 1];
            double d = 0.0;
            for (int l = 0; l < 3; l++) {
                double a = x[(k * N + i) * 3 + l] - x[(k * N + j) * 3 + l];
                d += a * a;
            }
            double dist = sqrt(d);
            distances[k * M + m] = dist;
            double u = cds[m] - dist;
            double q = sqrt(1.0 + alpha * alpha * u * u);
            sqrtdenoms[k * M + m] = q;
            res[m] += (alpha * u / q + 1.0) * 0.5 * norm;
        }
    }
    return 1;
}

/* likelihoods_c.pyx:27-75 -- d(-log L)/dx for the Poisson error model        */
/* (poisson_derivative, pyx:24-25).  result[S*N*3] structure-major, zeroed    */
/* here; md[M] is returned as a by-product (zeroed here too).                 */
EXPORT int oracle_likelihood_gradient(int64_t S, int64_t N, int64_t M,
                                      const double *x, double alpha, double norm,
                                      const double *cds, const int64_t *data,
                                      double *result, double *md)
{
    double *distances = (double *)malloc(sizeof(double) * (size_t)(S * M + 1));
    double *sqrtdenoms = (double *)malloc(sizeof(double) * (size_t)(S * M + 1));
    if (!distances || !sqrtdenoms) { free(distances); free(sqrtdenoms); return -1; }
    memset(result, 0, sizeof(double) * (size_t)(S * N * 3));
    memset(md, 0, sizeof(double) * (size_t)M);
    oracle_fwm(S, N, M, x, norm, cds, alpha, data, distances, sqrtdenoms, md);
    for (int64_t k = 0; k < S; k++) {
        int64_t offset = k * N * 3;
        for (int64_t u = 0; u < M; u++) {
            int64_t i = data[3 * u], j = data[3 * u + 1];
            double d = distances[k * M + u];
            double g = 1.0 + alpha * alpha * (cds[u] - d) * (cds[u] - d);
            double em = 1.0 - (double)data[3 * u + 2] / md[u];
            double f = 0.5 * em * norm * alpha / (g * sqrtdenoms[k * M + u] * d);
            for (int l = 0; l < 3; l++) {
                double value = (x[offset + j * 3 + l] - x[offset + i * 3 + l]) * f;
                result[offset + i * 3 + l] += value;
                result[offset + j * 3 + l] -= value;
            }
        }
    }
    free(distances);
    free(sqrtdenoms);
    return 0;
}

/* forcefield_c.pyx:11-31 -- O(N^2) quartic repulsion energy, one structure.  */
EXPORT double oracle_ff_energy(int64_t N, const double *s, const double *radii,
                               const double *radii2, double k)
{
    double res = 0.0;
    for (int64_t i = 0; i < N; i++) {
        for (int64_t j = i + 1; j < N; j++) {
            double d = (s[3*i] - s[3*j]) * (s[3*i] - s[3*j])
                     + (s[3*i+1] - s[3*j+1]) * (s[3*i+1] - s[3*j+1])
                     + (s[3*i+2] - s[3*j+2]) * (s[3*i+2] - s[3*j+2]);
            if (d < radii2[i] + radii2[j] + 2 * radii[i] * radii[j]) {
                d = sqrt(d);
                double a = d - radii[i] - radii[j];
                res += pow(a, 4.0);   /* Cython lowers "** 4" on a C double to pow(a, 4.0) */
            }
        }
    }
    return 0.5 * k * res;
}

/* forcefield_c.pyx:33-59 -- O(N^2) gradient, one structure; out[3N].         */
EXPORT int oracle_ff_gradient(int64_t N, const double *s, const double *radii,
                              const double *radii2, double k, double *out)
{
    memset(out, 0, sizeof(double) * (size_t)(3 * N));
    for (int64_t i = 0; i < N; i++) {
        for (int64_t j = i + 1; j < N; j++) {
            double d = 0.0;
            for (int l = 0; l < 3; l++)
                d += (s[3*i+l] - s[3*j+l]) * (s[3*i+l] - s[3*j+l]);
            if (d < radii2[i] + radii2[j] + 2 * radii[i] * radii[j]) {
                d = sqrt(d);
                double a = radii[i] + radii[j] - d;
                double g = a * a * a / d;
                for (int l = 0; l < 3; l++) {
                    out[3*i+l] -= (s[3*i+l] - s[3*j+l]) * g;
                    out[3*j+l] += (s[3*i+l] - s[3*j+l]) * g;
                }
            }
        }
    }
    for (int64_t t = 0; t < 3 * N; t++) out[t] = k * 2.0 * out[t];
    return 0;
}

/* backbone_prior_c.pyx:11-44 -- one molecule of n beads; ll/ul[n-1]; out[3n] */
EXPORT int oracle_backbone_gradient(int64_t n, const double *s, const double *ll,
                                    const double *ul, double k, double *out)
{
    memset(out, 0, sizeof(double) * (size_t)(3 * n));
    for (int64_t i = 1; i < n; i++) {
        double d = 0.0;
        for (int j = 0; j < 3; j++)
            d += (s[3*i+j] - s[3*(i-1)+j]) * (s[3*i+j] - s[3*(i-1)+j]);
        if (d > ul[i-1] * ul[i-1]) {
            d = sqrt(d);
            for (int j = 0; j < 3; j++) {
                double e = (d - ul[i-1]) * (s[3*i+j] - s[3*(i-1)+j]) / d;
                out[3*i+j] += e;
                out[3*(i-1)+j] -= e;
            }
        } else if (d < ll[i-1] * ll[i-1]) {
            d = sqrt(d);
            for (int j = 0; j < 3; j++) {
                double e = (ll[i-1] - d) * (s[3*i+j] - s[3*(i-1)+j]) / d;
                out[3*i+j] -= e;
                out[3*(i-1)+j] += e;
            }
        }
    }
    for (int64_t t = 0; t < 3 * n; t++) out[t] = k * out[t];
    return 0;
}

/* sphere_prior_c.pyx:12-38 -- one structure; beads == arange(N) (quirk B-5). */
EXPORT int oracle_sphere_gradient(int64_t N, const double *s, const double *center,
                                  double radius, double k, const double *bead_radii,
                                  const double *bead_radii2, double *out)
{
    double radius2 = radius * radius;
    memset(out, 0, sizeof(double) * (size_t)(3 * N));
    for (int64_t i = 0; i < N; i++) {
        double d = 0.0;
        for (int j = 0; j < 3; j++)
            d += (s[3*i+j] - center[j]) * (s[3*i+j] - center[j]);
        if (d < radius2 + bead_radii2[i] - 2 * radius * bead_radii[i]) continue;
        d = sqrt(d);
        double f = (d + bead_radii[i] - radius) / d;
        for (int j = 0; j < 3; j++) out[3*i+j] += (s[3*i+j] - center[j]) * f;
    }
    for (int64_t t = 0; t < 3 * N; t++) out[t] = k * out[t];
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Neighbour list: c/nblist.c + c/nblist.h + c/mathutils.c:13-24.             */

#define MAX_NO_NEIGHBORS 13                          /* nblist.h:7  */
#define NB_INDEX(i, j, k, n) ((n) * ((n) * (i) + (j)) + (k))  /* nblist.h:10 */
#define BOX_MARGIN 1e-5                              /* nblist.h:12 */

typedef struct { int id; int n_objects; int *objects; } ocell_t;   /* nblist.h:14-20 */

typedef struct oracle_nblist {
    int n_cells, n_atoms, n_per_cell, n_filled, max_n_contacts;
    double cellsize, origin;
    int *n_contacts;
    int *contacts;          /* [n_atoms][max_n_contacts], -1 padded like the getter nblist.c:452-462 */
    double *sq_distances;   /* [n_atoms][max_n_contacts] */
    ocell_t **cells;        /* (n_cells+2)^3 pointers */
    ocell_t *filled;        /* n_atoms cells */
    int neighbors[MAX_NO_NEIGHBORS];
    int n_dropped;          /* atoms dropped by the "no space left in cell" branch, nblist.c:166-169 */
} oracle_nblist_t;

static double vector_dot3(const double *a, const double *b)
{   /* mathutils.c:13-15: (a0*b0 + a1*b1) + a2*b2 */
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* nblist.c:9-29 set_neighbors */
static void nb_set_neighbors(oracle_nblist_t *nb)
{
    int j, k, n = nb->n_cells, counter = 1;
    nb->neighbors[0] = NB_INDEX(0, 0, 1, n);
    for (k = -1; k < 2; k++, counter++) nb->neighbors[counter] = NB_INDEX(0, 1, k, n);
    for (j = -1; j < 2; j++) for (k = -1; k < 2; k++) {
        nb->neighbors[counter] = NB_INDEX(1, j, k, n);
        counter++;
    }
}

EXPORT void oracle_nblist_free(oracle_nblist_t *nb)
{
    if (!nb) return;
    if (nb->filled) {
        for (int i = 0; i < nb->n_atoms; i++) free(nb->filled[i].objects);
        free(nb->filled);
    }
    free(nb->cells); free(nb->n_contacts); free(nb->contacts); free(nb->sq_distances);
    free(nb);
}

/* Construction order of nblist.py:51-83 / forcefields.py:175: cellsize, n_per_cell,
 * n_atoms (set_natoms, nblist.c:369-412: max_n = min(n_per_cell*14, n)), n_cells
 * (init_cells nblist.c:31-61 + set_neighbors). */
EXPORT oracle_nblist_t *oracle_nblist_new(double cellsize, int n_cells, int n_per_cell, int n_atoms)
{
    oracle_nblist_t *nb = (oracle_nblist_t *)calloc(1, sizeof(*nb));
    if (!nb) return NULL;
    nb->cellsize = cellsize; nb->n_cells = n_cells;
    nb->n_per_cell = n_per_cell; nb->n_atoms = n_atoms;
    int max_n = n_per_cell * (MAX_NO_NEIGHBORS + 1);
    if (max_n > n_atoms) max_n = n_atoms;
    nb->max_n_contacts = max_n;
    size_t grid = (size_t)(n_cells + 2) * (n_cells + 2) * (n_cells + 2);
    nb->n_contacts = (int *)calloc((size_t)n_atoms, sizeof(int));
    nb->contacts = (int *)malloc(sizeof(int) * (size_t)n_atoms * max_n);
    nb->sq_distances = (double *)malloc(sizeof(double) * (size_t)n_atoms * max_n);
    nb->cells = (ocell_t **)calloc(grid, sizeof(ocell_t *));
    nb->filled = (ocell_t *)calloc((size_t)n_atoms, sizeof(ocell_t));
    if (!nb->n_contacts || !nb->contacts || !nb->sq_distances || !nb->cells || !nb->filled) {
        oracle_nblist_free(nb); return NULL;
    }
    for (int i = 0; i < n_atoms; i++) {
        nb->filled[i].objects = (int *)malloc(sizeof(int) * (size_t)n_per_cell);
        nb->filled[i].n_objects = 0;
    }
    nb_set_neighbors(nb);
    return nb;
}

EXPORT void oracle_nblist_set_cellsize(oracle_nblist_t *nb, double c) { nb->cellsize = c; }

/* nblist.c:88-110 update_bbox over all 3n scalars (cubic box, quirk B-4) */
EXPORT double oracle_nblist_update_bbox(oracle_nblist_t *nb, const double *coords, int n)
{
    double x, x_min = coords[0], x_max = coords[0];
    for (int i = 1; i < n; i++) {
        x = coords[i];
        if (x < x_min) x_min = x;
        else if (x > x_max) x_max = x;
    }
    nb->origin = x_min - BOX_MARGIN;
    return x_max + BOX_MARGIN - x_min;
}

/* nblist.c:112-179 assign_atoms */
static int nb_assign_atoms(oracle_nblist_t *nb, const double *coords, int n_coords, int new_box)
{
    int n_filled = 0;
    double cellsize = nb->cellsize;
    int n_cells = nb->n_cells;
    if (new_box) {
        double size = oracle_nblist_update_bbox(nb, coords, 3 * n_coords);
        if ((int)floor(size / cellsize) >= n_cells) cellsize = size / (n_cells - 1. + 1.e-8);
    }
    for (int n = 0; n < n_coords; n++) {
        nb->n_contacts[n] = 0;
        int i = (int)floor((coords[3*n]   - nb->origin) / cellsize);
        int j = (int)floor((coords[3*n+1] - nb->origin) / cellsize);
        int k = (int)floor((coords[3*n+2] - nb->origin) / cellsize);
        int index = NB_INDEX(i + 1, j + 1, k + 1, n_cells);
        if (!nb->cells[index]) {
            nb->cells[index] = &nb->filled[n_filled];
            nb->cells[index]->id = index;
            n_filled++;
        }
        ocell_t *cur = nb->cells[index];
        if (cur->n_objects >= nb->n_per_cell) { nb->n_dropped++; continue; }
        cur->objects[cur->n_objects] = n;
        cur->n_objects++;
    }
    nb->n_filled = n_filled;
    return 0;
}

/* nblist.c:181-303 nblist_update: half list, membership |dx|^2 <= cellsize^2 with
 * self->cellsize (quirk B-2), list order as produced by the cell sweep. */
EXPORT int oracle_nblist_update(oracle_nblist_t *nb, const double *coords, int n_coords, int new_box)
{
    int total = 0;
    double cellsize2 = nb->cellsize * nb->cellsize;
    int mx = nb->max_n_contacts;
    nb->n_dropped = 0;
    if (nb_assign_atoms(nb, coords, n_coords, new_box)) return -1;
    for (int n = 0; n < nb->n_filled; n++) {
        ocell_t *cur = &nb->filled[n];
        for (int i = 0; i < cur->n_objects; i++) {
            int atom = cur->objects[i];
            int *ac = nb->contacts + (size_t)atom * mx;
            double *sq = nb->sq_distances + (size_t)atom * mx;
            for (int j = i + 1; j < cur->n_objects; j++) {
                int partner = cur->objects[j];
                double dx[3];
                for (int l = 0; l < 3; l++) dx[l] = coords[3*atom+l] - coords[3*partner+l];
                double d2 = vector_dot3(dx, dx);
                if (d2 > cellsize2) continue;
                ac[nb->n_contacts[atom]] = partner;
                sq[nb->n_contacts[atom]] = d2;
                nb->n_contacts[atom]++; total++;
            }
            for (int j = 0; j < MAX_NO_NEIGHBORS; j++) {
                ocell_t *nbr = nb->cells[cur->id + nb->neighbors[j]];
                if (!nbr) continue;
                for (int k = 0; k < nbr->n_objects; k++) {
                    int partner = nbr->objects[k];
                    double dx[3];
                    for (int l = 0; l < 3; l++) dx[l] = coords[3*atom+l] - coords[3*partner+l];
                    double d2 = vector_dot3(dx, dx);
                    if (d2 > cellsize2) continue;
                    ac[nb->n_contacts[atom]] = partner;
                    sq[nb->n_contacts[atom]] = d2;
                    nb->n_contacts[atom]++; total++;
                }
            }
        }
    }
    for (int i = 0; i < nb->n_filled; i++) {
        nb->cells[nb->filled[i].id] = NULL;
        nb->filled[i].n_objects = 0;
    }
    return total;
}

EXPORT int oracle_nblist_max_contacts(const oracle_nblist_t *nb) { return nb->max_n_contacts; }
EXPORT int oracle_nblist_n_dropped(const oracle_nblist_t *nb) { return nb->n_dropped; }
EXPORT int oracle_nblist_n_filled(const oracle_nblist_t *nb) { return nb->n_filled; }
EXPORT const int *oracle_nblist_n_contacts(const oracle_nblist_t *nb) { return nb->n_contacts; }
EXPORT const int *oracle_nblist_contacts(const oracle_nblist_t *nb) { return nb->contacts; }
EXPORT const double *oracle_nblist_sq_distances(const oracle_nblist_t *nb) { return nb->sq_distances; }

/* Flatten the half list to (i, j) rows in list order; returns the number written. */
EXPORT int oracle_nblist_pairs(const oracle_nblist_t *nb, int32_t *pairs, int capacity)
{
    int w = 0;
    for (int i = 0; i < nb->n_atoms; i++)
        for (int n = 0; n < nb->n_contacts[i]; n++) {
            if (w < capacity) { pairs[2*w] = i; pairs[2*w+1] = nb->contacts[(size_t)i * nb->max_n_contacts + n]; }
            w++;
        }
    return w;
}

/* prolsq.c:8-22 prolsq_energy */
static double prolsq_energy(double d, double d0, double k)
{
    double a = d0 - d;
    if (a < 0.) return 0.;
    a *= a; a *= a;
    return k * a / 2.;
}

/* prolsq.c:24-38 prolsq_gradient (E accumulates k*(d0-d)^4, NOT halved: quirk B-8) */
static double prolsq_gradient(double d, double d0, double k, double *E)
{
    double c = 0.;
    if (d < d0) {
        c = (d0 - d);
        c *= k * c * c;
        *E += (d0 - d) * c;
        c *= -2 / d;
    }
    return c;
}

/* forcefield.c:8-53 forcefield_energy over the current list, with the INTENDED
 * pair parameters d0 = d[i][j], k = k[i][j] of forcefields.py:181-212
 * (d = r_i + r_j, k = force constant); K = 1.  See quirk B-1 for the int32/int64
 * `types` aliasing of the reference, which is equivalent for uniform radii. */
EXPORT double oracle_nbl_energy(const oracle_nblist_t *nb, const double *coords,
                                const double *radii, double kconst, int n_particles)
{
    double E = 0.;
    for (int i = 0; i < n_particles; i++) {
        const int *ci = nb->contacts + (size_t)i * nb->max_n_contacts;
        for (int n = 0; n < nb->n_contacts[i]; n++) {
            int j = ci[n];
            double dx[3];
            for (int l = 0; l < 3; l++) dx[l] = coords[3*i+l] - coords[3*j+l];
            double r = sqrt(vector_dot3(dx, dx));
            E += prolsq_energy(r, radii[i] + radii[j], kconst);
        }
    }
    E *= 1.0;   /* self->K, forcefield.c:50 */
    return E;
}

/* forcefield.c:55-109 forcefield_gradient: accumulates INTO forces (caller zeroes,
 * forcefields.py:220); returns the (unhalved) E through *E_ptr. */
EXPORT int oracle_nbl_gradient(const oracle_nblist_t *nb, const double *coords, double *forces,
                               const double *radii, double kconst, int n_particles, double *E_ptr)
{
    double E = 0.;
    for (int i = 0; i < n_particles; i++) {
        const int *ci = nb->contacts + (size_t)i * nb->max_n_contacts;
        for (int n = 0; n < nb->n_contacts[i]; n++) {
            int j = ci[n];
            double dx[3];
            for (int l = 0; l < 3; l++) dx[l] = coords[3*i+l] - coords[3*j+l];
            double r = sqrt(vector_dot3(dx, dx));
            double c = prolsq_gradient(r, radii[i] + radii[j], kconst, &E);
            for (int l = 0; l < 3; l++) {
                forces[3*i+l] += c * dx[l];
                forces[3*j+l] -= c * dx[l];
            }
        }
    }
    if (E_ptr) *E_ptr = E;
    return 0;
}
