/* oracle/_ref/libref_nbl.so -- the REFERENCE'S OWN neighbour list and PROLSQ forcefield (ensemble_hic/c/nblist.c,
 * forcefield.c, prolsq.c, mathutils.c, compiled unmodified from where they lie against the shim headers in
 * oracle/shim/) behind a plain C interface.  Test infrastructure: it pins the C restatement in ehic_oracle.c, which in
 * turn checks the CUDA path (tests/test_oracle.py); nothing under ensemble_hic_b200/ loads it.
 *
 * The objects are created and configured exactly the way the reference's Python layer does it -- through the
 * constructors PyNBList_nblist / PyProlsq_New and the types' own tp_setattr slots:
 *   NBList(cellsize, n_cells, n_per_cell, n_particles)        nblist.py:51-83   (attribute order: cellsize, n_per_cell,
 *                                                             n_atoms, n_cells last, then enabled = 1)
 *   NBLForceField._make_isd_ff / bead_radii / force_constant   forcefields.py:166-212 (n_types = n_beads, types = arange,
 *                                                             NBList(1.0 + 1e-3, 90, n_beads, n_beads), d = r_i + r_j,
 *                                                             cellsize = max(d) + 1e-3, k = force constant)
 *   energy  = update_list + ctype.energy                       isd_forcefield.py: Forcefield.energy
 *   gradient = update_list + ctype.update_gradient(.., 1)      forcefields.py:214-224 */
#include <Python.h>
#include <numpy/arrayobject.h>
#include "ensemble_hic.h"

PyObject shim_none, shim_exc;
char shim_last_error[256];
PyTypeObject PyArray_Type;

int nblist_update(PyNBListObject *self, vector *coords, int n_coords, int new_box);   /* c/nblist.c:181 */

static int set_int(PyObject *o, char *name, long v)
{
    PyObject op; memset(&op, 0, sizeof(op)); op.shim_ival = v; op.shim_dval = (double)v;
    return o->ob_type->tp_setattr(o, name, &op);
}
static int set_double(PyObject *o, char *name, double v)
{
    PyObject op; memset(&op, 0, sizeof(op)); op.shim_dval = v; op.shim_ival = (long)v;
    return o->ob_type->tp_setattr(o, name, &op);
}
static int set_matrix(PyObject *o, char *name, double *data, int n)
{
    PyArrayObject a; npy_intp dims[2] = {n, n}, strides[2] = {(npy_intp)(n * sizeof(double)), (npy_intp)sizeof(double)};
    memset(&a, 0, sizeof(a));
    a.ob_type = &PyArray_Type; a.data = (char *)data; a.nd = 2; a.dimensions = dims; a.strides = strides;
    return o->ob_type->tp_setattr(o, name, (PyObject *)&a);
}

void *ref_nblist_new(double cellsize, int n_cells, int n_per_cell, int n_atoms)
{
    PyObject *o = PyNBList_nblist(NULL, NULL);
    if (!o) return NULL;
    int bad = set_double(o, "cellsize", cellsize) | set_int(o, "n_per_cell", n_per_cell) | set_int(o, "n_atoms", n_atoms) |
              set_int(o, "n_cells", n_cells) | set_int(o, "enabled", 1);
    if (bad) { o->ob_type->tp_dealloc(o); return NULL; }
    return o;
}
void ref_nblist_free(void *h) { if (h) ((PyObject *)h)->ob_type->tp_dealloc((PyObject *)h); }
void ref_nblist_set_cellsize(void *h, double c) { set_double((PyObject *)h, "cellsize", c); }
int ref_nblist_update(void *h, const double *coords, int n, int new_box)
{
    return nblist_update((PyNBListObject *)h, (vector *)coords, n, new_box);
}
/* half list in the reference's own list order: rows (i, contacts[i][q]); returns the number of pairs */
int ref_nblist_pairs(void *h, int *out, int capacity)
{
    PyNBListObject *nb = (PyNBListObject *)h;
    int n = 0;
    for (int i = 0; i < nb->n_atoms; i++)
        for (int q = 0; q < nb->n_contacts[i]; q++) {
            if (out && n < capacity) { out[2 * n] = i; out[2 * n + 1] = nb->contacts[i][q]; }
            n++;
        }
    return n;
}

typedef struct { PyObject *ff, *nb; int *types; int n; } RefFF;

void *ref_nblff_new(int n_beads, const double *radii, double force_constant)
{
    RefFF *f = (RefFF *)calloc(1, sizeof(RefFF));
    if (!f) return NULL;
    f->n = n_beads;
    f->ff = PyProlsq_New(NULL, NULL);
    f->nb = (PyObject *)ref_nblist_new(1.0 + 1e-3, 90, n_beads, n_beads);
    f->types = (int *)malloc(sizeof(int) * n_beads);
    double *d = (double *)malloc(sizeof(double) * n_beads * n_beads), *k = (double *)malloc(sizeof(double) * n_beads * n_beads);
    if (!f->ff || !f->nb || !f->types || !d || !k) return NULL;
    double dmax = 0.0;
    for (int i = 0; i < n_beads; i++) {
        f->types[i] = i;
        for (int j = 0; j < n_beads; j++) {
            d[i * n_beads + j] = radii[i] + radii[j]; k[i * n_beads + j] = force_constant;
            if (d[i * n_beads + j] > dmax) dmax = d[i * n_beads + j];
        }
    }
    set_int(f->ff, "n_types", n_beads);
    f->ff->ob_type->tp_setattr(f->ff, "nblist", f->nb);
    set_matrix(f->ff, "d", d, n_beads);
    set_matrix(f->ff, "k", k, n_beads);
    ref_nblist_set_cellsize(f->nb, dmax + 1e-3);
    free(d); free(k);
    return f;
}
void ref_nblff_free(void *h)
{
    RefFF *f = (RefFF *)h;
    if (!f) return;
    if (f->ff) f->ff->ob_type->tp_dealloc(f->ff);
    if (f->nb) f->nb->ob_type->tp_dealloc(f->nb);
    free(f->types); free(f);
}
double ref_nblff_energy(void *h, const double *coords)
{
    RefFF *f = (RefFF *)h;
    nblist_update((PyNBListObject *)f->nb, (vector *)coords, f->n, 1);
    return forcefield_energy((PyForceFieldObject *)f->ff, (double *)coords, f->types, f->n);
}
/* forces must be zero-initialised by the caller (forcefields.py:219); returns the energy accumulated by grad_f */
double ref_nblff_gradient(void *h, const double *coords, double *forces)
{
    RefFF *f = (RefFF *)h;
    double E = 0.0;
    nblist_update((PyNBListObject *)f->nb, (vector *)coords, f->n, 1);
    forcefield_gradient((PyForceFieldObject *)f->ff, (double *)coords, forces, f->types, f->n, &E);
    return E;
}
const char *ref_last_error(void) { return shim_last_error; }
