"""CPU oracle for the ensemble_hic posterior hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  The product package
(``ensemble_hic_b200``) never does, and has no CPU fallback.

Two layers:

* :class:`COracle` -- ctypes binding of ``libehic_oracle.so`` (``ehic_oracle.c``), the
  plain-C restatement of the reference's native kernels (each function cites the
  reference file:line it follows).
* numpy restatements of the log-probabilities the reference evaluates in numpy
  (backbone_prior.py:64-90, sphere_prior.py:62-72, error_models.py:85-97,
  gamma_prior.py:48-53) and :class:`OraclePosterior`, which assembles them exactly
  like setup_functions.make_posterior does (tempering: likelihoods.py:41-61,
  nonbonded_prior.py:103-121,158-164).

:func:`load_ref` imports the reference's own compiled Cython modules from
``oracle/_ref`` (built by ``oracle/build_ref.py``) when present; :class:`RefNBL` loads the
reference's own neighbour-list / PROLSQ C sources compiled as plain C (``build_ref.build_nbl``).

The HMC restatement (:func:`leapfrog`, :func:`hmc_step`) follows csb 1.2.5's
``FastLeapFrog`` / binf's ``HMCSampler`` as recalled from their public sources; those
packages are absent from /root/reference, and the reference has no test for them:
**parity unpinned** for that part.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "libehic_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build_c_oracle(force=False):
    """gcc the plain-C restatement (seconds)."""
    src = os.path.join(HERE, "ehic_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-fvisibility=hidden", src, "-o", _LIB, "-lm"])
    return _LIB


def load_ref():
    """Return a dict of the reference's own compiled Cython functions, or None."""
    ref_dir = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "ensemble_hic")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        from ensemble_hic.forward_models_c import ensemble_contacts_evaluate
        from ensemble_hic.likelihoods_c import calculate_gradient
        from ensemble_hic.forcefield_c import forcefield_energy, forcefield_gradient
        from ensemble_hic.backbone_prior_c import backbone_prior_gradient
        from ensemble_hic.sphere_prior_c import sphere_prior_gradient
    except ImportError:
        return None
    return dict(ensemble_contacts_evaluate=ensemble_contacts_evaluate,
                calculate_gradient=calculate_gradient,
                forcefield_energy=forcefield_energy,
                forcefield_gradient=forcefield_gradient,
                backbone_prior_gradient=backbone_prior_gradient,
                sphere_prior_gradient=sphere_prior_gradient)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(_dp)


class COracle(object):
    """ctypes face of ehic_oracle.c, same call signatures as the reference's native entries."""

    def __init__(self):
        build_c_oracle()
        L = C.CDLL(_LIB)
        L.oracle_fwm.argtypes = [C.c_int64] * 3 + [_dp, C.c_double, _dp, C.c_double, _ip, _dp, _dp, _dp]
        L.oracle_likelihood_gradient.argtypes = [C.c_int64] * 3 + [_dp, C.c_double, C.c_double, _dp, _ip, _dp, _dp]
        L.oracle_ff_energy.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double]
        L.oracle_ff_energy.restype = C.c_double
        L.oracle_ff_gradient.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, _dp]
        L.oracle_backbone_gradient.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, _dp]
        L.oracle_sphere_gradient.argtypes = [C.c_int64, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
        L.oracle_nblist_new.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        L.oracle_nblist_new.restype = C.c_void_p
        L.oracle_nblist_free.argtypes = [C.c_void_p]
        L.oracle_nblist_set_cellsize.argtypes = [C.c_void_p, C.c_double]
        L.oracle_nblist_update.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int]
        L.oracle_nblist_update_bbox.argtypes = [C.c_void_p, _dp, C.c_int]
        L.oracle_nblist_update_bbox.restype = C.c_double
        L.oracle_nblist_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.oracle_nblist_n_dropped.argtypes = [C.c_void_p]
        L.oracle_nblist_n_filled.argtypes = [C.c_void_p]
        L.oracle_nbl_energy.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int]
        L.oracle_nbl_energy.restype = C.c_double
        L.oracle_nbl_gradient.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int, _dp]
        self.L = L

    # forward_models_c.pyx:14-32 (argument order: norm, cds, alpha, data_points)
    def ensemble_contacts_evaluate(self, structures, norm, contact_distances, alpha, data_points):
        X = _d(structures); cds = _d(contact_distances); dp = _i(data_points)
        S, N = X.shape[0], X.shape[1]; M = dp.shape[0]
        dist = np.empty((S, max(M, 1))); sq = np.empty((S, max(M, 1))); res = np.zeros(M)
        self.L.oracle_fwm(S, N, M, _p(X), float(norm), _p(cds), float(alpha),
                          dp.ctypes.data_as(_ip), _p(dist), _p(sq), _p(res))
        return res

    # likelihoods_c.pyx:27-75 (argument order: structures, alpha, norm, cds, data_points)
    def calculate_gradient(self, structures, alpha, norm, cds, data_points):
        X = _d(structures); cds = _d(cds); dp = _i(data_points)
        S, N = X.shape[0], X.shape[1]; M = dp.shape[0]
        out = np.zeros(S * N * 3); md = np.zeros(max(M, 1))
        rc = self.L.oracle_likelihood_gradient(S, N, M, _p(X), float(alpha), float(norm), _p(cds),
                                               dp.ctypes.data_as(_ip), _p(out), _p(md))
        if rc:
            raise MemoryError("oracle scratch allocation failed")
        return out

    def forcefield_energy(self, s, radii, radii2, k):
        s = _d(s); r = _d(radii); r2 = _d(radii2)
        return self.L.oracle_ff_energy(len(s), _p(s), _p(r), _p(r2), float(k))

    def forcefield_gradient(self, s, radii, radii2, k):
        s = _d(s); r = _d(radii); r2 = _d(radii2); out = np.zeros(3 * len(s))
        self.L.oracle_ff_gradient(len(s), _p(s), _p(r), _p(r2), float(k), _p(out))
        return out

    def backbone_prior_gradient(self, structure, ll, ul, k):
        s = _d(structure).ravel(); n = len(s) // 3
        ll = _d(ll); ul = _d(ul); out = np.zeros(3 * n)
        self.L.oracle_backbone_gradient(n, _p(s), _p(ll), _p(ul), float(k), _p(out))
        return out

    def sphere_prior_gradient(self, structure, center, radius, k, beads, bead_radii, bead_radii2):
        s = _d(structure); c = _d(center); br = _d(bead_radii); br2 = _d(bead_radii2)
        out = np.zeros(3 * len(s))
        self.L.oracle_sphere_gradient(len(s), _p(s), _p(c), float(radius), float(k), _p(br), _p(br2), _p(out))
        return out


class OracleNBList(object):
    """nblist.py:51-83 / c/nblist.c restated; ``update`` returns the total contact count."""

    def __init__(self, co, cellsize, n_cells, n_per_cell, n_atoms):
        self.co = co
        self.n_atoms = n_atoms
        self.h = co.L.oracle_nblist_new(float(cellsize), int(n_cells), int(n_per_cell), int(n_atoms))
        if not self.h:
            raise MemoryError("oracle nblist allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.co.L.oracle_nblist_free(self.h); self.h = None

    def set_cellsize(self, c):
        self.co.L.oracle_nblist_set_cellsize(self.h, float(c))

    def update(self, coords, new_box=1):
        c = _d(coords).reshape(-1, 3)
        return self.co.L.oracle_nblist_update(self.h, _p(c), len(c), int(new_box))

    def update_bbox(self, coords):
        c = _d(coords).reshape(-1, 3)
        return self.co.L.oracle_nblist_update_bbox(self.h, _p(c), 3 * len(c))

    def pairs(self):
        """Half list as int32 (n, 2) rows (i, j) in the reference's list order."""
        n = self.co.L.oracle_nblist_pairs(self.h, None, 0)
        out = np.empty((max(n, 1), 2), dtype=np.int32)
        self.co.L.oracle_nblist_pairs(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]

    def canonical_pairs(self):
        """Sorted (min, max) rows -- the order-free form compared in parity tests (quirk B-3)."""
        p = self.pairs()
        p = np.sort(p, axis=1)
        return p[np.lexsort((p[:, 1], p[:, 0]))]

    def n_dropped(self):
        return self.co.L.oracle_nblist_n_dropped(self.h)


class OracleNBLForceField(object):
    """forcefields.py:152-224 NBLForceField: NBList(cell = max(r_i+r_j)+1e-3, 90 cells,
    n_per_cell = n_beads) + PROLSQ over the list.  energy() and gradient() each rebuild
    the list, as isd_forcefield.py:76-90 does."""

    def __init__(self, co, bead_radii, force_constant):
        self.co = co
        self.bead_radii = _d(bead_radii)
        self.force_constant = float(force_constant)
        n = len(self.bead_radii)
        self.n_beads = n
        self.nbl = OracleNBList(co, 1.0 + 1e-3, 90, n, n)
        d = np.add.outer(self.bead_radii, self.bead_radii)
        self.nbl.set_cellsize(d.max() + 1e-3)         # forcefields.py:199-201

    def energy(self, structure):
        c = _d(structure).reshape(-1, 3)
        self.nbl.update(c, 1)
        return self.co.L.oracle_nbl_energy(self.nbl.h, _p(c), _p(self.bead_radii),
                                           self.force_constant, len(c))

    def gradient(self, structure):
        c = _d(structure).reshape(-1, 3)
        forces = np.zeros(c.shape)
        self.nbl.update(c, 1)
        self.co.L.oracle_nbl_gradient(self.nbl.h, _p(c), _p(forces), _p(self.bead_radii),
                                      self.force_constant, len(c), None)
        return forces.ravel()


class RefNBL(object):
    """The REFERENCE'S OWN neighbour list + PROLSQ forcefield (c/nblist.c, c/forcefield.c, c/prolsq.c compiled unmodified
    into oracle/_ref/libref_nbl.so by oracle/build_ref.build_nbl; interface: oracle/ref_nbl_driver.c).  Used only to pin
    OracleNBList / OracleNBLForceField.  ``RefNBL.load()`` returns None when the library has not been built.
    Atom types are 0..n-1 as the reference intends (its int64 -> int aliasing of ``types``, SURVEY quirk B-1, is a property
    of the numpy buffer it is handed, not of this C code)."""

    @staticmethod
    def load():
        path = os.path.join(HERE, "_ref", "libref_nbl.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_nblist_new.restype = vp; L.ref_nblist_new.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_nblist_free.argtypes = [vp]
        L.ref_nblist_update.argtypes = [vp, vp, C.c_int, C.c_int]
        L.ref_nblist_pairs.argtypes = [vp, vp, C.c_int]
        L.ref_nblff_new.restype = vp; L.ref_nblff_new.argtypes = [C.c_int, vp, C.c_double]
        L.ref_nblff_free.argtypes = [vp]
        L.ref_nblff_energy.restype = C.c_double; L.ref_nblff_energy.argtypes = [vp, vp]
        L.ref_nblff_gradient.restype = C.c_double; L.ref_nblff_gradient.argtypes = [vp, vp, vp]
        return RefNBL(L)

    def __init__(self, L):
        self.L = L

    def pairs(self, coords, cellsize, n_cells, n_per_cell):
        """(total contacts, half list in the reference's list order) of NBList(cellsize, n_cells, n_per_cell, n).update(coords, 1)"""
        c = _d(coords).reshape(-1, 3)
        h = self.L.ref_nblist_new(float(cellsize), int(n_cells), int(n_per_cell), len(c))
        try:
            tot = self.L.ref_nblist_update(h, c.ctypes.data, len(c), 1)
            out = np.empty((max(tot, 1), 2), dtype=np.int32)
            self.L.ref_nblist_pairs(h, out.ctypes.data, tot)
            return tot, out[:tot]
        finally:
            self.L.ref_nblist_free(h)

    def energy_gradient(self, coords, bead_radii, force_constant):
        """NBLForceField(bead_radii, k): (energy, gradient, energy accumulated by the gradient pass)"""
        c = _d(coords).reshape(-1, 3); r = _d(bead_radii)
        h = self.L.ref_nblff_new(len(c), r.ctypes.data, float(force_constant))
        try:
            E = self.L.ref_nblff_energy(h, c.ctypes.data)
            F = np.zeros(c.shape)
            Eg = self.L.ref_nblff_gradient(h, c.ctypes.data, F.ctypes.data)
            return E, F.ravel(), Eg
        finally:
            self.L.ref_nblff_free(h)


class RefNBLForceField(object):
    """NBLForceField (forcefields.py:152-224) on the reference's own C objects: one neighbour list + PROLSQ forcefield kept
    for the lifetime of the object, the list rebuilt by every energy() and every gradient() as the reference does."""

    def __init__(self, ref, bead_radii, force_constant):
        self.L = ref.L
        self.radii = _d(bead_radii)
        self.n = len(self.radii)
        self.h = self.L.ref_nblff_new(self.n, self.radii.ctypes.data, float(force_constant))
        if not self.h:
            raise MemoryError("reference forcefield allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_nblff_free(self.h); self.h = None

    def energy(self, structure):
        c = _d(structure).reshape(-1, 3)
        return self.L.ref_nblff_energy(self.h, c.ctypes.data)

    def gradient(self, structure):
        c = _d(structure).reshape(-1, 3)
        forces = np.zeros(c.shape)
        self.L.ref_nblff_gradient(self.h, c.ctypes.data, forces.ctypes.data)
        return forces.ravel()


# --------------------------------------------------------------------------------------
# numpy restatements of the log-probabilities (the reference evaluates these in numpy)

def poisson_log_prob(mock_data, counts):
    """error_models.py:85-97: -sum(md) + sum(c * log(md)) (no log c! term)."""
    return -mock_data.sum() + np.sum(counts * np.log(mock_data))


def backbone_log_prob_single(structure, ll, ul, k_bb):
    """backbone_prior.py:64-90."""
    x = structure.reshape(-1, 3)
    d = np.sqrt(np.sum((x[1:] - x[:-1]) ** 2, 1))
    u = d > ul
    l = d < ll
    return -0.5 * k_bb * (np.sum((d[u] - ul[u]) ** 2) + np.sum((ll[l] - d[l]) ** 2))


def sphere_log_prob_single(structure, center, radius, k, bead_radii):
    """sphere_prior.py:62-72."""
    X = structure.reshape(-1, 3)
    norms = np.sqrt(np.sum((X - center[None, :]) ** 2, 1))
    v = norms + bead_radii > radius
    return -0.5 * k * np.sum((norms[v] + bead_radii[v] - radius) ** 2)


def gamma_log_prob(norm, shape, rate):
    """gamma_prior.py:48-53."""
    return (shape - 1.0) * np.log(norm) - norm * rate


class OraclePosterior(object):
    """The posterior assembled by setup_functions.make_posterior (setup_functions.py:64-98),
    evaluated with the reference's own arithmetic.  ``kernels`` is either a
    :class:`COracle` or the dict returned by :func:`load_ref` (the reference's compiled
    Cython).  The neighbour-list forcefield (``nonbonded='nbl'``, the production path) is the
    reference's own C (:class:`RefNBLForceField`) when ``kernels`` is the reference and
    oracle/_ref/libref_nbl.so has been built, else the bit-identical C restatement;
    ``nonbonded='n2'`` selects the O(N^2) forcefield_c path.  ``nbl_kind`` says which."""

    def __init__(self, n_structures, data_points, contact_distances, alpha, bead_radii,
                 k_nb, k_bb, bb_ll, bb_ul, mol_ranges, sphere=None, gamma=None,
                 kernels=None, nonbonded="nbl"):
        self.S = int(n_structures)
        self.data = _i(data_points)
        self.cds = _d(contact_distances)
        self.alpha = float(alpha)
        self.radii = _d(bead_radii)
        self.radii2 = self.radii * self.radii
        self.N = len(self.radii)
        self.k_nb = float(k_nb)
        self.k_bb = float(k_bb)
        self.ll = [_d(a) for a in bb_ll]
        self.ul = [_d(a) for a in bb_ul]
        self.mr = np.asarray(mol_ranges, dtype=int)
        self.sphere = sphere      # None | (radius, k, center[3])
        self.gamma = gamma        # None | (shape, rate)
        self.co = COracle()
        k = kernels if kernels is not None else self.co
        if isinstance(k, dict):
            self.fwm = k["ensemble_contacts_evaluate"]; self.lgrad = k["calculate_gradient"]
            self.ffE = k["forcefield_energy"]; self.ffG = k["forcefield_gradient"]
            self.bbG = k["backbone_prior_gradient"]; self.spG = k["sphere_prior_gradient"]
        else:
            self.fwm = k.ensemble_contacts_evaluate; self.lgrad = k.calculate_gradient
            self.ffE = k.forcefield_energy; self.ffG = k.forcefield_gradient
            self.bbG = k.backbone_prior_gradient; self.spG = k.sphere_prior_gradient
        self.nonbonded = nonbonded
        self.nbl_ff, self.nbl_kind = None, None
        if nonbonded == "nbl":
            ref_nbl = RefNBL.load() if isinstance(k, dict) else None
            if ref_nbl is not None:
                self.nbl_ff, self.nbl_kind = RefNBLForceField(ref_nbl, self.radii, self.k_nb), "reference C (oracle/_ref/libref_nbl.so)"
            else:
                self.nbl_ff, self.nbl_kind = OracleNBLForceField(self.co, self.radii, self.k_nb), "C restatement"

    # ---- pieces -------------------------------------------------------------------
    def mock_data(self, x, norm):
        X = _d(x).reshape(self.S, -1, 3)
        return np.asarray(self.fwm(X, float(norm), self.cds, self.alpha, self.data))

    def log_likelihood(self, x, norm):
        return poisson_log_prob(self.mock_data(x, norm), self.data[:, 2])

    def nb_energy(self, xk):
        if self.nonbonded == "nbl":
            return self.nbl_ff.energy(xk)
        return self.ffE(np.ascontiguousarray(xk.reshape(-1, 3)), self.radii, self.radii2, self.k_nb)

    def nb_gradient(self, xk):
        if self.nonbonded == "nbl":
            return self.nbl_ff.gradient(xk)
        return np.asarray(self.ffG(np.ascontiguousarray(xk.reshape(-1, 3)), self.radii, self.radii2, self.k_nb))

    def energies(self, x, norm):
        """(-log L, sum_k E_nb, -log p_bb, -log p_sphere): the untempered terms."""
        X = _d(x).reshape(self.S, -1, 3)
        U_L = -self.log_likelihood(x, norm)
        E_nb = float(np.sum([self.nb_energy(xk) for xk in X]))
        mr = self.mr
        U_bb = -float(np.sum([np.sum([backbone_log_prob_single(xk[mr[i]:mr[i + 1]], self.ll[i], self.ul[i], self.k_bb)
                                      for i in range(len(mr) - 1)]) for xk in X]))
        U_sp = 0.0
        if self.sphere is not None:
            R, k, c = self.sphere
            U_sp = -float(np.sum([sphere_log_prob_single(xk, _d(c), R, k, self.radii) for xk in X]))
        return U_L, E_nb, U_bb, U_sp

    # ---- posterior ----------------------------------------------------------------
    def log_prob(self, x, norm, lammda=1.0, beta=1.0):
        U_L, E_nb, U_bb, U_sp = self.energies(x, norm)
        lp = -lammda * U_L - beta * E_nb - U_bb - U_sp
        if self.gamma is not None:
            lp += gamma_log_prob(norm, *self.gamma)
        return lp

    def gradient(self, x, norm, lammda=1.0, beta=1.0, terms=("lik", "nb", "bb", "sph")):
        """Gradient of -log posterior w.r.t. structures."""
        X = _d(x).reshape(self.S, -1, 3)
        g = np.zeros(X.size)
        if "lik" in terms:
            g += lammda * np.asarray(self.lgrad(X, self.alpha, float(norm), self.cds, self.data))
        if "nb" in terms:
            g += beta * np.concatenate([self.nb_gradient(xk) for xk in X])
        if "bb" in terms:
            mr = self.mr
            g += np.concatenate([np.concatenate([np.asarray(self.bbG(np.ascontiguousarray(xk[mr[i]:mr[i + 1]]).ravel(),
                                                                      self.ll[i], self.ul[i], self.k_bb))
                                                 for i in range(len(mr) - 1)]) for xk in X])
        if "sph" in terms and self.sphere is not None:
            R, k, c = self.sphere
            g += np.concatenate([np.asarray(self.spG(np.ascontiguousarray(xk), _d(c), float(R), float(k),
                                                     np.arange(len(xk)), self.radii, self.radii2)) for xk in X])
        return g


# --------------------------------------------------------------------------------------
# HMC restatement (csb FastLeapFrog / binf HMCSampler -- sources absent: parity unpinned)

def leapfrog(grad, x, p, eps, L):
    """half-kick, (L-1) x (drift, kick), drift, half-kick: L+1 gradient evaluations."""
    x = x.copy(); p = p.copy()
    p -= 0.5 * eps * grad(x)
    for _ in range(L - 1):
        x += eps * p
        p -= eps * grad(x)
    x += eps * p
    p -= 0.5 * eps * grad(x)
    return x, p


def hmc_step(log_prob, grad, x, eps, L, rng_normal, rng_uniform):
    """One HMC transition.  Returns (x_new, accepted, E0+K0, E1+K1)."""
    p0 = rng_normal(x.shape)
    H0 = -log_prob(x) + 0.5 * np.sum(p0 * p0)
    x1, p1 = leapfrog(grad, x, p0, eps, L)
    H1 = -log_prob(x1) + 0.5 * np.sum(p1 * p1)
    u = rng_uniform()
    acc = u < np.exp(min(0.0, -(H1 - H0))) if np.isfinite(H1) else False
    return (x1 if acc else x), bool(acc), H0, H1
