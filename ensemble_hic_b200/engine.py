"""Thin Python face of the C-ABI model object (include/ehic_b200.h).

:class:`Model` owns one ``ehic_model_t``: the static contact list, bead radii and prior
parameters of one posterior (what setup_functions.make_posterior builds), resident on one
GPU, able to evaluate up to ``max_replicas`` tempered replicas per call.

Two families of methods:

* host (numpy in / numpy out, copies inside the call) -- used by the reference-named
  drop-in functions and PDF classes;
* device (torch CUDA tensors, raw pointers handed to the C ABI, asynchronous on the
  current torch stream) -- used by the batched HMC / replica-exchange driver.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (TERM_ALL, TERM_BACKBONE, TERM_LIKELIHOOD, TERM_NONBONDED, TERM_SPHERE,  # noqa: F401
                   FLAG_EXACT, FLAG_FP32, N_ENERGIES)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Model(object):
    def __init__(self, n_structures, bead_radii, data_points=None, contact_distances=None,
                 alpha=1.0, k_nb=0.0, k_bb=0.0, bb_lower=None, bb_upper=None, mol_ranges=None,
                 sphere=None, max_replicas=1, exact=False, device=0, fp32=False):
        L = _lib.lib()
        radii = _f64(bead_radii).ravel()
        N = len(radii)
        if data_points is None:
            data_points = np.zeros((0, 3), dtype=np.int64)
        dp = np.ascontiguousarray(data_points, dtype=np.int64).reshape(-1, 3)
        M = dp.shape[0]
        cds = _f64(contact_distances if contact_distances is not None else np.zeros(M)).ravel()
        if len(cds) < M:
            raise ValueError("contact_distances has fewer entries than data_points")
        d = _lib.ModelDesc()
        d.n_beads, d.n_structures, d.n_data = N, int(n_structures), M
        d.data_points = dp.ctypes.data_as(_lib._i64p)
        d.contact_distances = cds.ctypes.data_as(_lib._dp)
        d.alpha = float(alpha)
        d.bead_radii = radii.ctypes.data_as(_lib._dp)
        d.nonbonded_k, d.backbone_k = float(k_nb), float(k_bb)
        keep = [dp, cds, radii]
        if bb_lower is not None:
            bl = _f64(bb_lower).ravel(); keep.append(bl)
            if len(bl) != N - 1:
                raise ValueError("bb_lower must have n_beads - 1 entries")
            d.bb_lower = bl.ctypes.data_as(_lib._dp)
        if bb_upper is not None:
            bu = _f64(bb_upper).ravel(); keep.append(bu)
            if len(bu) != N - 1:
                raise ValueError("bb_upper must have n_beads - 1 entries")
            d.bb_upper = bu.ctypes.data_as(_lib._dp)
        if mol_ranges is not None:
            mr = np.ascontiguousarray(mol_ranges, dtype=np.int32).ravel(); keep.append(mr)
            d.mol_ranges = mr.ctypes.data_as(_lib._i32p)
            d.n_mol_ranges = len(mr)
        if sphere is not None:
            R, k, c = sphere
            c = np.zeros(3) if c is None else _f64(c)
            d.sphere_active, d.sphere_radius, d.sphere_k = 1, float(R), float(k)
            d.sphere_center[0], d.sphere_center[1], d.sphere_center[2] = float(c[0]), float(c[1]), float(c[2])
        d.max_replicas = int(max_replicas)
        d.flags = (FLAG_EXACT if exact else 0) | (FLAG_FP32 if fp32 else 0)
        h = C.c_void_p()
        _lib.check(L.ehic_model_create(C.byref(d), int(device), C.byref(h)))
        self._L, self._h = L, h
        self.N, self.S, self.M = N, int(n_structures), M
        self.max_replicas = int(max_replicas)
        self.device = int(device)
        self.exact = bool(exact)
        self.fp32 = bool(fp32)
        self.sphere_active = sphere is not None

    def close(self):
        if getattr(self, "_h", None):
            self._L.ehic_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------
    def set_param(self, which, value):
        _lib.check(self._L.ehic_model_set_param(self._h, int(which), float(value)))

    def get_param(self, which):
        v = C.c_double()
        _lib.check(self._L.ehic_model_get_param(self._h, int(which), C.byref(v)))
        return v.value

    def info(self):
        out = (C.c_int64 * 8)()
        _lib.check(self._L.ehic_model_info(self._h, out))
        path = "lanes" if out[7] >= 20000 else ("tiles" if out[7] >= 10000 else "rows")
        return dict(N=out[0], S=out[1], M=out[2], max_replicas=out[3], flags=out[4], slots=out[5],
                    device=out[6], kt=out[7], path=path)

    TRAJECTORY_MODES = ("graph", "one_cta", "cluster")

    def trajectory_plan(self, R=1, n_steps=1):
        """how hmc_trajectory_device will run: dict(mode, ctas_per_replica, smem_bytes, launches)"""
        out = (C.c_int64 * 4)()
        _lib.check(self._L.ehic_trajectory_plan(self._h, int(R), int(n_steps), out))
        return dict(mode=self.TRAJECTORY_MODES[out[0]], ctas_per_replica=out[1], smem_bytes=out[2], launches=out[3])

    # ---- per-kernel timing ------------------------------------------------------------------
    KERNEL_CLASSES = ("pack", "forward", "prior", "gather", "finalize", "other")

    def kernel_times(self, fn, repeats=5):
        """Average device milliseconds per call of ``fn`` spent in each kernel class (CUDA events
        recorded by the library on the launching stream)."""
        ms = (C.c_double * 8)(); cnt = (C.c_int64 * 8)()
        _lib.check(self._L.ehic_profile(self._h, 1))
        try:
            _lib.check(self._L.ehic_profile_read(self._h, ms, cnt))     # drop anything pending
            for _ in range(repeats):
                fn()
            _lib.check(self._L.ehic_profile_read(self._h, ms, cnt))
        finally:
            self._L.ehic_profile(self._h, 0)
        return {n: ms[q] / repeats for q, n in enumerate(self.KERNEL_CLASSES) if cnt[q]}

    # ---- host API -------------------------------------------------------------------------
    def _prep(self, x, norm, lammda, beta):
        x = _f64(x)
        per = self.S * self.N * 3
        if x.size % per:
            raise ValueError("structures has %d values, not a multiple of S*N*3 = %d" % (x.size, per))
        R = x.size // per
        bc = lambda v: None if v is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (R,)))
        return x, R, bc(norm), bc(lammda), bc(beta)

    def evaluate(self, x, norm=None, lammda=None, beta=None, terms=TERM_ALL,
                 grad=True, energies=False, md=False):
        """Returns a dict with the requested 'grad' [R, S*N*3], 'energies' [R, 4], 'md' [R, M]."""
        x, R, norm, lammda, beta = self._prep(x, norm, lammda, beta)
        out = {}
        g = np.empty((R, self.S * self.N * 3)) if grad else None
        E = np.empty((R, N_ENERGIES)) if energies else None
        mdo = np.empty((R, self.M)) if md else None
        _lib.check(self._L.ehic_evaluate_host(self._h, R, int(terms), _ptr(x), _ptr(norm), _ptr(lammda),
                                              _ptr(beta), _ptr(g), _ptr(E), _ptr(mdo)))
        if grad:
            out["grad"] = g
        if energies:
            out["energies"] = E
        if md:
            out["md"] = mdo
        return out

    def mock_data(self, x, norm):
        x, R, norm, _, _ = self._prep(x, norm, None, None)
        mdo = np.empty((R, self.M))
        _lib.check(self._L.ehic_mock_data_host(self._h, R, _ptr(x), _ptr(norm), _ptr(mdo)))
        return mdo

    # ---- device API (torch tensors; asynchronous on the current stream) ---------------------
    @staticmethod
    def _dptr(t):
        return None if t is None else t.data_ptr()

    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    def evaluate_device(self, R, x, norm, lammda=None, beta=None, grad=None, energies=None, md=None,
                        terms=TERM_ALL):
        p = self._dptr
        _lib.check(self._L.ehic_evaluate(self._h, int(R), int(terms), p(x), p(norm), p(lammda), p(beta),
                                         p(grad), p(energies), p(md), self._stream()))

    def mock_data_device(self, R, x, norm, md):
        p = self._dptr
        _lib.check(self._L.ehic_mock_data(self._h, int(R), p(x), p(norm), p(md), self._stream()))

    def hmc_trajectory_device(self, R, n_steps, x, p_mom, norm, lammda, beta, eps, H):
        p = self._dptr
        _lib.check(self._L.ehic_hmc_trajectory(self._h, int(R), int(n_steps), p(x), p(p_mom), p(norm),
                                               p(lammda), p(beta), p(eps), p(H), self._stream()))


def select_device(R, n, mask, src, dst):
    """dst[r] = src[r] where mask[r] (int32) is set; rows of n doubles."""
    import torch
    L = _lib.lib()
    _lib.check(L.ehic_select(int(R), int(n), mask.data_ptr(), src.data_ptr(), dst.data_ptr(),
                             torch.cuda.current_stream().cuda_stream))


def nblist_pairs(coords, cutoff, device=0):
    """Canonical (sorted, i<j) neighbour pairs of one structure: |x_i-x_j|^2 <= cutoff^2."""
    L = _lib.lib()
    c = _f64(coords).reshape(-1, 3)
    n = C.c_int64()
    _lib.check(L.ehic_nblist_host(len(c), c.ctypes.data, float(cutoff), None, 0, C.byref(n), int(device)))
    pairs = np.empty((max(n.value, 1), 2), dtype=np.int32)
    if n.value:
        _lib.check(L.ehic_nblist_host(len(c), c.ctypes.data, float(cutoff), pairs.ctypes.data, n.value,
                                      C.byref(n), int(device)))
    return pairs[:n.value]


def launch_count():
    return int(_lib.lib().ehic_launch_count())
