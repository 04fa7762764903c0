"""``Forcefield`` / ``PROLSQ`` with the surface of the reference's isd_forcefield.py (the Python
face of c/forcefield.c + c/prolsq.c): quartic repulsion over a neighbour list.

The GPU path supports the parameter tables the reference actually builds
(forcefields.py:181-212): d[i][j] = r_i + r_j and one force constant.  Other tables raise.
"""
import numpy as np

from . import _lib, _models
from .engine import Model, TERM_NONBONDED


class Forcefield(object):
    def __init__(self, name):
        self.name = name
        self.nblist = None
        self.enabled = 1
        self.n_types = 0
        self.types = None
        self._d = None
        self._k = None
        self._model = None

    # ---- parameter tables --------------------------------------------------------------------
    @property
    def d(self):
        return self._d

    @d.setter
    def d(self, value):
        self._d = np.array(value, dtype=np.float64)
        self._model = None

    @property
    def k(self):
        return self._k

    @k.setter
    def k(self, value):
        self._k = np.array(value, dtype=np.float64)

    def is_enabled(self):
        return self.enabled == 1

    def enable(self, enabled=1):
        self.enabled = int(enabled)

    def disable(self):
        self.enable(0)

    def _radii_and_k(self):
        d, k = self._d, self._k
        if d is None or k is None:
            raise ValueError("force field tables d and k are not set")
        r = 0.5 * np.diag(d)
        if not np.allclose(d, np.add.outer(r, r), rtol=0, atol=1e-12 * max(1.0, np.abs(d).max())):
            raise NotImplementedError("GPU PROLSQ supports d[i][j] = r_i + r_j tables only")
        if not np.all(k == k.flat[0]):
            raise NotImplementedError("GPU PROLSQ supports a single force constant only")
        return r, float(k.flat[0])

    def _engine(self):
        r, k = self._radii_and_k()
        if self._model is None:
            exact, device = _models.default_exact(), _models.default_device()
            key = ("ff", _models.digest(r), exact, device)
            self._model = _models.get(key, lambda: Model(1, r, exact=exact, device=device))
        self._model.set_param(_lib.PARAM_NONBONDED_K, k)
        return self._model

    # ---- the reference's methods (isd_forcefield.py:76-90) -------------------------------------
    def update_list(self, coords):
        if self.nblist is not None:
            self.nblist.update(np.asarray(coords).reshape(-1, 3), 1)

    def energy(self, coords, update=True):
        if not self.enabled:
            return 0.0
        x = np.asarray(coords, dtype=np.float64).reshape(1, -1)
        return float(self._engine().evaluate(x, terms=TERM_NONBONDED, grad=False, energies=True)["energies"][0, 1])

    def update_gradient(self, coords, forces):
        """accumulates dE/dx INTO ``forces`` (forcefield.c:99-102); returns the energy."""
        if not self.enabled:
            return None
        x = np.asarray(coords, dtype=np.float64).reshape(1, -1)
        out = self._engine().evaluate(x, terms=TERM_NONBONDED, grad=True, energies=True)
        forces += out["grad"][0].reshape(np.shape(forces))
        return float(out["energies"][0, 1])

    def __str__(self):
        return "%s(n_types=%.2f)" % (self.__class__.__name__, self.n_types)

    __repr__ = __str__


class PROLSQ(Forcefield):
    def __init__(self, name="PROLSQ"):
        super(PROLSQ, self).__init__(name)
