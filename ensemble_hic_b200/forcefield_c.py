"""GPU drop-in for the reference's ``ensemble_hic.forcefield_c`` (forcefield_c.pyx): the O(N^2)
quartic repulsion of one structure, evaluated by the tiled all-pairs prior kernel."""
import numpy as np

from . import _lib, _models
from .engine import Model, TERM_NONBONDED


def _model(radii):
    r = np.ascontiguousarray(radii, dtype=np.float64)
    exact, device = _models.default_exact(), _models.default_device()
    key = ("ff", _models.digest(r), exact, device)
    return _models.get(key, lambda: Model(1, r, exact=exact, device=device))


def forcefield_energy(s, radii, radii2, force_constant):
    """forcefield_c.pyx:11-31: 0.5 k sum_{i<j, d<r_i+r_j} (d - r_i - r_j)^4 -> float."""
    m = _model(radii)
    m.set_param(_lib.PARAM_NONBONDED_K, float(force_constant))
    s = np.asarray(s, dtype=np.float64).reshape(1, -1)
    return float(m.evaluate(s, terms=TERM_NONBONDED, grad=False, energies=True)["energies"][0, 1])


def forcefield_gradient(s, radii, radii2, force_constant):
    """forcefield_c.pyx:33-59 -> fresh ``ndarray[3N]``."""
    m = _model(radii)
    m.set_param(_lib.PARAM_NONBONDED_K, float(force_constant))
    s = np.asarray(s, dtype=np.float64).reshape(1, -1)
    return m.evaluate(s, terms=TERM_NONBONDED)["grad"][0].copy()
