"""GPU drop-in for the reference's ``ensemble_hic.backbone_prior_c`` (backbone_prior_c.pyx)."""
import numpy as np

from . import _lib, _models
from .engine import Model, TERM_BACKBONE


def backbone_prior_gradient(structure, lower_limits, upper_limits, force_constant):
    """backbone_prior_c.pyx:11-44: gradient of the consecutive-bead distance restraint of ONE
    molecule, ``structure[3n], lower_limits[n-1], upper_limits[n-1], k`` -> fresh ``ndarray[3n]``."""
    s = np.asarray(structure, dtype=np.float64).ravel()
    n = len(s) // 3
    ll = np.ascontiguousarray(lower_limits, dtype=np.float64)[:max(n - 1, 0)]
    ul = np.ascontiguousarray(upper_limits, dtype=np.float64)[:max(n - 1, 0)]
    exact, device = _models.default_exact(), _models.default_device()
    key = ("bb", n, _models.digest(ll, ul), exact, device)
    m = _models.get(key, lambda: Model(1, np.ones(n), bb_lower=ll, bb_upper=ul, exact=exact, device=device))
    m.set_param(_lib.PARAM_BACKBONE_K, float(force_constant))
    return m.evaluate(s.reshape(1, -1), terms=TERM_BACKBONE)["grad"][0].copy()
