"""GPU drop-in for the reference's ``ensemble_hic.sphere_prior_c`` (sphere_prior_c.pyx)."""
import numpy as np

from . import _lib, _models
from .engine import Model, TERM_SPHERE


def sphere_prior_gradient(structure, center, radius, k, beads, bead_radii, bead_radii2):
    """sphere_prior_c.pyx:12-38 -> fresh ``ndarray[3N]``.  ``beads`` must be ``arange(N)``: the
    reference indexes its output and the radii by loop position, which is
    only meaningful for that argument."""
    X = np.asarray(structure, dtype=np.float64).reshape(-1, 3)
    beads = np.asarray(beads)
    if len(beads) != len(X) or not np.array_equal(beads, np.arange(len(X))):
        raise ValueError("sphere_prior_gradient: beads must be arange(n_beads)")
    r = np.ascontiguousarray(bead_radii, dtype=np.float64)
    exact, device = _models.default_exact(), _models.default_device()
    key = ("sphere", _models.digest(r), exact, device)
    m = _models.get(key, lambda: Model(1, r, sphere=(1.0, 1.0, None), exact=exact, device=device))
    c = np.asarray(center, dtype=np.float64)
    for which, v in ((_lib.PARAM_SPHERE_RADIUS, radius), (_lib.PARAM_SPHERE_K, k), (_lib.PARAM_SPHERE_CX, c[0]),
                     (_lib.PARAM_SPHERE_CY, c[1]), (_lib.PARAM_SPHERE_CZ, c[2])):
        m.set_param(which, float(v))
    return m.evaluate(X.reshape(1, -1), terms=TERM_SPHERE)["grad"][0].copy()
