"""GPU drop-in for the reference's ``ensemble_hic.likelihoods_c`` (likelihoods_c.pyx)."""
import numpy as np

from . import _lib, _models
from .engine import TERM_LIKELIHOOD


def calculate_gradient(structures, alpha, norm, cds, data_points):
    """Gradient of -log L (Poisson error model) w.r.t. the structures; same signature and
    argument order as likelihoods_c.pyx:27-31 (note: ``alpha, norm, cds`` -- not the order of
    ensemble_contacts_evaluate) -> fresh ``ndarray[S*N*3]``, structure-major."""
    X = np.asarray(structures, dtype=np.float64)
    if X.ndim != 3 or X.shape[2] != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % X.ndim)
    m = _models.contact_model(X.shape[0], X.shape[1], data_points, cds, alpha)
    m.set_param(_lib.PARAM_ALPHA, float(alpha))
    return m.evaluate(X, float(norm), terms=TERM_LIKELIHOOD)["grad"][0].copy()
