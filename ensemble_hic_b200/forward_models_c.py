"""GPU drop-in for the reference's ``ensemble_hic.forward_models_c`` (forward_models_c.pyx)."""
import numpy as np

from . import _lib, _models


def ensemble_contacts_evaluate(structures, norm, contact_distances, alpha, data_points):
    """Mock contact counts of an ensemble: same signature, argument order and return value as
    forward_models_c.pyx:14-32 (``structures[S,N,3], norm, contact_distances[M], alpha,
    data_points[M,3]`` -> fresh ``ndarray[M]``), evaluated by the sm_100a forward kernel."""
    X = np.asarray(structures, dtype=np.float64)
    if X.ndim != 3 or X.shape[2] != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % X.ndim)
    dp = np.asarray(data_points)
    if dp.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % dp.ndim)
    m = _models.contact_model(X.shape[0], X.shape[1], dp, contact_distances, alpha)
    m.set_param(_lib.PARAM_ALPHA, float(alpha))
    return m.mock_data(X, float(norm))[0].copy()
