"""Cache of engine models behind the stateless, reference-signature entry points.

The reference's native functions take the whole problem description on every call
(forward_models_c.pyx:14-18, likelihoods_c.pyx:27-31, ...).  On the GPU the static part (contact
list layout, radii, bonds) lives in an ``ehic_model_t``; this module keys such models by a digest
of the static arguments so that repeated calls -- one per leapfrog step in the reference's
samplers -- reuse the resident model.
"""
import hashlib
import os
from collections import OrderedDict

import numpy as np

from .engine import Model

_CACHE = OrderedDict()
_MAX = int(os.environ.get("EHIC_MODEL_CACHE", "8"))


def default_fp32():
    """EHIC_FP32=1 selects the optional FP32 pair arithmetic (tile kernels for dense contact lists, lane kernels for all others) for posteriors built
    by make_posterior."""
    return os.environ.get("EHIC_FP32", "0") not in ("0", "", "false", "False")


def default_exact():
    """EHIC_EXACT=1 selects the bit-exact arithmetic mode for models built implicitly."""
    return os.environ.get("EHIC_EXACT", "0") not in ("0", "", "false", "False")


def default_device():
    return int(os.environ.get("EHIC_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"<none>")
            continue
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def get(key, factory):
    m = _CACHE.get(key)
    if m is not None:
        _CACHE.move_to_end(key)
        return m
    m = factory()
    _CACHE[key] = m
    while len(_CACHE) > _MAX:
        # only the cache's reference is dropped: priors, forward models and their clones may still hold the
        # evicted model (`_engine`); Model.__del__ frees the device memory when the last holder goes
        _CACHE.popitem(last=False)
    return m


def clear():
    """forget every cached model (device memory is released as soon as nothing else refers to them)"""
    _CACHE.clear()


def contact_model(n_structures, n_beads, data_points, contact_distances, alpha, exact=None, device=None,
                  max_replicas=1):
    """Model holding only a contact list (forward model / likelihood gradient entry points)."""
    exact = default_exact() if exact is None else exact
    device = default_device() if device is None else device
    dp = np.ascontiguousarray(data_points, dtype=np.int64)
    cds = np.ascontiguousarray(contact_distances, dtype=np.float64)[:len(dp)]
    key = ("contacts", int(n_structures), int(n_beads), digest(dp, cds), bool(exact), int(device), int(max_replicas))
    m = get(key, lambda: Model(n_structures, np.ones(n_beads), dp, cds, alpha=alpha, exact=exact,
                               device=device, max_replicas=max_replicas))
    return m
