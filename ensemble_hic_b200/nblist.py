"""Neighbour list with the attribute/method surface of the reference's ``NBList``
(ensemble_hic/nblist.py:51-83 over c/nblist.c), built on the GPU.

Membership is exactly the reference's: a pair is listed iff |x_i - x_j|^2 <= cellsize^2 with the
dot product associated as in mathutils.c:13-15 (the cell grid only prunes candidates, so as long
as no cell overflows -- forcefields.py:175 sizes cells for n_beads atoms -- the reference lists
precisely these pairs).  The *order* of the reference's list depends on cell fill order
(SURVEY quirk B-3); this class returns the canonical order: pair (i, j), i < j, stored in row i,
rows sorted by j.
"""
import numpy as np

from . import _models
from .engine import nblist_pairs


class NBList(object):
    def __init__(self, cellsize, n_cells, n_per_cell, n_atoms):
        self.cellsize = float(cellsize)
        self.n_per_cell = int(n_per_cell)
        self.n_atoms = int(n_atoms)
        self.n_cells = int(n_cells)
        self.enabled = 1
        self._pairs = np.zeros((0, 2), dtype=np.int32)
        self._coords = None

    @property
    def ctype(self):          # the reference exposes the C object here; we are our own "ctype"
        return self

    def enable(self, enabled=1):
        self.enabled = int(enabled)

    def disable(self):
        self.enabled = 0

    def update(self, coords, new_box=1):
        """Rebuild the list for ``coords[N,3]``; returns the total number of contacts
        (nblist_update, c/nblist.c:181-303), or -1 when disabled."""
        if not self.enabled:
            return -1
        c = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        if len(c) != self.n_atoms:
            raise ValueError("expected %d atoms, got %d" % (self.n_atoms, len(c)))
        self._coords = c
        self._pairs = nblist_pairs(c, self.cellsize, device=_models.default_device())
        return len(self._pairs)

    def update_bbox(self, coords):
        """size of the cubic bounding box incl. margin (c/nblist.c:88-110, BOX_MARGIN 1e-5)."""
        c = np.asarray(coords, dtype=np.float64).ravel()
        self.origin = c.min() - 1e-5
        return c.max() + 1e-5 - c.min()

    @property
    def pairs(self):
        return self._pairs

    @property
    def n_contacts(self):
        return np.bincount(self._pairs[:, 0], minlength=self.n_atoms).astype(np.int32)

    @property
    def max_n_contacts(self):
        return min(self.n_per_cell * 14, self.n_atoms)     # set_natoms, c/nblist.c:375-376

    @property
    def contacts(self):
        """[n_atoms, max_n_contacts] partner ids, -1 padded (getter c/nblist.c:452-462)."""
        nc = self.n_contacts
        width = max(self.max_n_contacts, int(nc.max()) if len(nc) else 0, 1)
        out = -np.ones((self.n_atoms, width), dtype=np.int32)
        start = np.concatenate([[0], np.cumsum(nc)[:-1]])
        col = np.arange(len(self._pairs)) - start[self._pairs[:, 0]] if len(self._pairs) else np.zeros(0, dtype=int)
        out[self._pairs[:, 0], col] = self._pairs[:, 1]
        return out

    @property
    def sq_distances(self):
        c = self.contacts
        out = np.zeros(c.shape)
        if self._coords is not None and len(self._pairs):
            i, j = self._pairs[:, 0], self._pairs[:, 1]
            d = self._coords[i] - self._coords[j]
            sq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            nc = self.n_contacts
            start = np.concatenate([[0], np.cumsum(nc)[:-1]])
            out[i, np.arange(len(i)) - start[i]] = sq
        return out
