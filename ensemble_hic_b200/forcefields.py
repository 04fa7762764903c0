"""Force fields for the repulsive nonbonded term, API of the reference's forcefields.py.

Both classes evaluate on the GPU.  ``ForceField`` mirrors the O(N^2) Cython path
(forcefield_c.pyx); ``NBLForceField`` mirrors the neighbour-list path that make_priors uses
(setup_functions.py:511).  For d0 = r_i + r_j the two are the same function (the reference's own
cross-check, forcefields.py:228-248), so both map onto the tiled all-pairs prior kernel.
"""
import numpy as np

from .forcefield_c import forcefield_energy, forcefield_gradient


class AbstractForceField(object):
    def __init__(self, bead_radii, force_constant):
        self._bead_radii = np.asarray(bead_radii, dtype=np.float64)
        self._bead_radii2 = self._bead_radii * self._bead_radii
        self._force_constant = force_constant

    def energy(self, structure):
        raise NotImplementedError

    def gradient(self, structure):
        raise NotImplementedError

    @property
    def bead_radii(self):
        return self._bead_radii

    @bead_radii.setter
    def bead_radii(self, value):
        self._bead_radii = np.asarray(value, dtype=np.float64)
        self._bead_radii2 = self._bead_radii * self._bead_radii

    @property
    def bead_radii2(self):
        return self._bead_radii2

    @property
    def force_constant(self):
        return self._force_constant

    @force_constant.setter
    def force_constant(self, value):
        self._force_constant = value


class ForceField(AbstractForceField):
    #: marks force fields whose ensemble energy/gradient the nonbonded prior may batch on the GPU
    gpu_native = True

    def energy(self, structure):
        s = np.asarray(structure, dtype=np.float64).reshape(-1, 3)
        return forcefield_energy(s, self.bead_radii, self.bead_radii2, float(self.force_constant))

    def gradient(self, structure):
        s = np.asarray(structure, dtype=np.float64).reshape(-1, 3)
        return forcefield_gradient(s, self.bead_radii, self.bead_radii2, float(self.force_constant))


class NBLForceField(ForceField):
    """forcefields.py:152-224.  Keeps the reference's attributes: ``n_beads`` and an ``_isd_ff``
    PROLSQ object carrying the (n_beads x n_beads) d / k tables and the NBList whose cutoff is
    max(r_i + r_j) + 1e-3."""

    def __init__(self, bead_radii, force_constant):
        super(NBLForceField, self).__init__(bead_radii, force_constant)
        self.n_beads = len(self._bead_radii)
        self._isd_ff = None

    def _make_isd_ff(self):
        from .isd_forcefield import PROLSQ
        from .nblist import NBList
        ff = PROLSQ("PROLSQ")
        n = self.n_beads
        ff.n_types = n
        ff.types = np.arange(n)
        ff.nblist = NBList(1.0 + 1e-3, 90, n, n)
        ff.d = np.add.outer(self._bead_radii, self._bead_radii)
        ff.k = np.zeros((n, n)) + float(self._force_constant)
        ff.nblist.cellsize = ff.d.max() + 1e-3
        return ff

    @property
    def isd_ff(self):
        """the PROLSQ/NBList pair, built on first use (the dense n x n tables are only
        materialised when somebody asks for them: 10 000 beads would be 1.6 GB)"""
        if self._isd_ff is None:
            self._isd_ff = self._make_isd_ff()
        return self._isd_ff
