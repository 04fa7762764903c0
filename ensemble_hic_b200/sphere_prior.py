"""Prior confining all beads to a sphere, e.g. a nuclear wall (reference: sphere_prior.py)."""
import numpy as np

from . import _lib, _models
from ._binf import AbstractPrior, ArrayParameter, Parameter
from .engine import Model, TERM_SPHERE
from .sphere_prior_c import sphere_prior_gradient


class SpherePrior(AbstractPrior):
    """-sphere_k/2 * sum_{|x_i - c| + r_i > R} (|x_i - c| + r_i - R)^2 per structure."""

    def __init__(self, name, sphere_radius, sphere_k, n_structures, bead_radii, sphere_center=None):
        super(SpherePrior, self).__init__(name)
        self.n_structures = n_structures
        self.bead_radii = np.asarray(bead_radii, dtype=np.float64)
        self.bead_radii2 = self.bead_radii ** 2
        self._register_variable("structures", differentiable=True)
        self._register("sphere_radius")
        self["sphere_radius"] = Parameter(sphere_radius, "sphere_radius")
        self._register("sphere_k")
        self["sphere_k"] = Parameter(sphere_k, "sphere_k")
        self._register("sphere_center")
        sphere_center = np.zeros(3) if sphere_center is None else sphere_center
        self["sphere_center"] = ArrayParameter(sphere_center, "sphere_center")
        self.update_var_param_types(structures=ArrayParameter)
        self._set_original_variables()
        self._engine = None

    def attach_engine(self, model):
        self._engine = model

    def _gpu_engine(self, n_structures=None):
        S = self.n_structures if n_structures is None else n_structures
        if self._engine is None or self._engine.S != S:
            exact, device = _models.default_exact(), _models.default_device()
            key = ("sph", int(S), _models.digest(self.bead_radii), exact, device)
            eng = _models.get(key, lambda: Model(S, self.bead_radii, sphere=(1.0, 1.0, None), exact=exact, device=device))
            if n_structures is None:
                self._engine = eng
        else:
            eng = self._engine
        c = self["sphere_center"].value
        for which, v in ((_lib.PARAM_SPHERE_RADIUS, self["sphere_radius"].value), (_lib.PARAM_SPHERE_K, self["sphere_k"].value),
                         (_lib.PARAM_SPHERE_CX, c[0]), (_lib.PARAM_SPHERE_CY, c[1]), (_lib.PARAM_SPHERE_CZ, c[2])):
            eng.set_param(which, float(v))
        return eng

    def _single_structure_log_prob(self, structure):
        x = np.asarray(structure, dtype=np.float64).reshape(1, -1)
        return -float(self._gpu_engine(1).evaluate(x, terms=TERM_SPHERE, grad=False, energies=True)["energies"][0, 3])

    def _single_structure_gradient(self, structure):
        X = np.asarray(structure).reshape(-1, 3)
        return sphere_prior_gradient(X, self["sphere_center"].value, self["sphere_radius"].value,
                                     self["sphere_k"].value, np.arange(len(X)), self.bead_radii, self.bead_radii2)

    def _evaluate_log_prob(self, structures):
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return -float(self._gpu_engine().evaluate(x, terms=TERM_SPHERE, grad=False, energies=True)["energies"][0, 3])

    def _evaluate_gradient(self, structures):
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return self._gpu_engine().evaluate(x, terms=TERM_SPHERE)["grad"][0].copy()

    def clone(self):
        copy = self.__class__(name=self.name, sphere_radius=self["sphere_radius"].value,
                              sphere_k=self["sphere_k"].value, n_structures=self.n_structures,
                              bead_radii=self.bead_radii, sphere_center=self["sphere_center"].value)
        copy._engine = self._engine
        copy.set_fixed_variables_from_pdf(self)
        return copy
