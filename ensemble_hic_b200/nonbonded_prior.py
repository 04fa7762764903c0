"""Priors on the nonbonded interactions between beads (reference: nonbonded_prior.py)."""
import numpy as np

from . import _lib, _models
from ._binf import AbstractPrior, ArrayParameter, Parameter
from .engine import Model, TERM_NONBONDED


class AbstractNonbondedPrior2(AbstractPrior):
    """log p = sum_k log_ensemble(E(x_k)); gradient of -log p = -concat_k(log_ensemble'(E_k) grad E(x_k)).

    Force fields flagged ``gpu_native`` are evaluated for all structures in one kernel launch;
    any other object with ``energy(structure)`` / ``gradient(structure)`` is driven structure by
    structure exactly like the reference does (nonbonded_prior.py:103-121)."""

    def __init__(self, name, forcefield, n_structures):
        super(AbstractNonbondedPrior2, self).__init__(name)
        self.n_structures = n_structures
        self.forcefield = forcefield
        self._register_variable("structures", differentiable=True)
        self.update_var_param_types(structures=ArrayParameter)
        self._set_original_variables()
        self._engine = None

    def _register_ensemble_parameters(self, **parameters):
        raise NotImplementedError

    def _log_ensemble(self, E):
        raise NotImplementedError

    def _log_ensemble_gradient(self, E):
        raise NotImplementedError

    def _forcefield_energy(self, structure):
        return self.forcefield.energy(structure)

    def _forcefield_gradient(self, structure):
        return self.forcefield.gradient(structure)

    def attach_engine(self, model):
        self._engine = model

    def _gpu_engine(self):
        ff = self.forcefield
        if not getattr(ff, "gpu_native", False):
            return None
        if self._engine is None:
            r = np.ascontiguousarray(ff.bead_radii, dtype=np.float64)
            exact, device = _models.default_exact(), _models.default_device()
            key = ("nb", int(self.n_structures), _models.digest(r), exact, device)
            self._engine = _models.get(key, lambda: Model(self.n_structures, r, exact=exact, device=device))
        self._engine.set_param(_lib.PARAM_NONBONDED_K, float(ff.force_constant))
        return self._engine

    def _ensemble_energies(self, structures):
        """per-call total sum_k E(x_k) on the GPU (Boltzmann ensembles only need the sum)"""
        m = self._gpu_engine()
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return float(m.evaluate(x, terms=TERM_NONBONDED, grad=False, energies=True)["energies"][0, 1])

    def _evaluate_log_prob(self, structures):
        X = np.asarray(structures).reshape(self.n_structures, -1, 3)
        return np.sum([self._log_ensemble(self._forcefield_energy(structure=x)) for x in X])

    def _evaluate_gradient(self, structures):
        X = np.asarray(structures).reshape(self.n_structures, -1, 3)
        res = np.zeros((X.shape[0], X.shape[1] * 3))
        for i, x in enumerate(X):
            E = self._forcefield_energy(structure=x)
            res[i] = self._log_ensemble_gradient(E) * self._forcefield_gradient(structure=x)
        return -res.ravel()


class BoltzmannNonbondedPrior2(AbstractNonbondedPrior2):
    """exp(-beta * E): log_ensemble(E) = -beta E."""

    def __init__(self, name, forcefield, n_structures, beta):
        super(BoltzmannNonbondedPrior2, self).__init__(name, forcefield, n_structures)
        self._register_ensemble_parameters(beta)

    def _register_ensemble_parameters(self, beta):
        self._register("beta")
        self["beta"] = Parameter(beta, "beta")

    def _log_ensemble(self, E):
        return -self["beta"].value * E

    def _log_ensemble_gradient(self, E):
        return -self["beta"].value

    def _evaluate_log_prob(self, structures):
        if self._gpu_engine() is None:
            return super(BoltzmannNonbondedPrior2, self)._evaluate_log_prob(structures)
        return -self["beta"].value * self._ensemble_energies(structures)

    def _evaluate_gradient(self, structures):
        m = self._gpu_engine()
        if m is None:
            return super(BoltzmannNonbondedPrior2, self)._evaluate_gradient(structures)
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return m.evaluate(x, beta=self["beta"].value, terms=TERM_NONBONDED)["grad"][0].copy()

    def clone(self):
        copy = self.__class__(self.name, self.forcefield, self.n_structures, self["beta"].value)
        copy._engine = self._engine
        copy.set_fixed_variables_from_pdf(self)
        return copy


class TsallisNonbondedPrior(AbstractNonbondedPrior2):
    """Tsallis ensemble; make_nonbonded_prior raises NotImplementedError for it in the reference
    (setup_functions.py:518-519) and so does this package."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the Tsallis ensemble is not supported (as in the reference)")
