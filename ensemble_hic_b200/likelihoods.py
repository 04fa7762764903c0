"""Tempered likelihood of the contact data (reference: likelihoods.py)."""
import numpy as np

from . import _lib
from ._binf import BinfLikelihood, Parameter
from .engine import TERM_LIKELIHOOD
from .error_models import PoissonEM


class Likelihood(BinfLikelihood):
    """``lammda`` * log p(data | forward model): forward model + error model with a
    replica-exchange temperature and a direct (fused) gradient evaluation."""

    def __init__(self, name, fwm, em, lammda):
        super(Likelihood, self).__init__(name, fwm, em)
        self._register("lammda")
        self["lammda"] = Parameter(lammda, "lammda")

    def _engine_for(self, structures):
        fwm = self.forward_model
        n_beads = np.size(structures) // (3 * fwm.n_structures)
        return fwm.engine(n_beads)

    def _evaluate_log_prob(self, **variables):
        if not isinstance(self.error_model, PoissonEM):
            return super(Likelihood, self)._evaluate_log_prob(**variables)
        # -log L is reduced on the device (no [M] transfer): error_models.py:97 semantics
        m = self._engine_for(variables["structures"])
        m.set_param(_lib.PARAM_ALPHA, float(variables["smooth_steepness"]))
        x = np.asarray(variables["structures"], dtype=np.float64).reshape(1, -1)
        E = m.evaluate(x, float(variables["norm"]), terms=TERM_LIKELIHOOD, grad=False, energies=True)["energies"]
        return -float(E[0, 0])

    def log_prob(self, **variables):
        self._complete_variables(variables)
        return self["lammda"].value * self._evaluate_log_prob(**variables)

    def _evaluate_gradient(self, **variables):
        """lammda * d(-log L)/d structures, flat [S*N*3] (likelihoods.py:48-61)."""
        if not isinstance(self.error_model, PoissonEM):
            raise NotImplementedError("only the Poisson error model has a gradient kernel")
        m = self._engine_for(variables["structures"])
        m.set_param(_lib.PARAM_ALPHA, float(variables["smooth_steepness"]))
        x = np.asarray(variables["structures"], dtype=np.float64).reshape(1, -1)
        out = m.evaluate(x, float(variables["norm"]), lammda=self["lammda"].value, terms=TERM_LIKELIHOOD)
        return out["grad"][0].copy()

    def clone(self):
        copy = self.__class__(self.name, self.forward_model.clone(), self.error_model.clone(),
                              self["lammda"].value)
        return copy

    def conditional_factory(self, **fixed_vars):
        fwm = self.forward_model.clone()
        fwm.fix_variables(**fwm._get_variables_intersection(fixed_vars))
        em = self.error_model.conditional_factory(**self.error_model._get_variables_intersection(fixed_vars))
        return self.__class__(self.name, fwm, em, self["lammda"].value)
