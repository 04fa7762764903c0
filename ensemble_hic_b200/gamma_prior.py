"""Gamma prior on the scaling factor ``norm`` (reference counterpart: gamma_prior.py).

Pure host-side scalar arithmetic: the density is
    log p(v) = (shape - 1) log v - rate v        (unnormalised, as the reference evaluates it)
and it does not depend on the structures, so it never enters a GPU kernel; the fused
posterior adds it on the host (posteriors.GPUPosterior.log_prob).
"""
from math import log

from ._binf import AbstractPrior, Parameter


def gamma_log_density(value, shape, rate):
    """unnormalised log-density of Gamma(shape, rate) at ``value``"""
    return (shape - 1.0) * log(value) - value * rate


class GammaPrior(AbstractPrior):
    """Gamma(shape, rate) prior on one scalar variable.

    ``name`` identifies the pdf inside a posterior, ``variable_name`` is the variable it is a prior
    for; 'shape' and 'rate' are registered Parameters (``prior['shape'].value``)."""

    def __init__(self, name, shape, rate, variable_name):
        super(GammaPrior, self).__init__(name)
        for pname, pvalue in (("shape", shape), ("rate", rate)):
            self._register(pname)
            self[pname] = Parameter(pvalue, pname)
        self._variable_name = variable_name
        self._register_variable(variable_name)
        self.update_var_param_types(**{variable_name: Parameter})
        self._set_original_variables()

    @property
    def hyperparameters(self):
        return self["shape"].value, self["rate"].value

    def _evaluate_log_prob(self, **variables):
        shape, rate = self.hyperparameters
        return gamma_log_density(float(variables[self._variable_name]), shape, rate)


class NormGammaPrior(GammaPrior):
    """the prior make_norm_prior builds: name 'norm_prior', variable 'norm'"""

    def __init__(self, shape, rate):
        super(NormGammaPrior, self).__init__("norm_prior", shape, rate, "norm")

    def clone(self):
        twin = NormGammaPrior(*self.hyperparameters)
        twin.set_fixed_variables_from_pdf(self)
        return twin
