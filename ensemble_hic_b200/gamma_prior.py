"""Gamma prior on the scaling factor 'norm' (reference: gamma_prior.py).  Scalar host math."""
import numpy as np

from ._binf import AbstractPrior, Parameter


class GammaPrior(AbstractPrior):
    def __init__(self, name, shape, rate, variable_name):
        super(GammaPrior, self).__init__(name)
        self._register("shape")
        self._register("rate")
        self["shape"] = Parameter(shape, "shape")
        self["rate"] = Parameter(rate, "rate")
        self._variable_name = variable_name
        self._register_variable(variable_name)
        self.update_var_param_types(**{variable_name: Parameter})
        self._set_original_variables()

    def _evaluate_log_prob(self, **variables):
        v = variables[self._variable_name]
        return (self["shape"].value - 1.0) * np.log(v) - v * self["rate"].value


class NormGammaPrior(GammaPrior):
    def __init__(self, shape, rate):
        super(NormGammaPrior, self).__init__("norm_prior", shape, rate, "norm")

    def clone(self):
        copy = self.__class__(self["shape"].value, self["rate"].value)
        copy.set_fixed_variables_from_pdf(self)
        return copy
