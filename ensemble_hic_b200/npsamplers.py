"""Gibbs update of the scaling factor ``norm`` (reference counterpart: npsamplers.py).

Poisson counts with rates norm * f_m and a Gamma(a, b) prior on norm are conjugate:

    norm | structures, data  ~  Gamma( lammda * sum_m c_m + a ,  lammda * sum_m f_m + b ),

where f = forward model evaluated at norm = 1 and lammda is the likelihood temperature.  The one
expensive ingredient, f, is a forward-kernel launch on the GPU (``L.forward_model(norm=1.0)``);
everything else is scalar host arithmetic with numpy's global RNG, as in the reference.
"""
import numpy as np


class AbstractGammaSampler(object):
    """Draws ``state ~ Gamma(shape, rate)``; subclasses say what shape and rate are.

    ``pdf`` is assigned by the Gibbs sampler before every draw: the posterior conditioned on all
    other variables."""

    pdf = None

    def __init__(self, likelihood_name, variable_name):
        self._likelihood_name = likelihood_name
        self._variable_name = variable_name
        self.state = None

    def _calculate_shape(self):
        raise NotImplementedError

    def _calculate_rate(self):
        raise NotImplementedError

    def conditional_parameters(self):
        """(shape, rate) of the conditional Gamma, rate first evaluated like the reference does"""
        rate = self._calculate_rate()
        return self._calculate_shape(), rate

    def sample(self, state=42):
        shape, rate = self.conditional_parameters()
        draw = np.random.gamma(shape) / rate
        self.state = draw if draw != 0.0 else 1e-10      # a zero scale would make log(md) diverge
        return self.state


class NormGammaSampler(AbstractGammaSampler):
    """sampler for 'norm' under the 'ensemble_contacts' likelihood with a Poisson error model"""

    def __init__(self):
        super(NormGammaSampler, self).__init__("ensemble_contacts", "norm")

    def _likelihood(self):
        return self.pdf.likelihoods[self._likelihood_name]

    def _get_prior(self):
        for prior in self.pdf.priors.values():
            if "norm" in prior._original_variables:
                return prior
        raise KeyError("the posterior has no prior on 'norm'")

    def _check_gamma_prior(self, prior):
        from .gamma_prior import GammaPrior
        return isinstance(prior, GammaPrior)

    def _calculate_shape(self):
        prior = self._get_prior()
        if not self._check_gamma_prior(prior):
            raise NotImplementedError("the scaling factor can only be sampled under a Gamma prior")
        L = self._likelihood()
        return L["lammda"].value * float(np.sum(L.error_model.data)) + prior["shape"].value

    def _calculate_rate(self):
        L = self._likelihood()
        unscaled = L.forward_model(norm=1.0)             # GPU forward kernel
        return L["lammda"].value * float(unscaled.sum()) + self._get_prior()["rate"].value
