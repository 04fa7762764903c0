"""Conjugate Gibbs sampler for the scaling factor (reference: npsamplers.py:42-116).

With a Poisson error model and a Gamma(shape a, rate b) prior, norm | structures is
Gamma(lammda * sum(c) + a, lammda * sum(fwm(norm=1)) + b).  The only expensive piece, the
forward-model evaluation at norm = 1, runs on the GPU."""
import numpy as np


class AbstractGammaSampler(object):
    pdf = None

    def __init__(self, likelihood_name, variable_name):
        self._likelihood_name = likelihood_name
        self._variable_name = variable_name

    def _calculate_shape(self):
        raise NotImplementedError

    def _calculate_rate(self):
        raise NotImplementedError

    def sample(self, state=42):
        rate = self._calculate_rate()
        shape = self._calculate_shape()
        sample = np.random.gamma(shape) / rate
        if sample == 0.0:
            sample += 1e-10
        self.state = sample
        return self.state


class NormGammaSampler(AbstractGammaSampler):
    def __init__(self):
        super(NormGammaSampler, self).__init__("ensemble_contacts", "norm")

    def _get_prior(self):
        return [p for p in self.pdf.priors.values() if "norm" in p._original_variables][0]

    def _check_gamma_prior(self, prior):
        from .gamma_prior import GammaPrior
        return isinstance(prior, GammaPrior)

    def _calculate_shape(self):
        prior = self._get_prior()
        if not self._check_gamma_prior(prior):
            raise NotImplementedError("only Gamma priors are supported for the scaling parameter")
        L = self.pdf.likelihoods[self._likelihood_name]
        return L["lammda"].value * np.sum(L.error_model.data) + prior["shape"].value

    def _calculate_rate(self):
        prior = self._get_prior()
        L = self.pdf.likelihoods[self._likelihood_name]
        md = L.forward_model(norm=1.0)
        return L["lammda"].value * md.sum() + prior["rate"].value
