"""Error models for ensemble-averaged contact data (reference: error_models.py).  Only the
Poisson model is wired into make_likelihood (setup_functions.py:620-624) and only it has a GPU
gradient; the Gaussian and log-normal classes keep their host-side log-probabilities."""
import numpy as np

from ._binf import AbstractErrorModel, ArrayParameter, Parameter


class PoissonEM(AbstractErrorModel):
    def __init__(self, name, data):
        super(PoissonEM, self).__init__(name)
        self._register_variable("mock_data", differentiable=True)
        self.data = data
        self.update_var_param_types(mock_data=ArrayParameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, mock_data):
        """-sum(md) + sum(c log md): Poisson log-probability without the log c! constant."""
        md = np.asarray(mock_data)
        return -md.sum() + np.sum(self.data * np.log(md))

    def _evaluate_gradient(self, **variables):
        return None     # fused with the forward-model Jacobian in the gather kernel

    def clone(self):
        copy = self.__class__(self.name, self.data)
        copy.set_fixed_variables_from_pdf(self)
        return copy


class GaussianEM(AbstractErrorModel):
    def __init__(self, name, data):
        super(GaussianEM, self).__init__(name)
        self.data = data
        self._register_variable("mock_data", differentiable=True)
        self._register_variable("precision")
        self.update_var_param_types(mock_data=ArrayParameter, precision=Parameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, mock_data, precision):
        chi2 = np.sum((np.asarray(mock_data) - self.data) ** 2)
        return -(0.5 * chi2 * precision - 0.5 * len(mock_data) * np.log(precision / 2.0 / np.pi))

    def clone(self):
        copy = self.__class__(self.name, self.data)
        copy.set_fixed_variables_from_pdf(self)
        return copy


class LognormalEM(AbstractErrorModel):
    def __init__(self, name, data):
        super(LognormalEM, self).__init__(name)
        self.data = self.targets = data
        self._register_variable("mock_data", differentiable=True)
        self._register_variable("precision")
        self.update_var_param_types(mock_data=ArrayParameter, precision=Parameter)
        self._set_original_variables()

    def _evaluate_log_prob(self, mock_data, precision):
        return -0.5 * precision * np.sum(np.log(np.asarray(mock_data) / self.data) ** 2)

    def clone(self):
        copy = self.__class__(self.name, self.data)
        copy.set_fixed_variables_from_pdf(self)
        return copy
