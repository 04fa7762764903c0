"""Error model for ensemble-averaged contact data (reference: error_models.py:75-112).  Only the Poisson
model is wired into make_likelihood (setup_functions.py:620-624); the reference's Gaussian and log-normal
classes are out of scope and not mirrored."""
import numpy as np

from ._binf import AbstractErrorModel, ArrayParameter


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
