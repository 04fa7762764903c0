"""Forward model back-calculating contact counts from a structure ensemble -- the class
surface of the reference's forward_models.py, evaluated on the GPU."""
import numpy as np

from . import _lib, _models
from ._binf import AbstractForwardModel, ArrayParameter, Parameter


class EnsembleContactsFWM(AbstractForwardModel):
    """md_m = norm * sum_k s(d_km): sum of smoothed single-structure contact matrices, scaled.

    name: unique name ('fwm' in make_likelihood); n_structures: ensemble size;
    contact_distances[M]; data_points[M,3] = (bead i, bead j, count).
    Variables: 'structures' (flat, x.reshape(S, N, 3)), 'smooth_steepness', 'norm'.
    """

    def __init__(self, name, n_structures, contact_distances, data_points):
        super(EnsembleContactsFWM, self).__init__(name)
        self.data_points = data_points
        self.n_structures = n_structures
        self._register("contact_distances")
        self["contact_distances"] = ArrayParameter(contact_distances, "contact_distance")
        self._register_variable("structures", differentiable=True)
        self._register_variable("smooth_steepness")
        self._register_variable("norm")
        self.update_var_param_types(structures=ArrayParameter, smooth_steepness=Parameter, norm=Parameter)
        self._set_original_variables()
        self._engine = None          # a shared posterior model may be injected by make_posterior
        self._engine_key = None

    def engine(self, n_beads):
        """the resident GPU model holding this contact list (built on first use)"""
        cds = self["contact_distances"].value
        if self._engine is not None and self._engine_key == (n_beads, id(cds)):
            return self._engine
        self._engine = _models.contact_model(self.n_structures, n_beads, self.data_points, cds, 1.0)
        self._engine_key = (n_beads, id(cds))
        return self._engine

    def attach_engine(self, model):
        self._engine = model
        self._engine_key = (model.N, id(self["contact_distances"].value))

    def _evaluate(self, structures, smooth_steepness, norm):
        X = np.asarray(structures, dtype=np.float64).reshape(self.n_structures, -1, 3)
        m = self.engine(X.shape[1])
        m.set_param(_lib.PARAM_ALPHA, float(smooth_steepness))
        return m.mock_data(X, float(norm))[0].copy()

    def clone(self):
        copy = self.__class__(name=self.name, n_structures=self.n_structures,
                              contact_distances=self["contact_distances"].value,
                              data_points=self.data_points)
        if self._engine is not None:      # the layout is static: clones share the resident model
            copy._engine = self._engine
            copy._engine_key = (self._engine_key[0], id(copy["contact_distances"].value))
        copy.set_fixed_variables_from_pdf(self)
        return copy

    def _evaluate_jacobi_matrix(self, structures, smooth_steepness, norm):
        raise NotImplementedError("the Jacobian is never formed: the likelihood gradient kernel "
                                  "contracts it with the error-model derivative (likelihoods_c)")
