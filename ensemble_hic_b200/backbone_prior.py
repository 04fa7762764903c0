"""Prior restraining distances between consecutive beads (reference: backbone_prior.py)."""
import numpy as np

from . import _lib, _models
from ._binf import AbstractPrior, ArrayParameter, Parameter
from .backbone_prior_c import backbone_prior_gradient
from .engine import Model, TERM_BACKBONE


class BackbonePrior(AbstractPrior):
    """-k_bb/2 [ sum_{d>ul} (d-ul)^2 + sum_{d<ll} (ll-d)^2 ] over consecutive beads of each
    molecule.  lower_limits / upper_limits: one array of length (beads in molecule - 1) per
    molecule; mol_ranges: bead offsets of the molecules (first 0, last n_beads)."""

    def __init__(self, name, lower_limits, upper_limits, k_bb, n_structures, mol_ranges):
        super(BackbonePrior, self).__init__(name)
        self.n_structures = n_structures
        self.lower_limits = lower_limits
        self.upper_limits = upper_limits
        self._register("k_bb")
        self["k_bb"] = Parameter(k_bb, "k_bb")
        self._register_variable("structures", differentiable=True)
        self.update_var_param_types(structures=ArrayParameter)
        self._set_original_variables()
        self._mol_ranges = np.asarray(mol_ranges, dtype=int)
        self._engine = None

    def _flat_limits(self):
        mr = self._mol_ranges
        n = int(mr[-1])
        ll = np.zeros(max(n - 1, 0)); ul = np.zeros(max(n - 1, 0))
        for i in range(len(mr) - 1):
            a, b = int(mr[i]), int(mr[i + 1])
            ll[a:b - 1] = np.asarray(self.lower_limits[i], dtype=np.float64)[:b - 1 - a]
            ul[a:b - 1] = np.asarray(self.upper_limits[i], dtype=np.float64)[:b - 1 - a]
        return ll, ul

    def attach_engine(self, model):
        self._engine = model

    def _gpu_engine(self):
        if self._engine is None:
            ll, ul = self._flat_limits()
            n = int(self._mol_ranges[-1])
            exact, device = _models.default_exact(), _models.default_device()
            key = ("bbp", int(self.n_structures), n, _models.digest(ll, ul, self._mol_ranges), exact, device)
            self._engine = _models.get(key, lambda: Model(self.n_structures, np.ones(n), bb_lower=ll, bb_upper=ul,
                                                          mol_ranges=self._mol_ranges, exact=exact, device=device))
        self._engine.set_param(_lib.PARAM_BACKBONE_K, self["k_bb"].value)
        return self._engine

    # ---- single-structure helpers kept for API parity (tests/cases/backbone_prior.py) -----------
    def _single_structure_log_prob(self, structure, ll, ul):
        n = np.size(structure) // 3
        m = _backbone_single_model(n, ll, ul)
        m.set_param(_lib.PARAM_BACKBONE_K, self["k_bb"].value)
        x = np.asarray(structure, dtype=np.float64).reshape(1, -1)
        return -float(m.evaluate(x, terms=TERM_BACKBONE, grad=False, energies=True)["energies"][0, 2])

    def _single_structure_gradient(self, structure, ll, ul):
        return backbone_prior_gradient(np.asarray(structure).ravel(), ll, ul, self["k_bb"].value)

    def _evaluate_log_prob(self, structures):
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return -float(self._gpu_engine().evaluate(x, terms=TERM_BACKBONE, grad=False, energies=True)["energies"][0, 2])

    def _evaluate_gradient(self, structures):
        x = np.asarray(structures, dtype=np.float64).reshape(1, -1)
        return self._gpu_engine().evaluate(x, terms=TERM_BACKBONE)["grad"][0].copy()

    def clone(self):
        copy = self.__class__(self.name, self.lower_limits, self.upper_limits, self["k_bb"].value,
                              self.n_structures, self._mol_ranges)
        copy._engine = self._engine
        copy.set_fixed_variables_from_pdf(self)
        return copy


def _backbone_single_model(n, ll, ul):
    ll = np.ascontiguousarray(ll, dtype=np.float64)[:max(n - 1, 0)]
    ul = np.ascontiguousarray(ul, dtype=np.float64)[:max(n - 1, 0)]
    exact, device = _models.default_exact(), _models.default_device()
    key = ("bb", n, _models.digest(ll, ul), exact, device)
    return _models.get(key, lambda: Model(1, np.ones(n), bb_lower=ll, bb_upper=ul, exact=exact, device=device))
