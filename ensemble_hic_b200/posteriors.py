"""Posterior with a fused GPU evaluation path.

``GPUPosterior`` is the object ``make_posterior`` returns.  It behaves like the binf Posterior
the reference builds (sum of component log-probabilities / gradients, shared Parameter objects,
``conditional_factory``), but when its components are the standard ones of
setup_functions.make_posterior -- Poisson ``Likelihood`` + Boltzmann nonbonded prior + backbone
prior (+ sphere prior) (+ Gamma norm prior) -- ``log_prob`` and ``gradient`` are ONE call into the
C ABI (`ehic_evaluate_host`) on a resident model that holds all static data, instead of one
native call per component and per structure as in SURVEY.md 3.3 / 3.4.
"""
import numpy as np

from . import _lib
from ._binf import Posterior
from .engine import TERM_ALL, TERM_BACKBONE, TERM_LIKELIHOOD, TERM_NONBONDED, TERM_SPHERE


class GPUPosterior(Posterior):
    def __init__(self, likelihoods, priors, name="", engine=None):
        super(GPUPosterior, self).__init__(likelihoods, priors, name)
        self.engine = engine

    # ---- composition ---------------------------------------------------------------------------
    def _standard(self):
        """the fused path applies iff every component is one the resident model evaluates"""
        from .backbone_prior import BackbonePrior
        from .error_models import PoissonEM
        from .gamma_prior import GammaPrior
        from .likelihoods import Likelihood
        from .nonbonded_prior import BoltzmannNonbondedPrior2
        from .sphere_prior import SpherePrior
        if self.engine is None:
            return False
        for L in self._likelihoods.values():
            if not (isinstance(L, Likelihood) and isinstance(L.error_model, PoissonEM)):
                return False
        if len(self._likelihoods) != 1:
            return False
        for p in self._priors.values():
            if isinstance(p, BoltzmannNonbondedPrior2):
                if not getattr(p.forcefield, "gpu_native", False):
                    return False
            elif not isinstance(p, (BackbonePrior, SpherePrior, GammaPrior)):
                return False
        return True

    def _terms(self):
        from .backbone_prior import BackbonePrior
        from .nonbonded_prior import BoltzmannNonbondedPrior2
        from .sphere_prior import SpherePrior
        t = TERM_LIKELIHOOD
        for p in self._priors.values():
            if isinstance(p, BoltzmannNonbondedPrior2):
                t |= TERM_NONBONDED
            elif isinstance(p, BackbonePrior):
                t |= TERM_BACKBONE
            elif isinstance(p, SpherePrior):
                t |= TERM_SPHERE
        return t

    def sync_engine(self):
        """push the current Parameter values (k_bb, sphere_*, force constant, alpha) into the model"""
        from .backbone_prior import BackbonePrior
        from .nonbonded_prior import BoltzmannNonbondedPrior2
        from .sphere_prior import SpherePrior
        m = self.engine
        L = next(iter(self._likelihoods.values()))
        if "smooth_steepness" in L.parameters:
            m.set_param(_lib.PARAM_ALPHA, L["smooth_steepness"].value)
        for p in self._priors.values():
            if isinstance(p, BoltzmannNonbondedPrior2):
                m.set_param(_lib.PARAM_NONBONDED_K, float(p.forcefield.force_constant))
            elif isinstance(p, BackbonePrior):
                m.set_param(_lib.PARAM_BACKBONE_K, p["k_bb"].value)
            elif isinstance(p, SpherePrior):
                c = p["sphere_center"].value
                m.set_param(_lib.PARAM_SPHERE_RADIUS, p["sphere_radius"].value)
                m.set_param(_lib.PARAM_SPHERE_K, p["sphere_k"].value)
                m.set_param(_lib.PARAM_SPHERE_CX, float(c[0]))
                m.set_param(_lib.PARAM_SPHERE_CY, float(c[1]))
                m.set_param(_lib.PARAM_SPHERE_CZ, float(c[2]))

    def tempering(self):
        """(lammda, beta) of this posterior (run_simulation.py:89-90)"""
        lam = self["lammda"].value if "lammda" in self.parameters else 1.0
        bet = self["beta"].value if "beta" in self.parameters else 1.0
        return lam, bet

    def _norm(self, variables):
        if "norm" in variables:
            return float(variables["norm"])
        return float(self["norm"].value)

    def _structures(self, variables):
        if "structures" in variables:
            return np.asarray(variables["structures"], dtype=np.float64)
        return np.asarray(self["structures"].value, dtype=np.float64)

    # ---- fused evaluation ------------------------------------------------------------------------
    def energies(self, **variables):
        """untempered (-log L, E_nb, -log p_bb, -log p_sphere) of the state"""
        self.sync_engine()
        x = self._structures(variables).reshape(1, -1)
        return self.engine.evaluate(x, self._norm(variables), terms=self._terms(), grad=False, energies=True)["energies"][0]

    def log_prob(self, **variables):
        if not self._standard():
            return super(GPUPosterior, self).log_prob(**variables)
        from .gamma_prior import GammaPrior
        E = self.energies(**variables)
        lam, bet = self.tempering()
        lp = -lam * E[0] - bet * E[1] - E[2] - E[3]
        for p in self._priors.values():
            if isinstance(p, GammaPrior):
                lp += p.log_prob(**p._get_variables_intersection(variables))
        return float(lp)

    def gradient(self, **variables):
        if not self._standard():
            return super(GPUPosterior, self).gradient(**variables)
        self.sync_engine()
        lam, bet = self.tempering()
        x = self._structures(variables).reshape(1, -1)
        out = self.engine.evaluate(x, self._norm(variables), lammda=lam, beta=bet, terms=self._terms())
        return out["grad"][0].copy()

    # ---- copies keep the resident model -------------------------------------------------------------
    def clone(self):
        c = self.__class__({k: v.clone() for k, v in self._likelihoods.items()},
                           {k: v.clone() for k, v in self._priors.items()}, self.name, engine=self.engine)
        return c

    def conditional_factory(self, **fixed_vars):
        cond = lambda f: f.conditional_factory(**f._get_variables_intersection(fixed_vars))
        return self.__class__({k: cond(v) for k, v in self._likelihoods.items()},
                              {k: cond(v) for k, v in self._priors.items()}, self.name, engine=self.engine)
