"""Minimal re-provision of the third-party PDF scaffolding the reference builds on.

The reference's PDF classes subclass ``binf`` (pinned at 7cc7910 in shell.nix:18-26) and use
``csb.statistics.pdf.parameterized.Parameter`` (csb 1.2.5).  Neither package is vendored in the
reference tree nor installed here, so this module restates -- from the reference's call sites
(grep over ensemble_hic/*.py) -- exactly the surface the hot path touches:

    Parameter, ArrayParameter, AbstractBinfPDF (log_prob / gradient / clone /
    conditional_factory / fix_variables / parameter sharing through ``pdf['name']``),
    AbstractForwardModel, AbstractErrorModel, AbstractPrior, BinfLikelihood, Posterior,
    BinfState.

Semantics worth knowing:
  * a *variable* is something ``log_prob(**variables)`` is a function of; fixing a variable
    turns it into a *parameter* (``pdf[name].value``) that ``_complete_variables`` feeds back
    into the evaluation, so ``_evaluate_*`` always sees every original variable;
  * ``Posterior`` shares the Parameter objects of its components, so
    ``posterior['lammda'].set(v)`` (run_simulation.py:89-90) reaches the likelihood.
"""
import numpy as np


class Parameter(object):
    """Scalar parameter; casts to float like csb's Parameter (config values are strings,
    setup_functions.py:390,627)."""

    def __init__(self, value=0.0, name=None):
        self._name = name
        self._value = None
        self.set(value)

    @property
    def name(self):
        return self._name

    @property
    def value(self):
        return self._value

    def set(self, value):
        self._value = float(value)

    def __repr__(self):
        return "%s(%r, %r)" % (self.__class__.__name__, self._value, self._name)


class ArrayParameter(Parameter):
    def set(self, value):
        self._value = np.asarray(value)     # no copy: contact-distance arrays are 16 MB at config D


class AbstractBinfPDF(object):
    def __init__(self, name=""):
        self.name = name
        self._params = {}
        self._variables = set()
        self._differentiable_variables = set()
        self._var_param_types = {}
        self._original_variables = set()

    # ---- parameters ------------------------------------------------------------------------
    def _register(self, name):
        if name not in self._params:
            self._params[name] = None

    def __getitem__(self, name):
        if name not in self._params:
            raise KeyError("%s has no parameter %r" % (self.__class__.__name__, name))
        return self._params[name]

    def __setitem__(self, name, value):
        if name not in self._params:
            raise KeyError("parameter %r is not registered" % (name,))
        self._params[name] = value

    @property
    def parameters(self):
        return tuple(self._params)

    def _set_parameters(self, other):
        """copy this pdf's parameter values into ``other`` (same registered names)"""
        for p in self.parameters:
            if p in other.parameters:
                other[p].set(self[p].value)

    # ---- variables -------------------------------------------------------------------------
    @property
    def variables(self):
        return set(self._variables)

    @property
    def differentiable_variables(self):
        return set(self._differentiable_variables)

    @property
    def var_param_types(self):
        return dict(self._var_param_types)

    def _register_variable(self, name, differentiable=False):
        self._variables.add(name)
        if differentiable:
            self._differentiable_variables.add(name)

    def _delete_variable(self, name):
        self._variables.discard(name)
        self._differentiable_variables.discard(name)

    def update_var_param_types(self, **types):
        self._var_param_types.update(types)

    def _set_original_variables(self):
        self._original_variables = set(self._variables)

    def _get_variables_intersection(self, test_variables):
        return {k: v for k, v in test_variables.items() if k in self._variables}

    def _complete_variables(self, variables):
        for p in self.parameters:
            if p in self._original_variables and p not in variables:
                variables[p] = self[p].value

    def fix_variables(self, **fixed_vars):
        for v, value in fixed_vars.items():
            if v not in self._variables:
                raise ValueError("%r is not a variable of %s" % (v, self.name or self.__class__.__name__))
            self._delete_variable(v)
            self._register(v)
            ptype = self._var_param_types.get(v, Parameter)
            self[v] = ptype(value, v)

    def set_fixed_variables_from_pdf(self, pdf):
        self.fix_variables(**{p: pdf[p].value for p in pdf.parameters
                              if p in pdf._original_variables and p in self._variables})

    # ---- evaluation ------------------------------------------------------------------------
    def log_prob(self, **variables):
        self._complete_variables(variables)
        return self._evaluate_log_prob(**variables)

    def gradient(self, **variables):
        self._complete_variables(variables)
        return self._evaluate_gradient(**variables)

    def _evaluate_log_prob(self, **variables):
        raise NotImplementedError

    def _evaluate_gradient(self, **variables):
        raise NotImplementedError

    def clone(self):
        raise NotImplementedError

    def conditional_factory(self, **fixed_vars):
        result = self.clone()
        result.fix_variables(**result._get_variables_intersection(fixed_vars))
        return result


class AbstractForwardModel(AbstractBinfPDF):
    def __call__(self, **variables):
        self._complete_variables(variables)
        return self._evaluate(**variables)

    def _evaluate(self, **variables):
        raise NotImplementedError


class AbstractErrorModel(AbstractBinfPDF):
    pass


class AbstractPrior(AbstractBinfPDF):
    pass


class BinfLikelihood(AbstractBinfPDF):
    """forward model + error model; variables = their union without 'mock_data'."""

    def __init__(self, name, forward_model, error_model):
        super(BinfLikelihood, self).__init__(name)
        self._forward_model = forward_model
        self._error_model = error_model
        for comp in (forward_model, error_model):
            for v in comp.variables:
                if v != "mock_data":
                    self._register_variable(v, v in comp.differentiable_variables)
            self.update_var_param_types(**{k: t for k, t in comp.var_param_types.items() if k != "mock_data"})
            for p in comp.parameters:
                self._register(p)
                self[p] = comp[p]
            self._original_variables |= {v for v in comp._original_variables if v != "mock_data"}

    @property
    def forward_model(self):
        return self._forward_model

    @property
    def error_model(self):
        return self._error_model

    def _split(self, variables):
        fwm_vars = {v: variables[v] for v in self._forward_model._original_variables if v in variables}
        em_vars = {v: variables[v] for v in self._error_model._original_variables
                   if v in variables and v != "mock_data"}
        return fwm_vars, em_vars

    def _evaluate_log_prob(self, **variables):
        fwm_vars, em_vars = self._split(variables)
        mock = self._forward_model._evaluate(**fwm_vars)
        return self._error_model._evaluate_log_prob(mock_data=mock, **em_vars)

    def fix_variables(self, **fixed_vars):
        for comp in (self._forward_model, self._error_model):
            comp.fix_variables(**comp._get_variables_intersection(fixed_vars))
        for v, value in fixed_vars.items():
            if v not in self._variables:
                raise ValueError("%r is not a variable of %s" % (v, self.name))
            self._delete_variable(v)
            self._register(v)
            owner = self._forward_model if v in self._forward_model.parameters else self._error_model
            self[v] = owner[v]


class Posterior(AbstractBinfPDF):
    """Product of likelihoods and priors; shares the components' Parameter objects."""

    def __init__(self, likelihoods, priors, name=""):
        super(Posterior, self).__init__(name)
        self._likelihoods = dict(likelihoods)
        self._priors = dict(priors)
        self._setup()

    def _setup(self):
        for comp in self._components.values():
            for v in comp.variables:
                self._register_variable(v, v in comp.differentiable_variables)
            self.update_var_param_types(**comp.var_param_types)
            for p in comp.parameters:
                if p not in self._params:
                    self._register(p)
                    self[p] = comp[p]
            self._original_variables |= set(comp._original_variables)
        # a variable fixed in one component but free in another stays a variable
        self._original_variables |= self._variables

    @property
    def _components(self):
        c = dict(self._likelihoods)
        c.update(self._priors)
        return c

    @property
    def likelihoods(self):
        return self._likelihoods

    @property
    def priors(self):
        return self._priors

    def _complete_variables(self, variables):
        for p in self.parameters:
            if p in self._original_variables and p not in variables and p not in self._variables:
                variables[p] = self[p].value

    def _evaluate_log_prob(self, **variables):
        return sum(f.log_prob(**f._get_variables_intersection(variables)) for f in self._components.values())

    def _evaluate_gradient(self, **variables):
        res = 0.0
        for f in self._components.values():
            if f.differentiable_variables & set(variables):
                res = res + f.gradient(**f._get_variables_intersection(variables))
        return res

    def clone(self):
        return self.__class__({k: v.clone() for k, v in self._likelihoods.items()},
                              {k: v.clone() for k, v in self._priors.items()}, self.name)

    def conditional_factory(self, **fixed_vars):
        cond = lambda f: f.conditional_factory(**f._get_variables_intersection(fixed_vars))
        return self.__class__({k: cond(v) for k, v in self._likelihoods.items()},
                              {k: cond(v) for k, v in self._priors.items()}, self.name)


class BinfState(object):
    """dict-of-variables state (binf.samplers.BinfState): ``.variables``, ``update_variables``."""

    def __init__(self, variables=None, momenta=None):
        self._variables = dict(variables or {})
        self._momenta = dict(momenta or {})

    @property
    def variables(self):
        return self._variables

    @property
    def momenta(self):
        return self._momenta

    def update_variables(self, **variables):
        self._variables.update(variables)

    def keys(self):
        return self._variables.keys()

    def __getitem__(self, k):
        return self._variables[k]

    def clone(self):
        return BinfState({k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in self._variables.items()},
                         dict(self._momenta))
