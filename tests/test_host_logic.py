"""Host-side logic that needs no GPU: config / data parsing, schedules, the PDF scaffolding
(parameters, conditioning, cloning), samplers on a toy density, replica-exchange bookkeeping."""
import os
import pickle

import numpy as np
import pytest

CFG = """
[general]
n_structures = 2
n_beads = 23
variables = structures,norm
error_model = poisson
data_file = %(data)s
output_folder = /tmp/hairpin_s/

[forward_model]
contact_distance_factor = 2.0
alpha = 10

[sphere_prior]
radius = 10
force_constant = 5
active=True

[nonbonded_prior]
force_constant = 50
bead_radii = 1

[backbone_prior]
force_constant = 1000

[norm_prior]
shape = 0.1
rate = 0.1

[data_filtering]
disregard_lowest = 0
ignore_sequential_neighbors = 2
include_zero_counts = True

[initial_state]
structures = elongated
norm = 2500

[replica]
n_replicas = 2
schedule = linear
ensemble = boltzmann
lambda_min = 0.0
lambda_max = 1
beta_min = 0.001
beta_max = 1
separate_prior_annealing = False
swap_interval = 5
statistics_update_interval = 40
print_status_interval = 50
n_samples = 20000
samples_dump_interval = 100
samples_dump_step = 20

[structures_hmc]
trajectory_length = 20
timestep = 1e-4
adaption_limit = 10000000
"""


@pytest.fixture()
def cfg(tmp_path):
    rng = np.random.default_rng(0)
    rows = [(i, j, int(rng.poisson(3.0)) if (i + j) % 3 else 0) for i in range(23) for j in range(i + 1, 23)]
    data = tmp_path / "data.txt"
    data.write_text("\n".join("%d\t%d\t%d" % r for r in rows) + "\n")
    path = tmp_path / "config.cfg"
    path.write_text(CFG % {"data": str(data)})
    return str(path), np.array(rows)


def test_parse_config_file_keeps_strings(cfg):
    from ensemble_hic_b200.setup_functions import parse_config_file
    s = parse_config_file(cfg[0])
    assert s["general"]["n_beads"] == "23" and s["sphere_prior"]["active"] == "True"
    assert s["structures_hmc"]["timestep"] == "1e-4"
    assert set(s) >= {"general", "forward_model", "sphere_prior", "nonbonded_prior", "backbone_prior",
                      "norm_prior", "data_filtering", "initial_state", "replica", "structures_hmc"}
    with pytest.raises(IOError):
        parse_config_file("/nonexistent.cfg")


def test_parse_data_filters_like_the_reference(cfg):
    from ensemble_hic_b200.setup_functions import parse_config_file, parse_data
    path, rows = cfg
    s = parse_config_file(path)
    d = parse_data(s["general"]["data_file"], s["data_filtering"])
    keep = rows[np.abs(rows[:, 0] - rows[:, 1]) > 2]
    assert d.shape == keep.shape
    assert np.all(np.diff(d[:, 2]) >= 0)                       # sorted by count
    canon = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(canon(d), canon(keep))
    flt = dict(s["data_filtering"], include_zero_counts="False")
    d2 = parse_data(s["general"]["data_file"], flt)
    assert (d2[:, 2] > 0).all() and len(d2) == (keep[:, 2] > 0).sum()
    flt = dict(s["data_filtering"], disregard_lowest="0.5")
    assert len(parse_data(s["general"]["data_file"], flt)) < len(d)


def test_schedules():
    from ensemble_hic_b200.setup_functions import expspace, load_schedule, make_replica_schedule
    rp = dict(schedule="linear", lambda_min="0.0", lambda_max="1", beta_min="0.001", beta_max="1",
              separate_prior_annealing="False")
    s = make_replica_schedule(rp, 5)
    assert np.allclose(s["lammda"], np.linspace(0, 1, 5)) and np.allclose(s["beta"], np.linspace(0.001, 1, 5))
    rp["separate_prior_annealing"] = "True"
    s = make_replica_schedule(rp, 5)
    assert np.allclose(s["lammda"], [0, 0, 0, 0.5, 1]) and np.allclose(s["beta"], [0.001, 1, 1, 1, 1])
    e = expspace(1e-6, 1.0, 0.1, 64)
    assert e[0] == pytest.approx(1e-6) and e[-1] == pytest.approx(1.0) and np.all(np.diff(e) > 0)
    g = lambda n, lo=1e-6, hi=1.0, a=0.1, N=64: (hi - lo) / (np.exp(a * (N - 1.0)) - 1.0) * (np.exp(a * (n - 1.0)) - 1.0) + lo
    assert np.allclose(e, [g(n) for n in range(1, 65)])
    rp.update(schedule="exponential", lambda_rate="0.1", beta_rate="0.2")
    s = make_replica_schedule(rp, 8)
    assert np.allclose(s["beta"], expspace(0.001, 1.0, 0.2, 8))
    with pytest.raises(ValueError):
        make_replica_schedule(dict(rp, schedule="bogus"), 4)
    s = load_schedule(dict(rp, schedule="x.py", gauss_mean="0.65", gauss_std="0.2"), 6)
    assert s["beta"][-1] == 1.0 and len(s["beta"]) == 7 and np.all(np.diff(s["beta"]) >= 0)


def test_load_pickled_schedule(tmp_path):
    from ensemble_hic_b200.setup_functions import load_schedule
    p = tmp_path / "s.pickle"
    with open(p, "wb") as f:
        pickle.dump({"lammda": np.linspace(0, 1, 4), "beta": np.linspace(0, 1, 4)}, f, protocol=2)
    s = load_schedule({"schedule": str(p)}, 99)
    assert len(s["lammda"]) == 4


# ---- PDF scaffolding ---------------------------------------------------------------------------

def _toy_pdfs():
    from ensemble_hic_b200._binf import AbstractPrior, ArrayParameter, Parameter

    class Quad(AbstractPrior):
        def __init__(self, name, k):
            super(Quad, self).__init__(name)
            self._register("k_" + name); self["k_" + name] = Parameter(k, "k_" + name)
            self._register_variable("structures", differentiable=True)
            self._register_variable("norm")
            self.update_var_param_types(structures=ArrayParameter, norm=Parameter)
            self._set_original_variables()

        def _evaluate_log_prob(self, structures, norm):
            return -0.5 * self["k_" + self.name].value * norm * np.sum(np.asarray(structures) ** 2)

        def _evaluate_gradient(self, structures, norm):
            return self["k_" + self.name].value * norm * np.asarray(structures)

        def clone(self):
            c = self.__class__(self.name, self["k_" + self.name].value)
            c.set_fixed_variables_from_pdf(self)
            return c
    return Quad


def test_posterior_plumbing():
    from ensemble_hic_b200._binf import Posterior
    Quad = _toy_pdfs()
    a, b = Quad("a", 2.0), Quad("b", 3.0)
    P = Posterior({"a": a}, {"b": b})
    x = np.arange(6.0)
    assert P.variables == {"structures", "norm"}
    assert P.log_prob(structures=x, norm=2.0) == a.log_prob(structures=x, norm=2.0) + b.log_prob(structures=x, norm=2.0)
    assert np.array_equal(P.gradient(structures=x, norm=2.0), (2.0 + 3.0) * 2.0 * x)
    P["k_a"].set(10.0)                                   # shared Parameter objects
    assert a["k_a"].value == 10.0
    C = P.conditional_factory(norm="4")                  # strings are cast like csb Parameters
    assert C.variables == {"structures"} and C["norm"].value == 4.0
    assert np.array_equal(C.gradient(structures=x), (10.0 + 3.0) * 4.0 * x)
    assert P.variables == {"structures", "norm"}         # conditioning copies
    C["k_a"].set(1.0)
    assert a["k_a"].value == 10.0
    with pytest.raises(ValueError):
        a.fix_variables(nonsense=1.0)
    with pytest.raises(KeyError):
        P["nonsense"]


def test_nonbonded_prior_conventions_with_mock_forcefield():
    """tests/cases/nonbonded_prior.py:8-94: log_prob = sum_k log_ens(E_k); gradient = -concat(log_ens' grad E_k)"""
    from ensemble_hic_b200.forcefields import AbstractForceField
    from ensemble_hic_b200.nonbonded_prior import AbstractNonbondedPrior2, BoltzmannNonbondedPrior2

    class MockFF(AbstractForceField):
        def energy(self, structure):
            return structure.sum()

        def gradient(self, structure):
            return structure.ravel()

    class MockPrior(AbstractNonbondedPrior2):
        def _log_ensemble(self, E):
            return E

        def _log_ensemble_gradient(self, E):
            return 1.0

    ff = MockFF(np.array([1.0]), 42.0)
    prior = MockPrior("bla", ff, 2)
    X = np.random.default_rng(0).uniform(size=(2, 2, 3))
    assert prior._evaluate_log_prob(X.ravel()) == ff.energy(X[0]) + ff.energy(X[1])
    assert np.array_equal(prior._evaluate_gradient(X.ravel()), -np.concatenate((ff.gradient(X[0]), ff.gradient(X[1]))))
    B = BoltzmannNonbondedPrior2("bla", ff, 2, 42.0)
    assert B._log_ensemble(5.0) == -42.0 * 5.0 and B._log_ensemble_gradient(5.0) == -42.0
    assert B.log_prob(structures=X.ravel()) == (-42.0 * ff.energy(X[0])) + (-42.0 * ff.energy(X[1]))
    assert np.array_equal(B.gradient(structures=X.ravel()), 42.0 * X.ravel())
    assert B.clone()["beta"].value == 42.0 and B.clone().forcefield is ff      # force field shared by reference


def test_backbone_flat_limits_and_gamma_prior():
    from ensemble_hic_b200.backbone_prior import BackbonePrior
    from ensemble_hic_b200.gamma_prior import NormGammaPrior
    ll = [np.array([0.0, 0.5]), np.array([1.0, 0.0, 0.0])]
    ul = [np.array([0.5, 1.0]), np.array([2.0, 2.0, 1.0])]
    p = BackbonePrior("bla", ll, ul, 2.0, 2, np.array([0, 3, 7]))
    fl, fu = p._flat_limits()
    assert list(fl) == [0.0, 0.5, 0.0, 1.0, 0.0, 0.0] and list(fu) == [0.5, 1.0, 0.0, 2.0, 2.0, 1.0]
    assert p.clone()["k_bb"].value == 2.0
    g = NormGammaPrior(0.1, 0.2)
    assert g.log_prob(norm=3.0) == (0.1 - 1.0) * np.log(3.0) - 3.0 * 0.2
    assert g.conditional_factory(norm=3.0).log_prob() == g.log_prob(norm=3.0)


def test_hmc_sampler_matches_oracle_restatement_on_a_toy_density():
    from oracle.oracle import hmc_step
    from ensemble_hic_b200._binf import Posterior
    from ensemble_hic_b200.samplers import HMCSampler
    Quad = _toy_pdfs()
    P = Posterior({}, {"a": Quad("a", 2.0)}).conditional_factory(norm=1.5)
    x0 = np.linspace(-1, 1, 12)
    np.random.seed(7)
    s = HMCSampler(P, x0, 0.3, 10, timestep_adaption_limit=5)
    states, acc, eps = [], [], []
    for _ in range(12):
        states.append(s.sample().copy()); acc.append(s.last_move_accepted); eps.append(s.timestep)
    np.random.seed(7)
    x, e = x0.copy(), 0.3
    for it in range(12):
        x, a, H0, H1 = hmc_step(lambda q: P.log_prob(structures=q), lambda q: P.gradient(structures=q), x, e, 10,
                                lambda shape: np.random.normal(size=shape), np.random.random)
        assert a == acc[it] and np.allclose(x, states[it], rtol=0, atol=1e-13)
        if it + 1 < 5:
            e *= 1.05 if a else 0.95
        assert e == pytest.approx(eps[it])
    assert 0 < s.acceptance_rate <= 1 and s.n_samples == 12


def test_gibbs_sampler_order_and_conditioning():
    from ensemble_hic_b200._binf import BinfState, Posterior
    from ensemble_hic_b200.samplers import GibbsSampler, HMCSampler
    Quad = _toy_pdfs()
    P = Posterior({}, {"a": Quad("a", 2.0)})
    seen = []

    class NormSampler(object):
        pdf = None

        def sample(self):
            seen.append(("norm", sorted(self.pdf.variables), self.pdf["structures"].value.copy()))
            return 2.0

    x0 = np.ones(6)
    hmc = HMCSampler(P.conditional_factory(norm=1.0), x0, 0.1, 3)
    np.random.seed(0)
    g = GibbsSampler(P, BinfState({"structures": x0, "norm": 1.0}), {"norm": NormSampler(), "structures": hmc})
    st = g.sample()
    assert st.variables["norm"] == 2.0
    assert seen[0][1] == ["norm"] and np.array_equal(seen[0][2], st.variables["structures"])   # structures first
    assert hmc.pdf.variables == {"structures"} and hmc.pdf["norm"].value == 1.0


# ---- replica-exchange bookkeeping ---------------------------------------------------------------

def test_partition_and_swap_pairs():
    from ensemble_hic_b200.rex import partition, swap_pairs
    assert list(partition(64, 8)) == list(range(0, 65, 8))
    b = partition(302, 8)
    assert b[0] == 0 and b[-1] == 302 and set(np.diff(b)) == {37, 38}
    assert list(partition(2, 1)) == [0, 2]
    assert swap_pairs(5, 0) == [(0, 1), (2, 3)] and swap_pairs(5, 1) == [(1, 2), (3, 4)]
    assert swap_pairs(2, 1) == [] and swap_pairs(1, 0) == []


def test_swap_decisions_use_energies_only():
    from ensemble_hic_b200.rex import decide_swaps
    lam = np.array([0.0, 0.5, 1.0]); bet = np.array([0.1, 0.5, 1.0])
    U = np.array([[10.0, 4.0], [8.0, 3.0], [2.0, 1.0]])             # slot -> (U_L, U_nb)
    slot_of_temp = np.array([2, 0, 1])                              # temperature 0 lives in slot 2 ...
    acc, works = decide_swaps([(0, 1)], slot_of_temp, U, lam, bet, [0.999999])
    a, b = 2, 0
    w = (lam[0] - lam[1]) * (U[b, 0] - U[a, 0]) + (bet[0] - bet[1]) * (U[b, 1] - U[a, 1])
    # the same work from full tempered energies: E_t(s) = lam_t U_L(s) + beta_t U_nb(s) + U_rest(s)
    E = lambda t, s: lam[t] * U[s, 0] + bet[t] * U[s, 1] + 123.0
    assert works[0] == pytest.approx((E(0, b) - E(0, a)) + (E(1, a) - E(1, b))) == pytest.approx(w)
    assert acc[0] == (np.log(0.999999) < -w)
    assert decide_swaps([(0, 1)], slot_of_temp, np.full((3, 2), np.nan), lam, bet, [0.5])[0] == [False]


def test_stored_sample_loader_and_reweighting_matrix(tmp_path):
    """host side of the batch re-evaluation (reference: analysis_functions.py:38-73, 460-464)"""
    import pickle
    from ensemble_hic_b200 import reevaluation
    from ensemble_hic_b200._binf import BinfState
    folder = str(tmp_path) + "/"
    for first in range(0, 100, 20):                       # dump interval 20, dump step 5: 4 states per file
        states = [BinfState({"structures": np.full(6, first + q), "norm": 1.0 + first + q}) for q in (0, 5, 10, 15)]
        with open(folder + "samples_replica2_%d-%d.pickle" % (first, first + 20), "wb") as f:
            pickle.dump(states, f)
    s = reevaluation.load_sr_samples(folder, 2, 101, 20, 40)
    assert [x.variables["norm"] for x in s] == [1.0 + v for v in (40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95)]
    s = reevaluation.load_sr_samples(folder, 2, 101, 20, 5, interval=2)      # burn-in below one file: skip 5 states
    assert len(s) == (20 - 5 + 1) // 2 and s[0].variables["norm"] == 1.0 + 25
    E = np.arange(12.0).reshape(2, 3, 2)
    q = reevaluation.reweighting_matrix(E, {"lammda": [0.0, 1.0], "beta": [0.5, 1.0]})
    assert q.shape == (2, 6)
    np.testing.assert_allclose(q[0], 0.5 * E.reshape(-1, 2)[:, 1])
    np.testing.assert_allclose(q[1], E.reshape(-1, 2).sum(1))
    q = reevaluation.reweighting_matrix(E, {"beta": [0.5, 1.0]})
    np.testing.assert_allclose(q[1], E.reshape(-1, 2)[:, 1])


def test_model_cache_eviction_keeps_models_alive_for_their_holders():
    """Evicting a model from the cache of the stateless entry points must not destroy it: priors and forward
    models (and their clones) keep using the engine they were built with."""
    from ensemble_hic_b200 import _models

    class Dummy(object):
        closed = False

        def close(self):
            self.closed = True

    _models.clear()
    held = [_models.get(("k", q), Dummy) for q in range(_models._MAX + 3)]
    assert len(_models._CACHE) == _models._MAX
    assert not any(d.closed for d in held)
    assert _models.get(("k", _models._MAX + 2), Dummy) is held[-1]          # still cached
    assert _models.get(("k", 0), Dummy) is not held[0]                        # evicted: rebuilt on demand
    _models.clear()
    assert not any(d.closed for d in held)
