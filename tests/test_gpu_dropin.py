"""GPU parity of the reference-named Python surface (drop-in functions, PDF classes, posterior
built from a config file, samplers) against the oracle and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from helpers import make_problem, relerr
from test_host_logic import CFG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


@pytest.fixture(params=["0", "1"], ids=["fast", "exact"])
def mode(request, monkeypatch):
    from ensemble_hic_b200 import _models
    monkeypatch.setenv("EHIC_EXACT", request.param)
    _models.clear()
    yield request.param == "1"
    _models.clear()


def test_reference_kats_on_the_gpu(mode):
    """the reference's own known-answer tests, run against the CUDA kernels"""
    from ensemble_hic_b200.backbone_prior import BackbonePrior
    from ensemble_hic_b200.error_models import PoissonEM
    from ensemble_hic_b200.forward_models import EnsembleContactsFWM
    kats = json.load(open(os.path.join(GOLD, "reference_kats.json")))
    k = kats["forward_models.py:18-46"]
    fwm = EnsembleContactsFWM("bla", 2, np.array(k["contact_distances"]), np.array(k["data_points"]))
    md = fwm(structures=np.array(k["structures"], dtype=float).ravel(), smooth_steepness=1.0, norm=2.0)
    if mode:
        assert md[0] == 2.0                                    # exact symmetric-sigmoid identity
    assert abs(md[0] - 2.0) < 1e-15 and abs(md[1] - 0.39846603) < 1e-8
    k = kats["backbone_prior.py:38-55"]
    ll = [np.array(a) for a in k["ll"]]; ul = [np.array(a) for a in k["ul"]]
    prior = BackbonePrior("bla", ll, ul, k["k_bb"], 2, np.array(k["mol_ranges"]))
    z = np.array([k["structure0_z"], k["structure1_z"]])
    X = np.zeros((2, 7, 3)); X[:, :, 2] = z
    assert prior._single_structure_log_prob(X[0, :3], ll[0], ul[0]) == k["log_prob_mol0"]
    g = prior._single_structure_gradient(X[0, 3:], ll[1], ul[1])
    expect = np.zeros(12); expect[2::3] = k["grad_mol1_z"]
    assert np.array_equal(g, expect)
    total = prior._evaluate_gradient(X.ravel())
    parts = [prior._single_structure_gradient(X[s, a:b].ravel(), ll[m], ul[m])
             for s in range(2) for m, (a, b) in enumerate(((0, 3), (3, 7)))]
    assert np.array_equal(total, np.concatenate(parts))
    lp = prior._evaluate_log_prob(X.ravel())
    assert lp == pytest.approx(sum(prior._single_structure_log_prob(X[s, a:b], ll[m], ul[m])
                                   for s in range(2) for m, (a, b) in enumerate(((0, 3), (3, 7)))), rel=1e-14)
    k = kats["error_models.py:14-20"]
    em = PoissonEM("bla", np.array(k["data"]))
    mdk = np.exp(np.array(k["mock_data_log"]))
    assert em._evaluate_log_prob(mock_data=mdk) == -mdk.sum() + 3 * 3 + 4 * 4


def test_dropin_functions_match_the_oracle(co, mode):
    from ensemble_hic_b200.backbone_prior_c import backbone_prior_gradient
    from ensemble_hic_b200.forcefield_c import forcefield_energy, forcefield_gradient
    from ensemble_hic_b200.forward_models_c import ensemble_contacts_evaluate
    from ensemble_hic_b200.likelihoods_c import calculate_gradient
    from ensemble_hic_b200.sphere_prior_c import sphere_prior_gradient
    pb = make_problem(seed=5, S=3, N=45, density=0.6, uniform_radii=False, step=0.7)
    X, r, dp, cds = pb["X"], pb["radii"], pb["data"], pb["cds"]
    same = (lambda a, b: np.array_equal(a, b)) if mode else (lambda a, b: relerr(a, b) < TOL)
    assert same(ensemble_contacts_evaluate(X, 3.3, cds, 10.0, dp), co.ensemble_contacts_evaluate(X, 3.3, cds, 10.0, dp))
    assert same(calculate_gradient(X, 10.0, 3.3, cds, dp), co.calculate_gradient(X, 10.0, 3.3, cds, dp))
    x0 = np.ascontiguousarray(X[0])
    assert same(forcefield_gradient(x0, r, r * r, 50.0), co.forcefield_gradient(x0, r, r * r, 50.0))
    assert forcefield_energy(x0, r, r * r, 50.0) == pytest.approx(co.forcefield_energy(x0, r, r * r, 50.0), rel=1e-13)
    ll = np.full(44, 0.45); ul = np.full(44, 0.65)
    assert np.array_equal(backbone_prior_gradient(x0.ravel(), ll, ul, 7.0), co.backbone_prior_gradient(x0.ravel(), ll, ul, 7.0))
    c = np.array([0.1, -0.3, 0.2])
    assert np.array_equal(sphere_prior_gradient(x0, c, 2.0, 5.0, np.arange(45), r, r * r),
                          co.sphere_prior_gradient(x0, c, 2.0, 5.0, np.arange(45), r, r * r))
    # fresh arrays every call, inputs untouched, argument-order trap (alpha vs norm) respected
    a = calculate_gradient(X, 10.0, 3.3, cds, dp); b = calculate_gradient(X, 10.0, 3.3, cds, dp)
    assert a is not b and np.array_equal(a, b)
    assert not np.array_equal(calculate_gradient(X, 3.3, 10.0, cds, dp), a)
    with pytest.raises(ValueError):
        ensemble_contacts_evaluate(X.ravel(), 3.3, cds, 10.0, dp)
    with pytest.raises(ValueError):
        sphere_prior_gradient(x0, c, 2.0, 5.0, np.arange(45)[::-1], r, r * r)


@pytest.mark.parametrize("name", ["hairpin_s", "proteins", "nora2012"])
def test_golden_vectors_of_the_shipped_configurations(name, mode):
    """inputs = the reference's data files through its own filtering; expected = its own Cython"""
    from ensemble_hic_b200.engine import Model
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, data, cds, radii = g["X"], g["data"], g["cds"], g["radii"]
    N, S = int(g["N"]), int(g["S"])
    sphere = (g["sphere"][0], g["sphere"][1], np.zeros(3)) if len(g["sphere"]) else None
    m = Model(S, radii, data, cds, alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]),
              sphere=sphere, exact=mode)
    norm, lam, bet = float(g["norm"]), float(g["lammda"]), float(g["beta"])
    same = (lambda a, b: np.array_equal(a, b)) if mode else (lambda a, b: relerr(a, b) < TOL)
    out = m.evaluate(X, norm, terms=1, md=True)
    assert same(out["md"][0], g["md"]) and same(out["grad"][0], g["grad_lik"])
    assert same(m.evaluate(X, norm, terms=2)["grad"][0], g["ff_grad"])
    assert np.array_equal(m.evaluate(X, norm, terms=4)["grad"][0], g["bb_grad"])
    if sphere:
        assert np.array_equal(m.evaluate(X, norm, terms=8)["grad"][0], g["sphere_grad"])
    full = m.evaluate(X, norm, lam, bet, energies=True)
    assert relerr(full["grad"][0], g["grad"]) < TOL
    E = full["energies"][0]
    for q in range(4):
        assert abs(E[q] - g["energies"][q]) <= 1e-11 * max(abs(g["energies"][q]), 1e-300)
    lp = -lam * E[0] - bet * E[1] - E[2] - E[3] + (g["gamma"][0] - 1.0) * np.log(norm) - norm * g["gamma"][1]
    assert abs(lp - float(g["log_prob"])) <= 1e-10 * abs(float(g["log_prob"]))


@pytest.fixture()
def hairpin(tmp_path):
    g = np.load(os.path.join(GOLD, "hairpin_s.npz"))
    # a data file whose filtering reproduces the golden contact list (|i-j| <= 2 rows get filtered out)
    rows = [tuple(r) for r in g["data"]] + [(i, i + 1, 9) for i in range(22)] + [(i, i + 2, 7) for i in range(21)]
    data = tmp_path / "hairpin.txt"
    data.write_text("\n".join("%d\t%d\t%d" % r for r in rows) + "\n")
    cfg = tmp_path / "config.cfg"
    cfg.write_text((CFG % {"data": str(data)}).replace("/tmp/hairpin_s/", str(tmp_path / "out") + "/"))
    return str(cfg), g


def test_posterior_from_config_matches_the_oracle(hairpin, mode):
    from oracle.oracle import OraclePosterior
    from ensemble_hic_b200 import setup_functions as sf
    cfg, g = hairpin
    settings = sf.parse_config_file(cfg)
    p = sf.make_posterior(settings)
    fwm = p.likelihoods["ensemble_contacts"].forward_model
    canon = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(canon(np.asarray(fwm.data_points)), canon(g["data"]))        # contact list bit-exact
    assert p.variables == {"structures", "norm"}
    assert set(p.priors) == {"nonbonded_prior", "backbone_prior", "sphere_prior", "norm_prior"}
    assert np.array_equal(p.priors["nonbonded_prior"].forcefield.bead_radii, np.ones(23))
    N, S = 23, 2
    ul = np.full(N - 1, 2.0)
    op = OraclePosterior(S, np.asarray(fwm.data_points), fwm["contact_distances"].value, 10.0, np.ones(N), 50.0, 1000.0,
                         [np.zeros(N - 1)], [ul], [0, N], sphere=(10.0, 5.0, np.zeros(3)), gamma=(0.1, 0.1))
    x = g["X"].ravel(); norm = 31.5
    for lam, bet in ((1.0, 1.0), (0.0, 0.001), (0.37, 0.61)):
        p["lammda"].set(lam); p["beta"].set(bet)
        lp, ref = p.log_prob(structures=x, norm=norm), op.log_prob(x, norm, lam, bet)
        assert abs(lp - ref) <= 1e-10 * abs(ref)
        assert relerr(p.gradient(structures=x, norm=norm), op.gradient(x, norm, lam, bet)) < TOL
        # component by component (the generic binf-style path) == fused path
        comp = sum(f.log_prob(**f._get_variables_intersection(dict(structures=x, norm=norm))) for f in p._components.values())
        assert abs(comp - lp) <= 1e-10 * abs(lp)
    # conditioning on norm (what the HMC sampler sees) and the replica adaptor
    from ensemble_hic_b200._binf import BinfState
    from ensemble_hic_b200.replica import CompatibleReplica
    c = p.conditional_factory(norm=norm)
    assert c.variables == {"structures"} and c.engine is p.engine
    assert c.log_prob(structures=x) == p.log_prob(structures=x, norm=norm)
    rep = CompatibleReplica("replica1", None, p)
    assert rep.get_energy(BinfState({"structures": x, "norm": norm})) == -p.log_prob(structures=x, norm=norm)
    # norm Gibbs sampler: rate / shape of the conditional Gamma
    from ensemble_hic_b200.npsamplers import NormGammaSampler
    s = NormGammaSampler(); s.pdf = p.conditional_factory(structures=x)
    md1 = op.mock_data(x, 1.0)
    assert s._calculate_rate() == pytest.approx(0.61 * 0 + p["lammda"].value * md1.sum() + 0.1, rel=1e-12)
    assert s._calculate_shape() == pytest.approx(p["lammda"].value * g["data"][:, 2].sum() + 0.1, rel=1e-14)


def test_hmc_acceptance_decisions_match_the_cpu_restatement(hairpin):
    """north star: with a fixed RNG stream, identical HMC accept/reject decisions on the hairpin_s toy"""
    from oracle.oracle import OraclePosterior, hmc_step
    from ensemble_hic_b200 import setup_functions as sf
    cfg, g = hairpin
    settings = sf.parse_config_file(cfg)
    settings["structures_hmc"].update(timestep="2e-3", trajectory_length="20", adaption_limit="6")
    p = sf.make_posterior(settings)
    fwm = p.likelihoods["ensemble_contacts"].forward_model
    N, S = 23, 2
    op = OraclePosterior(S, np.asarray(fwm.data_points), fwm["contact_distances"].value, 10.0, np.ones(N), 50.0, 1000.0,
                         [np.zeros(N - 1)], [np.full(N - 1, 2.0)], [0, N], sphere=(10.0, 5.0, np.zeros(3)), gamma=(0.1, 0.1))
    x0, norm = g["X"].ravel().copy(), 31.5
    state = {"structures": x0, "norm": norm}
    subs = sf.make_subsamplers(p, state, settings["structures_hmc"])
    hmc = subs["structures"]
    np.random.seed(11)
    acc, xs = [], []
    for _ in range(10):
        xs.append(hmc.sample().copy()); acc.append(hmc.last_move_accepted)
    np.random.seed(11)
    x, eps, ref_acc = x0.copy(), 2e-3, []
    for it in range(10):
        x, a, H0, H1 = hmc_step(lambda q: op.log_prob(q, norm), lambda q: op.gradient(q, norm), x, eps, 20,
                                lambda shape: np.random.normal(size=shape), np.random.random)
        ref_acc.append(a)
        assert relerr(xs[it], x) < 1e-10
        if it + 1 < 6:
            eps *= 1.05 if a else 0.95
    assert acc == ref_acc and any(acc)
    assert hmc.timestep == pytest.approx(eps)


def test_gibbs_sweep_runs_end_to_end(hairpin):
    from ensemble_hic_b200 import setup_functions as sf
    from ensemble_hic_b200.samplers import GibbsSampler
    cfg, g = hairpin
    settings = sf.parse_config_file(cfg)
    p = sf.make_posterior(settings)
    np.random.seed(3)
    st = sf.setup_initial_state(settings["initial_state"], p)
    assert st.variables["structures"].shape == (2 * 23 * 3,) and st.variables["norm"] == 2500.0
    subs = sf.make_subsamplers(p, st.variables, settings["structures_hmc"])
    assert set(subs) == {"structures", "norm"}
    gibbs = GibbsSampler(p, st, subs)
    for _ in range(3):
        s = gibbs.sample()
    assert np.isfinite(s.variables["structures"]).all() and s.variables["norm"] > 0


def test_nblist_and_forcefield_objects(co):
    from oracle.oracle import OracleNBLForceField
    from ensemble_hic_b200.forcefields import ForceField, NBLForceField
    pb = make_problem(seed=8, S=1, N=150, step=0.5)
    x, r = pb["X"][0], pb["radii"]
    ff = NBLForceField(r, 50.0)
    oracle = OracleNBLForceField(co, r, 50.0)
    assert ff.energy(x) == pytest.approx(oracle.energy(x), rel=1e-12)
    assert relerr(ff.gradient(x), oracle.gradient(x)) < TOL
    assert ForceField(r, 50.0).energy(x) == pytest.approx(co.forcefield_energy(x, r, r * r, 50.0), rel=1e-12)
    nbl = ff.isd_ff.nblist
    assert nbl.cellsize == 1.0 + 1e-3
    n = nbl.update(x, 1)
    assert n == oracle.nbl.update(x, 1)
    assert np.array_equal(nbl.pairs, oracle.nbl.canonical_pairs())
    c = nbl.contacts
    assert c.shape[0] == 150 and (c >= -1).all() and (c >= 0).sum() == n
    assert np.array_equal(nbl.n_contacts, np.bincount(nbl.pairs[:, 0], minlength=150))
    forces = np.zeros((150, 3))
    E = ff.isd_ff.update_gradient(x, forces)
    assert relerr(forces.ravel(), oracle.gradient(x)) < TOL and E == pytest.approx(oracle.energy(x), rel=1e-12)


def test_run_simulation_and_continuation_single_gpu(hairpin, tmp_path):
    """the launcher end to end on one GPU: output layout of the reference, then a continuation"""
    import glob
    import pickle
    from ensemble_hic_b200 import run_simulation
    cfg, g = hairpin
    txt = open(cfg).read().replace("n_replicas = 2", "n_replicas = 4").replace("samples_dump_interval = 100", "samples_dump_interval = 20") \
                          .replace("samples_dump_step = 20", "samples_dump_step = 5").replace("statistics_update_interval = 40", "statistics_update_interval = 10") \
                          .replace("timestep = 1e-4", "timestep = 1e-3")
    open(cfg, "w").write(txt)
    rex = run_simulation.main([cfg, "--n-samples", "60", "--quiet", "--seed", "3"])
    out = str(tmp_path / "out") + "/"
    assert os.path.exists(out + "config.cfg") and os.path.exists(out + "schedule.pickle")
    sched = pickle.load(open(out + "schedule.pickle", "rb"))
    assert np.allclose(sched["lammda"], np.linspace(0, 1, 4)) and np.allclose(sched["beta"], np.linspace(0.001, 1, 4))
    files = sorted(os.path.basename(f) for f in glob.glob(out + "samples/*.pickle"))
    assert files == sorted("samples_replica%d_%d-%d.pickle" % (t, a, a + 20) for t in range(1, 5) for a in (0, 20, 40))
    states = pickle.load(open(out + "samples/samples_replica4_40-60.pickle", "rb"))
    assert len(states) == 4 and states[-1].variables["structures"].shape == (2 * 23 * 3,) and states[-1].variables["norm"] > 0
    stats = np.loadtxt(out + "statistics/mcmc_stats.txt")
    assert stats.shape[1] == 9 and stats[-1, 0] == 60
    assert sorted(rex.temp_of_slot.tolist()) == [0, 1, 2, 3] and rex.n_hmc_T.sum() > 0
    # continuation: starts at offset 60 from the dumped states and timesteps
    rex2 = run_simulation.main([cfg, "--n-samples", "40", "--quiet", "--continue", "1"])
    assert rex2.step == 60 + 40
    assert os.path.exists(out + "init_continue1/init_states.npy") and os.path.exists(out + "init_continue1/timesteps.npy")
    assert np.allclose(np.load(out + "init_continue1/timesteps.npy"), stats[-1, 2::2])
    assert os.path.exists(out + "samples/samples_replica1_60-80.pickle") and os.path.exists(out + "samples/samples_replica3_80-100.pickle")
    assert os.path.exists(out + "init_continue1/statistics/mcmc_stats.txt")
    first = pickle.load(open(out + "samples/samples_replica2_60-80.pickle", "rb"))[0]
    assert np.isfinite(first.variables["structures"]).all()


def test_batch_reevaluation_of_stored_samples(hairpin, tmp_path):
    """the energy table calculate_DOS builds (analysis_functions.py:440-458), from the stored
    samples of a finished run, batched on the GPU -- against the per-sample, per-component evaluation"""
    from ensemble_hic_b200 import reevaluation, run_simulation
    from ensemble_hic_b200.setup_functions import make_posterior, parse_config_file
    cfg, g = hairpin
    txt = open(cfg).read().replace("n_replicas = 2", "n_replicas = 3").replace("samples_dump_interval = 100", "samples_dump_interval = 20") \
                          .replace("samples_dump_step = 20", "samples_dump_step = 5").replace("timestep = 1e-4", "timestep = 1e-3")
    open(cfg, "w").write(txt)
    run_simulation.main([cfg, "--n-samples", "60", "--quiet", "--seed", "5"])
    energies, sels = reevaluation.replica_energies(cfg, 60, 2, 20, max_replicas=3, rng=np.random.default_rng(0))
    assert energies.shape == (3, 4, 2) and sels.shape == (3, 4)          # 8 stored samples after burn-in, half of them
    settings = parse_config_file(cfg)
    p = make_posterior(settings)
    L, P = p.likelihoods["ensemble_contacts"], p.priors["nonbonded_prior"]
    # the oracle: calculate_DOS's arithmetic (analysis_functions.py:446-468) with the reference's own kernels
    from oracle.oracle import OraclePosterior, load_ref
    fwm = L.forward_model
    N, S = 23, 2
    op = OraclePosterior(S, np.asarray(fwm.data_points), fwm["contact_distances"].value, 10.0, np.ones(N), 50.0, 1000.0,
                         [np.zeros(N - 1)], [np.full(N - 1, 2.0)], [0, N], sphere=(10.0, 5.0, np.zeros(3)),
                         kernels=load_ref(), nonbonded="nbl")
    out = str(tmp_path / "out") + "/"
    for r in range(3):
        samples = reevaluation.load_sr_samples(out + "samples/", r + 1, 61, 20, 20)
        assert len(samples) == 8
        for q, idx in enumerate(sels[r]):
            x = samples[idx]
            U = op.energies(x.variables["structures"], x.variables["norm"])
            np.testing.assert_allclose(energies[r, q], [U[0], U[1]], rtol=1e-10)
            ref = [-L.log_prob(**x.variables), -P.log_prob(structures=x.variables["structures"])]
            np.testing.assert_allclose(energies[r, q], ref, rtol=1e-11)
    sched = {"lammda": np.linspace(0, 1, 3), "beta": np.linspace(0.001, 1, 3)}
    qm = reevaluation.reweighting_matrix(energies, sched)
    assert qm.shape == (3, 12)
    np.testing.assert_allclose(qm[1], 0.5 * energies.reshape(-1, 2)[:, 0] + sched["beta"][1] * energies.reshape(-1, 2)[:, 1])


def test_api_errors():
    from ensemble_hic_b200 import _lib
    from ensemble_hic_b200.engine import Model
    pb = make_problem(seed=1, S=2, N=20)
    m = Model(2, pb["radii"], pb["data"], pb["cds"], alpha=10.0, max_replicas=2)
    X3 = np.stack([pb["X"]] * 3)
    with pytest.raises(_lib.EhicError) as e:
        m.evaluate(X3, 1.0)                                   # 3 replicas > max_replicas
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(ValueError):
        m.evaluate(pb["X"].ravel()[:-3], 1.0)                 # not a multiple of S*N*3
    with pytest.raises(_lib.EhicError):
        m.evaluate(pb["X"], None)                             # likelihood term needs norm
    with pytest.raises(_lib.EhicError) as e:
        Model(2, pb["radii"], np.array([[0, 25, 1]]), np.array([1.0]))      # bead index out of range
    assert "outside" in str(e.value)
    with pytest.raises(_lib.EhicError):
        Model(2, pb["radii"], np.array([[3, 3, 1]]), np.array([1.0]))       # self pair
    with pytest.raises(ValueError):
        Model(2, pb["radii"], pb["data"], pb["cds"][:5])
    empty = Model(2, pb["radii"])                              # no data: priors only
    out = empty.evaluate(pb["X"], None, terms=_lib.TERM_ALL, energies=True)
    assert out["energies"][0, 0] == 0.0 and np.isfinite(out["grad"]).all()
    assert empty.mock_data(pb["X"], 1.0).shape == (1, 0)
