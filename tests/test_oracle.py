"""Pins the CPU oracle: reference known-answer tests, bit-for-bit agreement of the plain-C
restatement with the reference's own compiled Cython (oracle/_ref), committed golden vectors,
and the neighbour-list restatement against brute force / the O(N^2) forcefield."""
import json
import os

import numpy as np
import pytest

from helpers import make_problem, relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    """the reference's own Cython, when oracle/_ref has been built (it travels with the tree)"""
    from oracle.oracle import load_ref
    from oracle import build_ref
    build_ref.build(verbose=False)          # no-op without /root/reference
    r = load_ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return r


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        return json.load(f)


def _impls(co, request):
    out = [("c", co.ensemble_contacts_evaluate, co.backbone_prior_gradient)]
    return out


def test_kat_forward_model(co, kats):
    """tests/cases/forward_models.py:18-46: md[0] == norm exactly, md[1] closed form"""
    k = kats["forward_models.py:18-46"]
    X = np.array(k["structures"], dtype=float)
    cds = np.array(k["contact_distances"]); dp = np.array(k["data_points"])
    md = co.ensemble_contacts_evaluate(X, k["norm"], cds, k["smooth_steepness"], dp)
    assert md[0] == k["norm"] * 1.0 == k["md0"]
    s = lambda d, g, dc: (g * (dc - d) / np.sqrt(1 + (g * (dc - d)) ** 2) + 1) / 2
    c1 = s(np.linalg.norm(X[0, 2] - X[0, 0]), 1.0, cds[1]); c2 = s(np.linalg.norm(X[1, 2] - X[1, 0]), 1.0, cds[1])
    assert md[1] == k["norm"] * (c1 + c2)


def test_kat_backbone(co, kats):
    """tests/cases/backbone_prior.py:38-55 (both the upper- and lower-limit branches)"""
    from oracle.oracle import backbone_log_prob_single
    k = kats["backbone_prior.py:38-55"]
    z0 = np.array(k["structure0_z"]); s0 = np.stack([np.zeros(7), np.zeros(7), z0], 1)
    lp = backbone_log_prob_single(s0[:3], np.array(k["ll"][0]), np.array(k["ul"][0]), k["k_bb"])
    assert lp == k["log_prob_mol0"]
    g = co.backbone_prior_gradient(s0[3:].ravel(), np.array(k["ll"][1]), np.array(k["ul"][1]), k["k_bb"])
    expect = np.zeros(12); expect[2::3] = k["grad_mol1_z"]
    assert np.array_equal(g, expect)


def test_kat_poisson(kats):
    """tests/cases/error_models.py:14-20"""
    from oracle.oracle import poisson_log_prob
    k = kats["error_models.py:14-20"]
    md = np.exp(np.array(k["mock_data_log"])); c = np.array(k["data"])
    assert poisson_log_prob(md, c) == -md.sum() + 3 * 3 + 4 * 4


def test_likelihood_gradient_matches_finite_differences(co):
    """tests/cases/likelihoods.py:46-64 (seeded here; the reference's is unseeded), tol 1e-3"""
    from oracle.oracle import poisson_log_prob
    dp = np.array([[0, 1, 42], [0, 2, 23], [1, 2, 44], [1, 4, 10], [0, 4, 54]])
    cds = np.array([1.0, 0.7, 1.0, 0.4, 1.15, 0.3])
    rng = np.random.default_rng(0)
    for _ in range(10):
        X = rng.uniform(-0.6, 0.6, size=(3, 5, 3))
        g = co.calculate_gradient(X, 3.0, 2.0, cds, dp)
        E = lambda x: -poisson_log_prob(co.ensemble_contacts_evaluate(x.reshape(3, 5, 3), 2.0, cds, 3.0, dp), dp[:, 2])
        x = X.ravel().copy(); E0 = E(x); ng = np.zeros(len(x))
        for i in range(len(x)):
            x[i] += 1e-7; ng[i] = (E(x) - E0) / 1e-7; x[i] -= 1e-7
        assert np.max(np.abs(g - ng)) < 1e-3


@pytest.mark.parametrize("seed,S,N,density", [(0, 3, 40, 1.0), (1, 1, 7, 1.0), (2, 5, 64, 0.3)])
def test_c_oracle_equals_reference_cython_bit_for_bit(co, ref, seed, S, N, density):
    pb = make_problem(seed=seed, S=S, N=N, density=density, uniform_radii=False, step=0.7)
    X, r = pb["X"], pb["radii"]
    assert np.array_equal(co.ensemble_contacts_evaluate(X, 3.3, pb["cds"], 10.0, pb["data"]),
                          ref["ensemble_contacts_evaluate"](X, 3.3, pb["cds"], 10.0, pb["data"]))
    assert np.array_equal(co.calculate_gradient(X, 10.0, 3.3, pb["cds"], pb["data"]),
                          ref["calculate_gradient"](X, 10.0, 3.3, pb["cds"], pb["data"]))
    x0 = np.ascontiguousarray(X[0])
    assert co.forcefield_energy(x0, r, r * r, 50.0) == ref["forcefield_energy"](x0, r, r * r, 50.0)
    assert np.array_equal(co.forcefield_gradient(x0, r, r * r, 50.0), ref["forcefield_gradient"](x0, r, r * r, 50.0))
    ll = np.full(N - 1, 0.4); ul = np.full(N - 1, 0.6)
    assert np.array_equal(co.backbone_prior_gradient(x0.ravel(), ll, ul, 7.0),
                          ref["backbone_prior_gradient"](x0.ravel(), ll, ul, 7.0))
    c = np.array([0.1, 0.2, -0.1])
    assert np.array_equal(co.sphere_prior_gradient(x0, c, 2.0, 5.0, np.arange(N), r, r * r),
                          ref["sphere_prior_gradient"](x0, c, 2.0, 5.0, np.arange(N), r, r * r))


@pytest.mark.parametrize("name", ["hairpin_s", "proteins", "nora2012"])
def test_c_oracle_reproduces_golden_vectors(co, name):
    """golden vectors were produced by the reference's own Cython on the shipped data files"""
    from oracle.oracle import OraclePosterior
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, data, cds, radii = g["X"], g["data"], g["cds"], g["radii"]
    N, S = int(g["N"]), int(g["S"])
    assert np.array_equal(co.ensemble_contacts_evaluate(X, float(g["norm"]), cds, float(g["alpha"]), data), g["md"])
    assert np.array_equal(co.calculate_gradient(X, float(g["alpha"]), float(g["norm"]), cds, data), g["grad_lik"])
    ffG = np.concatenate([co.forcefield_gradient(x, radii, radii * radii, float(g["k_nb"])) for x in X])
    assert np.array_equal(ffG, g["ff_grad"])
    sphere = (g["sphere"][0], g["sphere"][1], np.zeros(3)) if len(g["sphere"]) else None
    ul = radii[:-1] + radii[1:]
    op = OraclePosterior(S, data, cds, float(g["alpha"]), radii, float(g["k_nb"]), float(g["k_bb"]),
                         [np.zeros(N - 1)], [ul], [0, N], sphere=sphere, gamma=tuple(g["gamma"]), nonbonded="nbl")
    assert np.array_equal(op.gradient(X, float(g["norm"]), float(g["lammda"]), float(g["beta"])), g["grad"])
    assert op.log_prob(X, float(g["norm"]), float(g["lammda"]), float(g["beta"])) == float(g["log_prob"])


@pytest.mark.parametrize("seed,N,step", [(0, 60, 0.5), (1, 300, 0.45), (2, 23, 1.8)])
def test_neighbour_list_restatement(co, seed, N, step):
    """c/nblist.c restated: the pair set is exactly {|dx|^2 <= cutoff^2}; energies / gradients over
    the list equal the O(N^2) forcefield_c path (the reference's own cross-check, forcefields.py:228-248)."""
    from oracle.oracle import OracleNBLForceField
    pb = make_problem(seed=seed, S=1, N=N, step=step, uniform_radii=(seed != 1))
    x, r = pb["X"][0], pb["radii"]
    ff = OracleNBLForceField(co, r, 50.0)
    n = ff.nbl.update(x, 1)
    cut2 = (np.add.outer(r, r).max() + 1e-3) ** 2
    brute = [(i, j) for i in range(N) for j in range(i + 1, N)
             if ((x[i, 0] - x[j, 0]) ** 2 + (x[i, 1] - x[j, 1]) ** 2) + (x[i, 2] - x[j, 2]) ** 2 <= cut2]
    assert n == len(brute) and ff.nbl.n_dropped() == 0
    assert np.array_equal(ff.nbl.canonical_pairs(), np.array(brute, dtype=np.int32).reshape(-1, 2))
    E = ff.energy(x); E2 = co.forcefield_energy(x, r, r * r, 50.0)
    assert abs(E - E2) <= 1e-13 * max(abs(E2), 1e-300)
    assert relerr(ff.gradient(x), co.forcefield_gradient(x, r, r * r, 50.0)) < 1e-13


def test_neighbour_list_edge_cases(co):
    from oracle.oracle import OracleNBList
    nb = OracleNBList(co, 1.001, 90, 4, 4)
    x = np.array([[0, 0, 0], [0, 0, 0.5], [100.0, 0, 0], [100.0, 0.9, 0]])     # box >> 90 cells: cell size inflates
    assert nb.update(x, 1) == 2
    assert np.array_equal(nb.canonical_pairs(), [[0, 1], [2, 3]])
    one = OracleNBList(co, 1.001, 90, 1, 1)
    assert one.update(np.zeros((1, 3)), 1) == 0


# ---- neighbour list + PROLSQ against the reference's own C (c/nblist.c, c/forcefield.c, c/prolsq.c) ------------------

@pytest.fixture(scope="module")
def ref_nbl():
    """the reference's neighbour-list / PROLSQ C sources compiled as plain C (oracle/_ref/libref_nbl.so; travels with the tree)"""
    from oracle import build_ref
    from oracle.oracle import RefNBL
    build_ref.build_nbl(verbose=False)        # no-op without /root/reference
    r = RefNBL.load()
    if r is None:
        pytest.skip("oracle/_ref/libref_nbl.so not built (no /root/reference here)")
    return r


@pytest.mark.parametrize("seed,N,uniform,step", [(0, 50, True, 0.8), (1, 300, True, 0.8), (2, 300, False, 0.9), (3, 1000, False, 0.7),
                                                  (4, 64, True, 0.3)])
def test_c_oracle_neighbour_list_equals_reference_c_bit_for_bit(co, ref_nbl, seed, N, uniform, step):
    """c/nblist.c:112-303 + c/forcefield.c:8-109 + c/prolsq.c:8-38, compiled unmodified, against the restatement in
    ehic_oracle.c: the same contacts in the same list order, the same energy and gradient bits"""
    from oracle.oracle import OracleNBList, OracleNBLForceField
    from helpers import random_walk
    rng = np.random.default_rng(seed)
    radii = np.full(N, 0.5) if uniform else rng.uniform(0.35, 0.65, N)
    x = np.ascontiguousarray(random_walk(rng, 1, N, step)[0])
    cell = float(np.add.outer(radii, radii).max() + 1e-3)
    tot, pairs = ref_nbl.pairs(x, cell, 90, N)
    o = OracleNBList(co, cell, 90, N, N)
    assert o.update(x, 1) == tot
    assert np.array_equal(o.pairs(), pairs)
    E, F, Eg = ref_nbl.energy_gradient(x, radii, 11.4)
    ff = OracleNBLForceField(co, radii, 11.4)
    assert ff.energy(x) == E
    assert np.array_equal(ff.gradient(x), F)


def test_c_oracle_neighbour_list_reproduces_the_reference_c_golden_vectors(co):
    """tests/golden/nbl_reference.npz (tests/golden/make_golden_nbl.py: outputs of the reference's own C code) -- runs
    wherever the repository is, with or without /root/reference"""
    from oracle.oracle import OracleNBList, OracleNBLForceField
    g = np.load(os.path.join(GOLD, "nbl_reference.npz"))
    assert len(g["names"]) == 5
    for name in g["names"]:
        x, radii, k = g[name + "_x"], g[name + "_radii"], float(g[name + "_k"])
        o = OracleNBList(co, float(g[name + "_cell"]), 90, len(x), len(x))
        assert o.update(x, 1) == int(g[name + "_total"]), name
        assert np.array_equal(o.pairs(), g[name + "_pairs"]), name
        ff = OracleNBLForceField(co, radii, k)
        assert ff.energy(x) == float(g[name + "_energy"]), name
        assert np.array_equal(ff.gradient(x), g[name + "_gradient"]), name
