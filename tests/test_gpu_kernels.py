"""GPU parity of the CUDA kernels (through the C ABI) against the CPU oracle.

FP64 tolerances: north_star asks for 1e-10 relative on log-posterior and gradient; the
tests below hold the fast mode to 1e-12 (max-norm relative) and the exact mode to
bit-for-bit equality with the reference's Cython arithmetic (oracle == oracle/_ref bit for
bit, see tests/test_oracle.py)."""
import numpy as np
import pytest

from helpers import make_problem, relerr

pytestmark = pytest.mark.gpu

FAST_TOL = 1e-12


def _model(pb, exact, **kw):
    from ensemble_hic_b200.engine import Model
    kw.setdefault("k_nb", 50.0)
    kw.setdefault("k_bb", 500.0)
    return Model(pb["S"], pb["radii"], pb["data"], pb["cds"], alpha=pb["alpha"], exact=exact, **kw)


@pytest.mark.parametrize("S,N", [(1, 5), (2, 23), (3, 40), (5, 70), (7, 33)])
def test_mock_data_exact_is_bitexact(co, S, N):
    pb = make_problem(seed=S * 100 + N, S=S, N=N)
    m = _model(pb, True)
    md = m.mock_data(pb["X"], pb["norm"])[0]
    ref = co.ensemble_contacts_evaluate(pb["X"], pb["norm"], pb["cds"], pb["alpha"], pb["data"])
    assert np.array_equal(md, ref)


@pytest.mark.parametrize("S,N", [(1, 5), (2, 23), (3, 40), (5, 70), (8, 64)])
def test_mock_data_fast(co, S, N):
    pb = make_problem(seed=S * 100 + N, S=S, N=N)
    m = _model(pb, False)
    md = m.mock_data(pb["X"], pb["norm"])[0]
    ref = co.ensemble_contacts_evaluate(pb["X"], pb["norm"], pb["cds"], pb["alpha"], pb["data"])
    assert relerr(md, ref) < FAST_TOL
    np.testing.assert_allclose(md, ref, rtol=1e-10, atol=1e-14 * pb["norm"] * S)


@pytest.mark.parametrize("S,N,density", [(1, 5, 1.0), (2, 23, 1.0), (3, 40, 0.3), (5, 70, 1.0), (6, 100, 0.2)])
def test_likelihood_gradient_exact_is_bitexact(co, S, N, density):
    from ensemble_hic_b200.engine import TERM_LIKELIHOOD
    pb = make_problem(seed=7 + S * 100 + N, S=S, N=N, density=density)
    m = _model(pb, True)
    g = m.evaluate(pb["X"], pb["norm"], terms=TERM_LIKELIHOOD)["grad"][0]
    ref = co.calculate_gradient(pb["X"], pb["alpha"], pb["norm"], pb["cds"], pb["data"])
    assert np.array_equal(g, ref)


@pytest.mark.parametrize("S,N,density", [(1, 5, 1.0), (2, 23, 1.0), (3, 40, 0.3), (5, 70, 1.0), (9, 100, 0.2)])
def test_likelihood_gradient_fast(co, S, N, density):
    from ensemble_hic_b200.engine import TERM_LIKELIHOOD
    pb = make_problem(seed=7 + S * 100 + N, S=S, N=N, density=density)
    m = _model(pb, False)
    g = m.evaluate(pb["X"], pb["norm"], terms=TERM_LIKELIHOOD)["grad"][0]
    ref = co.calculate_gradient(pb["X"], pb["alpha"], pb["norm"], pb["cds"], pb["data"])
    assert relerr(g, ref) < FAST_TOL


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("uniform", [True, False])
def test_nonbonded_gradient_and_energy(co, exact, uniform):
    from ensemble_hic_b200.engine import TERM_NONBONDED
    pb = make_problem(seed=11, S=4, N=90, uniform_radii=uniform, step=0.6)
    m = _model(pb, exact, k_nb=50.0)
    out = m.evaluate(pb["X"], pb["norm"], terms=TERM_NONBONDED, energies=True)
    r = pb["radii"]
    ref_g = np.concatenate([co.forcefield_gradient(x, r, r * r, 50.0) for x in pb["X"]])
    ref_E = sum(co.forcefield_energy(x, r, r * r, 50.0) for x in pb["X"])
    assert np.abs(ref_g).max() > 0
    if exact:
        assert np.array_equal(out["grad"][0], ref_g)
    else:
        assert relerr(out["grad"][0], ref_g) < FAST_TOL
    assert abs(out["energies"][0, 1] - ref_E) <= 1e-12 * abs(ref_E)


def test_nonbonded_matches_neighbour_list_path(co):
    """production reference path: NBList + PROLSQ (c/nblist.c, c/forcefield.c, c/prolsq.c)."""
    from oracle.oracle import OracleNBLForceField
    from ensemble_hic_b200.engine import TERM_NONBONDED
    pb = make_problem(seed=12, S=3, N=120, step=0.6)
    ff = OracleNBLForceField(co, pb["radii"], 50.0)
    m = _model(pb, False, k_nb=50.0)
    out = m.evaluate(pb["X"], pb["norm"], terms=TERM_NONBONDED, energies=True)
    ref_g = np.concatenate([ff.gradient(x) for x in pb["X"]])
    ref_E = sum(ff.energy(x) for x in pb["X"])
    assert relerr(out["grad"][0], ref_g) < FAST_TOL
    assert abs(out["energies"][0, 1] - ref_E) <= 1e-12 * abs(ref_E)


@pytest.mark.parametrize("exact", [False, True])
def test_nonbonded_against_the_reference_c_golden_vectors(exact):
    """tests/golden/nbl_reference.npz holds what the reference's OWN c/nblist.c + c/forcefield.c + c/prolsq.c return on
    seeded structures (make_golden_nbl.py): uniform and mixed radii, a compact globule, a chain far from the origin, a
    nora2012 structure.  The CUDA nonbonded term against those numbers directly"""
    import os
    from ensemble_hic_b200.engine import Model, TERM_NONBONDED
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nbl_reference.npz"))
    for name in g["names"]:
        x, radii, k = g[name + "_x"], g[name + "_radii"], float(g[name + "_k"])
        m = Model(1, radii, k_nb=k, exact=exact)
        out = m.evaluate(x.reshape(1, -1), 1.0, terms=TERM_NONBONDED, energies=True)
        E, G = float(g[name + "_energy"]), g[name + "_gradient"]
        assert relerr(out["grad"][0], G) < (1e-13 if exact else FAST_TOL), name
        assert abs(out["energies"][0, 1] - E) <= 1e-12 * abs(E), name
        m.close()


@pytest.mark.parametrize("exact", [True, False])
def test_backbone_gradient_bitexact_and_energy(co, exact):
    from oracle.oracle import backbone_log_prob_single
    from ensemble_hic_b200.engine import TERM_BACKBONE
    pb = make_problem(seed=13, S=3, N=50, step=1.0)
    N = pb["N"]
    rng = np.random.default_rng(5)
    ll = rng.uniform(0.6, 0.9, N - 1); ul = rng.uniform(0.95, 1.2, N - 1)
    mr = np.array([0, 20, 21, 50])      # includes a single-bead molecule
    m = _model(pb, exact, k_bb=7.5, bb_lower=ll, bb_upper=ul, mol_ranges=mr)
    out = m.evaluate(pb["X"], pb["norm"], terms=TERM_BACKBONE, energies=True)
    ref = []; E = 0.0
    for x in pb["X"]:
        for a, b in zip(mr[:-1], mr[1:]):
            ref.append(co.backbone_prior_gradient(x[a:b].ravel(), ll[a:b - 1], ul[a:b - 1], 7.5))
            E -= backbone_log_prob_single(x[a:b], ll[a:b - 1], ul[a:b - 1], 7.5)
    ref = np.concatenate(ref)
    assert np.abs(ref).max() > 0
    assert np.array_equal(out["grad"][0], ref)      # the backbone term uses IEEE ops in both modes
    assert abs(out["energies"][0, 2] - E) <= 1e-12 * abs(E)


@pytest.mark.parametrize("exact", [True, False])
def test_sphere_gradient_bitexact_and_energy(co, exact):
    from oracle.oracle import sphere_log_prob_single
    from ensemble_hic_b200.engine import TERM_SPHERE
    pb = make_problem(seed=14, S=3, N=60, uniform_radii=False, step=1.5)
    c = np.array([0.3, -0.2, 0.1])
    m = _model(pb, exact, sphere=(3.0, 5.0, c))
    out = m.evaluate(pb["X"], pb["norm"], terms=TERM_SPHERE, energies=True)
    r = pb["radii"]
    ref = np.concatenate([co.sphere_prior_gradient(x, c, 3.0, 5.0, np.arange(len(x)), r, r * r) for x in pb["X"]])
    E = -sum(sphere_log_prob_single(x, c, 3.0, 5.0, r) for x in pb["X"])
    assert np.abs(ref).max() > 0
    assert np.array_equal(out["grad"][0], ref)
    assert abs(out["energies"][0, 3] - E) <= 1e-12 * abs(E)


@pytest.mark.parametrize("exact", [True, False])
def test_full_posterior_batched_replicas(co, exact):
    """R replicas with different (x, norm, lammda, beta) in one call == R oracle evaluations."""
    from oracle.oracle import OraclePosterior
    pb = make_problem(seed=21, S=4, N=48, step=0.8)
    N, S = pb["N"], pb["S"]
    ul = pb["radii"][:-1] + pb["radii"][1:]
    sphere = (4.0, 5.0, np.zeros(3))
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], sphere=sphere, nonbonded="nbl")
    R = 5
    rng = np.random.default_rng(3)
    X = np.stack([pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape) for _ in range(R)])
    norm = rng.uniform(1, 5, R); lam = np.linspace(0.0, 1.0, R); bet = np.linspace(0.001, 1.0, R)
    m = _model(pb, exact, sphere=sphere, max_replicas=R)
    out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
    for r in range(R):
        g = op.gradient(X[r], norm[r], lam[r], bet[r])
        assert relerr(out["grad"][r], g) < FAST_TOL
        E = op.energies(X[r], norm[r])
        for q in range(4):
            assert abs(out["energies"][r, q] - E[q]) <= 1e-11 * max(abs(E[q]), 1e-300), (r, q)
        lp = -lam[r] * out["energies"][r, 0] - bet[r] * out["energies"][r, 1] - out["energies"][r, 2] - out["energies"][r, 3]
        assert abs(lp - op.log_prob(X[r], norm[r], lam[r], bet[r])) <= 1e-10 * abs(lp)
        assert relerr(out["md"][r], op.mock_data(X[r], norm[r])) < FAST_TOL


def test_neighbour_list_is_bitexact(co):
    from oracle.oracle import OracleNBList
    from ensemble_hic_b200.engine import nblist_pairs
    pb = make_problem(seed=31, S=1, N=300, step=0.5)
    x = pb["X"][0]
    cutoff = 1.0 + 1e-3
    nb = OracleNBList(co, cutoff, 90, 300, 300)
    n = nb.update(x, 1)
    pairs = nblist_pairs(x, cutoff)
    assert n == len(pairs) and n > 300
    assert np.array_equal(pairs, nb.canonical_pairs())


def test_hmc_trajectory_matches_oracle_leapfrog(co):
    import torch
    from oracle.oracle import OraclePosterior, leapfrog
    pb = make_problem(seed=41, S=2, N=23, radius=1.0, cd_factor=2.0, min_sep=2, step=1.8)
    N, S = pb["N"], pb["S"]
    ul = pb["radii"][:-1] + pb["radii"][1:]
    sphere = (10.0, 5.0, np.zeros(3))
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 1000.0,
                         [np.zeros(N - 1)], [ul], [0, N], sphere=sphere, nonbonded="nbl")
    R, L = 2, 20
    rng = np.random.default_rng(0)
    X = np.stack([pb["X"], pb["X"] + rng.normal(scale=0.1, size=pb["X"].shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0]); eps = np.array([1e-3, 2e-3])
    m = _model(pb, False, k_nb=50.0, k_bb=1000.0, sphere=sphere, max_replicas=R)
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
    m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
    torch.cuda.synchronize()
    H = H_t.cpu().numpy()
    for r in range(R):
        grad = lambda x: op.gradient(x, norm[r], lam[r], bet[r])
        x1, p1 = leapfrog(grad, X[r], P[r], eps[r], L)
        assert relerr(x_t[r].cpu().numpy(), x1) < 1e-11
        assert relerr(p_t[r].cpu().numpy(), p1) < 1e-9
        E0 = -op.log_prob(X[r], norm[r], lam[r], bet[r]); E1 = -op.log_prob(x1, norm[r], lam[r], bet[r])
        assert abs(H[r, 0] - E0) <= 1e-10 * abs(E0)
        assert abs(H[r, 2] - E1) <= 1e-10 * abs(E1)
        assert abs(H[r, 1] - 0.5 * np.sum(P[r] ** 2)) <= 1e-12 * H[r, 1]
        assert abs(H[r, 3] - 0.5 * np.sum(p1 ** 2)) <= 1e-9 * H[r, 3]


# ---- symmetric tile path (dense lists, fast mode) ---------------------------------------------------

@pytest.mark.parametrize("S,N,density,min_sep", [(1, 64, 1.0, 1), (3, 100, 1.0, 1), (4, 257, 1.0, 2), (5, 130, 0.4, 1),
                                                 (7, 300, 0.9, 1), (2, 96, 0.05, 3)])
def test_tile_path_matches_oracle(co, monkeypatch, S, N, density, min_sep):
    from oracle.oracle import OraclePosterior
    monkeypatch.setenv("EHIC_PATH", "tiles")
    pb = make_problem(seed=S * 1000 + N, S=S, N=N, density=density, min_sep=min_sep, uniform_radii=False, step=0.7)
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R = 3
    rng = np.random.default_rng(1)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)])
    norm = rng.uniform(1, 5, R); lam = np.array([0.2, 1.0, 0.6]); bet = np.array([1.0, 0.3, 0.0])
    m = _model(pb, False, max_replicas=R)
    assert m.info()["path"] == "tiles", "tile path not selected"
    out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
    for r in range(R):
        assert relerr(out["md"][r], op.mock_data(X[r], norm[r])) < FAST_TOL
        assert relerr(out["grad"][r], op.gradient(X[r], norm[r], lam[r], bet[r])) < FAST_TOL
        E = op.energies(X[r], norm[r])
        assert abs(out["energies"][r, 0] - E[0]) <= 1e-11 * abs(E[0])
    glik = m.evaluate(X[0], norm[0], terms=1)["grad"][0]
    assert relerr(glik, co.calculate_gradient(X[0], pb["alpha"], norm[0], pb["cds"], pb["data"])) < FAST_TOL
    monkeypatch.setenv("EHIC_PATH", "rows")                      # the two GPU paths agree with each other
    m2 = _model(pb, False, max_replicas=R)
    assert m2.info()["path"] != "tiles"
    out2 = m2.evaluate(X, norm, lam, bet, md=True)
    assert relerr(out["grad"], out2["grad"]) < 1e-13 and relerr(out["md"], out2["md"]) < 1e-13


def test_tile_path_is_chosen_for_dense_lists_only(monkeypatch):
    monkeypatch.delenv("EHIC_PATH", raising=False)
    dense = make_problem(seed=3, S=2, N=256, density=1.0)
    sparse = make_problem(seed=3, S=2, N=256, density=0.1)
    assert _model(dense, False).info()["path"] == "tiles"
    assert _model(sparse, False).info()["path"] != "tiles"
    assert _model(dense, True).info()["path"] != "tiles"           # exact mode keeps the row path (reference order)
    dup = dict(dense)
    dup["data"] = np.concatenate([dense["data"], dense["data"][:1][:, [1, 0, 2]]])     # (j, i) duplicate of a pair
    dup["cds"] = np.concatenate([dense["cds"], dense["cds"][:1]])
    assert _model(dup, False).info()["path"] != "tiles"


def test_tile_path_trajectory(co, monkeypatch):
    import torch
    from oracle.oracle import OraclePosterior, leapfrog
    monkeypatch.setenv("EHIC_PATH", "tiles")
    monkeypatch.setenv("EHIC_TRAJ", "multi")          # the multi-kernel (CUDA graph) trajectory, not the one-CTA kernel
    pb = make_problem(seed=77, S=3, N=140, step=0.8)
    N, S = pb["N"], pb["S"]
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R, L = 2, 6
    rng = np.random.default_rng(0)
    X = np.stack([pb["X"], pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0]); eps = np.array([5e-4, 1e-3])
    m = _model(pb, False, max_replicas=R)
    assert m.info()["path"] == "tiles"
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
    m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
    torch.cuda.synchronize()
    H = H_t.cpu().numpy()
    for r in range(R):
        x1, p1 = leapfrog(lambda x: op.gradient(x, norm[r], lam[r], bet[r]), X[r], P[r], eps[r], L)
        assert relerr(x_t[r].cpu().numpy(), x1) < 1e-11 and relerr(p_t[r].cpu().numpy(), p1) < 1e-9
        E1 = -op.log_prob(x1, norm[r], lam[r], bet[r])
        assert abs(H[r, 2] - E1) <= 1e-10 * abs(E1)
        assert abs(H[r, 3] - 0.5 * np.sum(p1 ** 2)) <= 1e-9 * H[r, 3]


# ---- optional FP32 mode (north star: 1e-5 relative on log-posterior and gradient) ----------------------

FP32_TOL = 1e-5


@pytest.mark.parametrize("S,N,density,shift", [(5, 257, 1.0, 0.0), (7, 300, 0.9, 0.0), (3, 100, 1.0, 300.0), (12, 130, 0.6, 0.0),
                                               (4, 64, 1.0, -1000.0)])
def test_fp32_mode_within_1e5_of_oracle(co, monkeypatch, S, N, density, shift):
    """FP32 pair arithmetic in the tile kernels; `shift` moves the whole ensemble away from the origin
    (the (hi, lo) coordinate split keeps differences accurate)."""
    from oracle.oracle import OraclePosterior
    monkeypatch.setenv("EHIC_TILE_MIN_FILL", "0.2")            # small dense lists fill few of their 128 x 32 tiles
    pb = make_problem(seed=S * 1000 + N, S=S, N=N, density=density, uniform_radii=False, step=0.7)
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R = 3
    rng = np.random.default_rng(1)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)]) + shift
    norm = rng.uniform(1, 5, R); lam = np.array([0.2, 1.0, 0.6]); bet = np.array([1.0, 0.3, 0.0])
    m = _model(pb, False, max_replicas=R, fp32=True)
    assert m.info()["path"] == "tiles" and m.info()["flags"] & 2
    out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
    for r in range(R):
        ref_md = op.mock_data(X[r], norm[r])
        np.testing.assert_allclose(out["md"][r], ref_md, rtol=2e-5)                # element-wise, far pairs included
        assert relerr(out["md"][r], ref_md) < FP32_TOL
        assert relerr(out["grad"][r], op.gradient(X[r], norm[r], lam[r], bet[r])) < FP32_TOL
        E = op.energies(X[r], norm[r])
        assert abs(out["energies"][r, 0] - E[0]) <= FP32_TOL * abs(E[0])             # -log L
        np.testing.assert_allclose(out["energies"][r, 1:], E[1:], rtol=1e-11)        # priors stay FP64
    glik = m.evaluate(X[0], norm[0], terms=1)["grad"][0]
    assert relerr(glik, co.calculate_gradient(X[0], pb["alpha"], norm[0], pb["cds"], pb["data"])) < FP32_TOL
    # and it is a different arithmetic, not the FP64 path under another name
    m64 = _model(pb, False, max_replicas=R)
    g64 = m64.evaluate(X, norm, lam, bet)["grad"]
    assert 1e-9 < relerr(out["grad"], g64) < FP32_TOL


def test_fp32_mode_excludes_exact_and_the_row_path(monkeypatch):
    dense = make_problem(seed=3, S=2, N=128, density=1.0)
    with pytest.raises(Exception, match="exclude"):
        _model(dense, True, fp32=True)
    sparse = make_problem(seed=3, S=2, N=256, density=0.1)
    monkeypatch.setenv("EHIC_PATH", "rows")
    with pytest.raises(Exception, match="no row-path kernels"):
        _model(sparse, False, fp32=True)


@pytest.mark.parametrize("S,N,density,min_sep,shift", [(30, 64, 0.3, 1, 0.0), (33, 50, 0.5, 1, 0.0), (20, 100, 0.1, 2, 300.0),
                                                       (2, 60, 0.4, 1, 0.0), (100, 37, 0.6, 1, -1000.0), (32, 200, 0.02, 3, 0.0)])
def test_fp32_mode_sparse_lists_within_1e5_of_oracle(co, S, N, density, min_sep, shift):
    """sparse lists in FP32 mode run the FP32 lane kernels whatever the ensemble size (S = 2 fills two lanes);
    `shift` moves the ensemble away from the origin"""
    from oracle.oracle import OraclePosterior
    pb = make_problem(seed=S * 1000 + N, S=S, N=N, density=density, min_sep=min_sep, uniform_radii=False, step=0.7)
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R = 3
    rng = np.random.default_rng(1)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)]) + shift
    norm = rng.uniform(1, 5, R); lam = np.array([0.2, 1.0, 0.6]); bet = np.array([1.0, 0.3, 0.0])
    m = _model(pb, False, max_replicas=R, fp32=True)
    assert m.info()["path"] == "lanes" and m.info()["flags"] & 2
    out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
    for r in range(R):
        ref_md = op.mock_data(X[r], norm[r])
        np.testing.assert_allclose(out["md"][r], ref_md, rtol=2e-5)
        assert relerr(out["grad"][r], op.gradient(X[r], norm[r], lam[r], bet[r])) < FP32_TOL
        E = op.energies(X[r], norm[r])
        assert abs(out["energies"][r, 0] - E[0]) <= FP32_TOL * abs(E[0])
        np.testing.assert_allclose(out["energies"][r, 1:], E[1:], rtol=1e-11)
    glik = m.evaluate(X[0], norm[0], terms=1)["grad"][0]
    assert relerr(glik, co.calculate_gradient(X[0], pb["alpha"], norm[0], pb["cds"], pb["data"])) < FP32_TOL
    m64 = _model(pb, False, max_replicas=R)
    g64 = m64.evaluate(X, norm, lam, bet)["grad"]
    assert 1e-9 < relerr(out["grad"], g64) < FP32_TOL


@pytest.mark.parametrize("name", ["hairpin_s", "proteins", "nora2012"])
def test_fp32_mode_on_the_golden_configs(name):
    """the three shipped configurations (all sparse or small): FP32 mode against the reference's own outputs"""
    import os
    from ensemble_hic_b200.engine import Model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    sphere = (float(g["sphere"][0]), float(g["sphere"][1]), np.zeros(3)) if len(g["sphere"]) else None
    m = Model(int(g["S"]), g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]),
              sphere=sphere, fp32=True)
    assert m.info()["path"] in ("lanes", "tiles") and m.info()["flags"] & 2
    out = m.evaluate(g["X"].reshape(1, -1), float(g["norm"]), float(g["lammda"]), float(g["beta"]), energies=True, md=True)
    np.testing.assert_allclose(out["md"][0], g["md"], rtol=2e-5)
    assert relerr(out["grad"][0], g["grad"]) < FP32_TOL
    glik = m.evaluate(g["X"].reshape(1, -1), float(g["norm"]), terms=1)["grad"][0]
    assert relerr(glik, g["grad_lik"]) < FP32_TOL
    assert abs(out["energies"][0, 0] - g["energies"][0]) <= FP32_TOL * abs(g["energies"][0])


def test_fp32_mode_sparse_trajectory(monkeypatch):
    """a leapfrog trajectory on the FP32 lane kernels (captured multi-kernel graph) stays within 1e-5 of the FP64 one"""
    import torch
    pb = make_problem(seed=78, S=20, N=90, density=0.3, step=0.8)
    R, L = 2, 6
    rng = np.random.default_rng(0)
    X = np.stack([pb["X"], pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0]); eps = np.array([5e-4, 1e-3])
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    res = {}
    for fp32 in (False, True):
        m = _model(pb, False, max_replicas=R, fp32=fp32)
        assert m.trajectory_plan(R, L)["mode"] == ("graph" if fp32 else m.trajectory_plan(R, L)["mode"])
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        for rep in range(2):
            m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[fp32] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy())
        m.close()
    assert relerr(res[True][0], res[False][0]) < 1e-8
    assert relerr(res[True][1], res[False][1]) < FP32_TOL
    np.testing.assert_allclose(res[True][2], res[False][2], rtol=FP32_TOL)
    assert not np.array_equal(res[True][1], res[False][1])


def test_fp32_mode_trajectory(monkeypatch):
    """a leapfrog trajectory in FP32 mode stays within 1e-5 of the FP64 trajectory over a few steps"""
    import torch
    monkeypatch.setenv("EHIC_TRAJ", "multi")
    monkeypatch.setenv("EHIC_TILE_MIN_FILL", "0.2")
    pb = make_problem(seed=77, S=3, N=140, step=0.8)
    R, L = 2, 6
    rng = np.random.default_rng(0)
    X = np.stack([pb["X"], pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0]); eps = np.array([5e-4, 1e-3])
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    res = {}
    for fp32 in (False, True):
        m = _model(pb, False, max_replicas=R, fp32=fp32)
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[fp32] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy())
        m.close()
    assert relerr(res[True][0], res[False][0]) < 1e-8          # positions move by eps^2 * gradient error
    assert relerr(res[True][1], res[False][1]) < FP32_TOL
    np.testing.assert_allclose(res[True][2], res[False][2], rtol=FP32_TOL)
    assert not np.array_equal(res[True][1], res[False][1])


# ---- lane path (sparse lists, fast mode: lane <-> structure) -------------------------------------------

@pytest.mark.parametrize("S,N,density,min_sep", [(30, 64, 0.3, 1), (33, 50, 0.5, 1), (64, 40, 1.0, 1), (20, 100, 0.1, 2),
                                                 (1, 30, 0.5, 1), (100, 37, 0.6, 1), (32, 200, 0.02, 3)])
def test_lane_path_matches_oracle(co, monkeypatch, S, N, density, min_sep):
    from oracle.oracle import OraclePosterior
    from ensemble_hic_b200.engine import TERM_ALL
    monkeypatch.setenv("EHIC_PATH", "lanes")
    pb = make_problem(seed=S * 1000 + N, S=S, N=N, density=density, min_sep=min_sep, uniform_radii=False, step=0.7)
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R = 3
    rng = np.random.default_rng(1)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)])
    norm = rng.uniform(1, 5, R); lam = np.array([0.2, 1.0, 0.6]); bet = np.array([1.0, 0.3, 0.0])
    m = _model(pb, False, max_replicas=R)
    assert m.info()["path"] == "lanes", "lane path not selected"
    out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
    for r in range(R):
        assert relerr(out["md"][r], op.mock_data(X[r], norm[r])) < FAST_TOL
        assert relerr(out["grad"][r], op.gradient(X[r], norm[r], lam[r], bet[r])) < FAST_TOL
        E = op.energies(X[r], norm[r])
        assert abs(out["energies"][r, 0] - E[0]) <= 1e-11 * abs(E[0])
    glik = m.evaluate(X[0], norm[0], terms=1)["grad"][0]
    assert relerr(glik, co.calculate_gradient(X[0], pb["alpha"], norm[0], pb["cds"], pb["data"])) < FAST_TOL
    gprior = m.evaluate(X[0], norm[0], terms=TERM_ALL & ~1)["grad"][0]       # no contact term: epilogue only
    assert relerr(gprior, op.gradient(X[0], norm[0], 0.0, 1.0)) < FAST_TOL
    monkeypatch.setenv("EHIC_PATH", "rows")                      # the two GPU paths agree with each other
    m2 = _model(pb, False, max_replicas=R)
    assert m2.info()["path"] == "rows"
    out2 = m2.evaluate(X, norm, lam, bet, md=True)
    assert relerr(out["grad"], out2["grad"]) < 1e-13 and relerr(out["md"], out2["md"]) < 1e-13


def test_lane_path_reversed_and_duplicate_pairs(co, monkeypatch):
    """data points given as (j, i) with j > i, and the same bead pair listed twice (allowed by the reference:
    every row of data_points is its own Poisson term)"""
    monkeypatch.setenv("EHIC_PATH", "lanes")
    pb = make_problem(seed=5, S=32, N=45, density=0.4)
    data, cds = pb["data"].copy(), pb["cds"].copy()
    data[::3, :2] = data[::3, 1::-1]                            # every third pair reversed
    data = np.concatenate([data, data[:40]]); cds = np.concatenate([cds, cds[:40] * 1.1])     # 40 duplicates, other dc
    pb2 = dict(pb, data=np.ascontiguousarray(data), cds=np.ascontiguousarray(cds))
    m = _model(pb2, False)
    assert m.info()["path"] == "lanes"
    out = m.evaluate(pb["X"], pb["norm"], terms=1, md=True)
    assert relerr(out["md"][0], co.ensemble_contacts_evaluate(pb["X"], pb["norm"], cds, pb["alpha"], data)) < FAST_TOL
    assert relerr(out["grad"][0], co.calculate_gradient(pb["X"], pb["alpha"], pb["norm"], cds, data)) < FAST_TOL


@pytest.mark.parametrize("fp32", [False, True])
def test_lane_replica_resident_launches_equal_the_plain_grid(monkeypatch, fp32):
    """replica-resident launches (one CTA per SM walking a contiguous, work-balanced range of a replica's virtual
    blocks so that its lane copy stays in L1) run the same bodies as the plain grid: results are identical bit for bit"""
    import torch
    pb = make_problem(seed=5, S=20, N=400, density=0.2, min_sep=2, step=0.8)
    R, L = 24, 3
    rng = np.random.default_rng(2)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = rng.uniform(1, 5, R); lam = np.linspace(0.1, 1.0, R); bet = np.linspace(1.0, 0.1, R); eps = np.full(R, 1e-3)
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    monkeypatch.setenv("EHIC_PATH", "lanes")
    monkeypatch.setenv("EHIC_TRAJ", "multi")
    monkeypatch.setenv("EHIC_LANE_SPLIT", "0")           # the resident backward launch walks the plain work list
    res = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("EHIC_LANE_RESIDENT", mode)
        m = _model(pb, False, max_replicas=R, fp32=fp32)
        out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[mode] = [out["grad"], out["md"], out["energies"], x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy()]
        m.close()
    for a, b in zip(res["0"], res["2"]):
        assert np.isfinite(a).all() and np.array_equal(a, b)


@pytest.mark.parametrize("fp32", [False, True])
def test_lane_split_rows_equal_whole_rows(co, monkeypatch, fp32):
    """backward lane kernels: long rows cut into 2 or 4 segments over the warps of a CTA (the work list of launches with
    few replicas) against one warp per bead -- same terms, partial sums combined in segment order"""
    pb = make_problem(seed=8, S=30, N=260, density=0.5, min_sep=1, step=0.8)       # rows of ~130 entries: 4 segments
    R = 3
    rng = np.random.default_rng(2)
    X = np.stack([pb["X"] + rng.normal(scale=0.03, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    norm = rng.uniform(1, 5, R); lam = np.array([0.2, 1.0, 0.6]); bet = np.array([1.0, 0.3, 0.0])
    monkeypatch.setenv("EHIC_PATH", "lanes")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EHIC_LANE_SPLIT", mode)
        m = _model(pb, False, max_replicas=R, fp32=fp32)
        res[mode] = m.evaluate(X, norm, lam, bet, energies=True)
        m.close()
    assert not np.array_equal(res["0"]["grad"], res["1"]["grad"])                  # a different summation order
    assert relerr(res["1"]["grad"], res["0"]["grad"]) < (1e-6 if fp32 else 1e-13)
    np.testing.assert_array_equal(res["1"]["energies"], res["0"]["energies"])
    g = co.calculate_gradient(X[0].reshape(pb["X"].shape), pb["alpha"], norm[0], pb["cds"], pb["data"])
    monkeypatch.setenv("EHIC_LANE_SPLIT", "1")
    m = _model(pb, False, max_replicas=R, fp32=fp32)
    glik = m.evaluate(X[0], norm[0], terms=1)["grad"][0]
    assert relerr(glik, g) < (FP32_TOL if fp32 else 1e-12)
    m.close()


def test_lane_path_is_chosen_for_sparse_lists_with_enough_structures(monkeypatch):
    monkeypatch.delenv("EHIC_PATH", raising=False)
    assert _model(make_problem(seed=3, S=30, N=256, density=0.1), False).info()["path"] == "lanes"
    assert _model(make_problem(seed=3, S=40, N=64, density=0.1), False).info()["path"] == "lanes"     # 40 of 64 lanes
    assert _model(make_problem(seed=3, S=36, N=64, density=0.1), False).info()["path"] == "rows"      # 36 of 64 lanes
    assert _model(make_problem(seed=3, S=2, N=256, density=0.1), False).info()["path"] == "rows"
    assert _model(make_problem(seed=3, S=30, N=256, density=0.1), True).info()["path"] == "rows"      # exact: reference order
    assert _model(make_problem(seed=3, S=30, N=256, density=1.0), False).info()["path"] == "tiles"


def test_lane_path_trajectory(co, monkeypatch):
    import torch
    from oracle.oracle import OraclePosterior, leapfrog
    monkeypatch.setenv("EHIC_PATH", "lanes")
    monkeypatch.setenv("EHIC_TRAJ", "multi")          # the multi-kernel (CUDA graph) trajectory, not the one-CTA kernel
    pb = make_problem(seed=78, S=35, N=61, density=0.3, step=0.8)
    N, S = pb["N"], pb["S"]
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 500.0,
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="n2")
    R, L = 2, 6
    rng = np.random.default_rng(0)
    X = np.stack([pb["X"], pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0]); eps = np.array([5e-4, 1e-3])
    m = _model(pb, False, max_replicas=R)
    assert m.info()["path"] == "lanes"
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
    m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
    torch.cuda.synchronize()
    H = H_t.cpu().numpy()
    for r in range(R):
        x1, p1 = leapfrog(lambda x: op.gradient(x, norm[r], lam[r], bet[r]), X[r], P[r], eps[r], L)
        assert relerr(x_t[r].cpu().numpy(), x1) < 1e-11 and relerr(p_t[r].cpu().numpy(), p1) < 1e-9
        E0 = -op.log_prob(X[r], norm[r], lam[r], bet[r]); E1 = -op.log_prob(x1, norm[r], lam[r], bet[r])
        assert abs(H[r, 0] - E0) <= 1e-10 * abs(E0) and abs(H[r, 2] - E1) <= 1e-10 * abs(E1)
        assert abs(H[r, 1] - 0.5 * np.sum(P[r] ** 2)) <= 1e-12 * H[r, 1]
        assert abs(H[r, 3] - 0.5 * np.sum(p1 ** 2)) <= 1e-9 * H[r, 3]


@pytest.mark.parametrize("S,N,density,sphere", [(2, 23, 1.0, True), (2, 56, 1.0, True), (10, 62, 0.5, False), (1, 40, 0.3, False)])
def test_one_launch_trajectory_equals_multi_kernel_trajectory(monkeypatch, S, N, density, sphere):
    """small systems: whole trajectory in one kernel (state in shared memory) vs the kernel-per-phase path"""
    import torch
    pb = make_problem(seed=S + N, S=S, N=N, density=density, radius=1.0, cd_factor=2.0, min_sep=2, step=1.8)
    R, L = 3, 25
    rng = np.random.default_rng(4)
    X = np.stack([pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5, 1.5]); lam = np.array([0.3, 1.0, 0.0]); bet = np.array([0.5, 1.0, 0.01]); eps = np.array([1e-3, 2e-3, 5e-3])
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    res = {}
    for mode in ("multi", "small"):
        if mode == "multi":
            monkeypatch.setenv("EHIC_TRAJ", "multi")
        else:
            monkeypatch.delenv("EHIC_TRAJ", raising=False)
        m = _model(pb, False, k_nb=50.0, k_bb=1000.0, sphere=(10.0, 5.0, np.zeros(3)) if sphere else None, max_replicas=R)
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        from ensemble_hic_b200.engine import launch_count
        l0 = launch_count()
        m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[mode] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy(), launch_count() - l0)
        m.close()
    assert res["small"][3] == 1 and res["multi"][3] > 3 * L
    assert relerr(res["small"][0], res["multi"][0]) < 1e-12
    assert relerr(res["small"][1], res["multi"][1]) < 1e-10
    np.testing.assert_allclose(res["small"][2], res["multi"][2], rtol=1e-11)


@pytest.mark.parametrize("S,N,density,sphere,C,skin", [
    (30, 70, 0.35, False, 2, None), (30, 70, 0.35, False, 4, None), (30, 70, 0.35, True, 8, None),
    (9, 100, 0.2, True, 4, "0.02"),           # lanes per structure group > structures per CTA; tiny skin: rebuilds inside the launch
    (5, 64, 0.5, False, 2, None),             # S not divisible: the last CTA holds fewer structures
    (13, 40, 1.0, False, 8, None)])
def test_cluster_trajectory_equals_multi_kernel_trajectory_and_oracle(monkeypatch, S, N, density, sphere, C, skin):
    """medium systems: the whole trajectory in ONE launch, one thread-block cluster per replica, the structures split
    over the cluster's CTAs and the ensemble sums exchanged through distributed shared memory -- against the
    kernel-per-phase path and against the oracle's leapfrog (reference kernels)"""
    import torch
    from oracle.oracle import OraclePosterior, leapfrog
    from ensemble_hic_b200.engine import launch_count
    pb = make_problem(seed=S + N + C, S=S, N=N, density=density, radius=1.0, cd_factor=2.0, min_sep=2, step=1.8)
    if skin is not None:
        monkeypatch.setenv("EHIC_VERLET_SKIN", skin)
    R, L = 3, 12
    rng = np.random.default_rng(4)
    X = np.stack([pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5, 1.5]); lam = np.array([0.3, 1.0, 0.0]); bet = np.array([0.5, 1.0, 0.01]); eps = np.array([1e-3, 2e-3, 5e-3])
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    sph = (10.0, 5.0, np.zeros(3)) if sphere else None
    res = {}
    for mode in ("multi", "cluster"):
        monkeypatch.setenv("EHIC_TRAJ", mode)
        monkeypatch.setenv("EHIC_TRAJ_CLUSTER", str(C))
        m = _model(pb, False, k_nb=50.0, k_bb=1000.0, sphere=sph, max_replicas=R)
        plan = m.trajectory_plan(R, L)
        assert plan["mode"] == ("graph" if mode == "multi" else "cluster"), plan
        if mode == "cluster":
            assert plan["ctas_per_replica"] == C and plan["smem_bytes"] <= 227 * 1024
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        l0 = launch_count()
        for rep in range(2):                                  # twice: the second call starts from the end point of the first
            m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[mode] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy(), launch_count() - l0)
        if mode == "cluster":                                 # the same model still evaluates through the other entry points
            g = m.evaluate(X, norm, lam, bet)["grad"]
            assert np.isfinite(g).all()
        m.close()
    assert res["cluster"][3] == 2 and res["multi"][3] > 3 * L
    assert relerr(res["cluster"][0], res["multi"][0]) < 1e-12
    assert relerr(res["cluster"][1], res["multi"][1]) < 1e-10
    np.testing.assert_allclose(res["cluster"][2], res["multi"][2], rtol=1e-11)
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S, pb["data"], pb["cds"], pb["alpha"], pb["radii"], 50.0, 1000.0, [np.zeros(N - 1)], [ul], [0, N],
                         sphere=sph, nonbonded="n2")
    r = 1
    x1, p1 = leapfrog(lambda x: op.gradient(x, norm[r], lam[r], bet[r]), X[r], P[r], eps[r], L)
    x2, p2 = leapfrog(lambda x: op.gradient(x, norm[r], lam[r], bet[r]), x1, p1, eps[r], L)
    assert relerr(res["cluster"][0][r], x2) < 1e-11 and relerr(res["cluster"][1][r], p2) < 1e-9
    E0 = -op.log_prob(x1, norm[r], lam[r], bet[r]); E1 = -op.log_prob(x2, norm[r], lam[r], bet[r])
    H = res["cluster"][2]
    assert abs(H[r, 0] - E0) <= 1e-10 * abs(E0) and abs(H[r, 2] - E1) <= 1e-10 * abs(E1)
    assert abs(H[r, 1] - 0.5 * np.sum(p1 ** 2)) <= 1e-9 * H[r, 1] and abs(H[r, 3] - 0.5 * np.sum(p2 ** 2)) <= 1e-9 * H[r, 3]


def test_cluster_trajectory_on_the_nora2012_shape():
    """308 beads x 30 structures, M = 10 145 (config_files/nora2012/30structures_3kb.cfg): 222 KB of coordinates do not fit
    one SM's shared memory beside the weights.  EHIC_TRAJ=cluster runs a whole trajectory of that shape as ONE launch
    (a cluster of 2, 4 or 8 CTAs per replica) and reproduces the multi-kernel trajectory; without the override the
    captured multi-kernel graph stays the choice, because it measured faster at every batch size"""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nora2012.npz"))
    from ensemble_hic_b200.engine import Model, launch_count
    R, L = 5, 8
    kw = dict(alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]), max_replicas=302)
    rng = np.random.default_rng(9)
    X = np.stack([g["X"] + 0.01 * rng.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
    lam = np.linspace(0.2, 1.0, R); eps = np.full(R, 1e-3); nrm = np.full(R, float(g["norm"]))
    res = {}
    for mode in ("auto", "cluster", "multi"):
        if mode != "auto":
            os.environ["EHIC_TRAJ"] = mode
        try:
            m = Model(int(g["S"]), g["radii"], g["data"], g["cds"], **kw)
        finally:
            os.environ.pop("EHIC_TRAJ", None)
        for Rq in (302, 38):
            plan = m.trajectory_plan(Rq, 100)
            if mode == "cluster":
                assert plan["mode"] == "cluster" and plan["ctas_per_replica"] in (2, 4, 8) and plan["launches"] == 1, plan
            else:
                assert plan["mode"] == "graph", plan
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        l0 = launch_count()
        m.hmc_trajectory_device(R, L, x_t, p_t, t(nrm), t(lam), t(lam), t(eps), H_t)
        torch.cuda.synchronize()
        res[mode] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy(), launch_count() - l0)
        m.close()
    assert res["cluster"][3] == 1 and res["auto"][3] > 3 * L
    for mode in ("auto", "cluster"):
        assert relerr(res[mode][0], res["multi"][0]) < 1e-12 and relerr(res[mode][1], res["multi"][1]) < 1e-10
        np.testing.assert_allclose(res[mode][2], res["multi"][2], rtol=1e-11)


# ---- Verlet lists inside trajectories ---------------------------------------------------------------

@pytest.mark.parametrize("scenario", ["chain", "tiny_skin", "overflow"])
def test_verlet_list_trajectory_equals_all_pairs_trajectory(monkeypatch, scenario):
    """the nonbonded term of a multi-kernel trajectory comes from Verlet lists reused across leapfrog steps;
    the trajectory must be the one the all-pairs sweep gives (rebuilds forced by a tiny skin; a ball of
    beads with more than 64 neighbours each exercises the overflow fallback)"""
    import torch
    from ensemble_hic_b200.engine import launch_count
    monkeypatch.setenv("EHIC_TRAJ", "multi")
    monkeypatch.setenv("EHIC_PATH", "rows")
    if scenario == "tiny_skin":
        monkeypatch.setenv("EHIC_VERLET_SKIN", "0.02")
    S, N = 3, 150
    pb = make_problem(seed=11, S=S, N=N, density=0.2, step=0.8)
    X0 = pb["X"]
    if scenario == "overflow":
        X0 = np.stack([_ball(np.random.default_rng(k), N, 1.2) for k in range(S)])     # everybody touches everybody
    R, L = 2, 60
    rng = np.random.default_rng(0)
    X = np.stack([X0, X0 + rng.normal(scale=0.02, size=X0.shape)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = np.array([2.5, 3.5]); lam = np.array([0.3, 1.0]); bet = np.array([0.5, 1.0])
    eps = np.array([2e-3, 4e-3]) if scenario != "overflow" else np.array([2e-4, 4e-4])
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    res = {}
    for mode in ("fused", "kernels", "off"):             # one verlet_step_kernel per step / check + build + force kernels / all-pairs
        monkeypatch.setenv("EHIC_VERLET", "0" if mode == "off" else "1")
        monkeypatch.setenv("EHIC_VERLET_FUSED", "0" if mode == "kernels" else "1")
        m = _model(pb, False, max_replicas=R)
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        l0 = launch_count()
        for _ in range(2):                               # second trajectory starts from a list built in the first
            m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[mode] = (x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy(), launch_count() - l0)
        m.close()
    assert res["kernels"][3] > res["off"][3] > res["fused"][3]      # the list kernels did run; the fused step is one launch
    for mode in ("fused", "kernels"):
        assert relerr(res[mode][0], res["off"][0]) < 1e-12
        assert relerr(res[mode][1], res["off"][1]) < 1e-10
        np.testing.assert_allclose(res[mode][2], res["off"][2], rtol=1e-11)
        assert np.all(np.isfinite(res[mode][2]))
    assert np.array_equal(res["fused"][0], res["kernels"][0])      # same lists, same order, same arithmetic


def test_verlet_lists_survive_changing_batch_sizes(monkeypatch):
    """the lists belong to replica slots: calls with fewer replicas than max_replicas, then more, then other
    coordinates in the same slots, must all give the all-pairs result"""
    import torch
    monkeypatch.setenv("EHIC_TRAJ", "multi")
    pb = make_problem(seed=21, S=4, N=90, density=0.3, step=0.8)
    rng = np.random.default_rng(3)
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
    calls = []
    for R in (2, 4, 3, 4):
        X = np.stack([pb["X"] + rng.normal(scale=0.3, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
        calls.append((R, X, rng.normal(size=X.shape)))
    res = {}
    for verlet in ("1", "0"):
        monkeypatch.setenv("EHIC_VERLET", verlet)
        m = _model(pb, False, max_replicas=4)
        out = []
        for R, X, P in calls:
            x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
            m.hmc_trajectory_device(R, 8, x_t, p_t, t(np.full(R, 3.0)), t(np.linspace(0.2, 1, R)), t(np.linspace(1, 0.2, R)),
                                    t(np.full(R, 2e-3)), H_t)
            torch.cuda.synchronize()
            out.append((x_t.cpu().numpy(), p_t.cpu().numpy(), H_t.cpu().numpy()))
        res[verlet] = out
        m.close()
    for a, b in zip(res["1"], res["0"]):
        assert relerr(a[0], b[0]) < 1e-12 and relerr(a[1], b[1]) < 1e-10
        np.testing.assert_allclose(a[2], b[2], rtol=1e-11)


@pytest.mark.parametrize("path,S,N,density", [("rows", 3, 70, 0.4), ("lanes", 33, 48, 0.4), ("tiles", 4, 160, 1.0)])
def test_split_batches_equal_single_batches(monkeypatch, path, S, N, density):
    """large batches are evaluated / integrated as two halves on two streams in their own workspace slots
    (EHIC_SPLIT, automatic for heavy models): same results as one batch, for evaluations and trajectories"""
    import torch
    monkeypatch.setenv("EHIC_PATH", path)
    monkeypatch.setenv("EHIC_TRAJ", "multi")
    monkeypatch.setenv("EHIC_TILE_MIN_FILL", "0.2")
    pb = make_problem(seed=31, S=S, N=N, density=density, step=0.8)
    R, L = 5, 7
    rng = np.random.default_rng(2)
    X = np.stack([pb["X"] + rng.normal(scale=0.05, size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    P = rng.normal(size=X.shape)
    norm = rng.uniform(2, 4, R); lam = np.linspace(0.2, 1, R); bet = np.linspace(1, 0.3, R); eps = np.full(R, 1e-3)
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
    res = {}
    for split in ("1", "2"):
        monkeypatch.setenv("EHIC_SPLIT", split)
        m = _model(pb, False, max_replicas=R)
        assert m.info()["path"] == path
        g = torch.zeros(R, X.shape[1], dtype=torch.float64, device=dev); E = torch.zeros(R, 4, dtype=torch.float64, device=dev)
        md = torch.zeros(R, len(pb["data"]), dtype=torch.float64, device=dev)
        m.evaluate_device(R, t(X), t(norm), t(lam), t(bet), grad=g, energies=E, md=md)
        x_t, p_t, H_t = t(X), t(P), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        for _ in range(2):
            m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(eps), H_t)
        torch.cuda.synchronize()
        res[split] = [a.cpu().numpy() for a in (g, E, md, x_t, p_t, H_t)]
        m.close()
    for a, b in zip(res["2"], res["1"]):
        assert np.array_equal(a, b)          # the same kernels on the same data: bit-identical


# ---- cell-grid nonbonded path ------------------------------------------------------------------------

def _ball(rng, n, radius):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * (radius * rng.random(n) ** (1 / 3.0))[:, None]


@pytest.mark.parametrize("N,force", [(4100, False), (1100, True), (150, True), (64, True), (333, True)])
def test_cell_grid_nonbonded_matches_oracle(co, monkeypatch, N, force):
    """compact structures take the cell grid, extended ones the pruned all-pairs sweep -- in the same call"""
    from ensemble_hic_b200.engine import TERM_NONBONDED, TERM_ALL
    if force:
        monkeypatch.setenv("EHIC_NB", "cells")
    rng = np.random.default_rng(N)
    radii = rng.uniform(0.4, 0.6, N)
    compact = _ball(rng, N, 0.62 * N ** (1 / 3.0))                       # ~1 bead per unit volume: many contacts
    extended = make_problem(seed=N, S=1, N=N, step=0.9)["X"][0]           # random walk: grid would not fit -> all-pairs
    walk = make_problem(seed=N + 1, S=1, N=N, step=0.45)["X"][0]
    degenerate = np.zeros((N, 3)); degenerate[:, 0] = np.arange(N) * 0.8  # a straight line: 1 x 1 x many cells
    X = np.stack([compact, extended, walk, degenerate])
    from ensemble_hic_b200.engine import Model
    m = Model(4, radii, k_nb=50.0, k_bb=0.0, max_replicas=2)
    Xr = np.stack([X, X[::-1]])
    out = m.evaluate(Xr, None, None, np.array([1.0, 0.3]), terms=TERM_NONBONDED, energies=True)
    ref_g = np.concatenate([co.forcefield_gradient(x, radii, radii * radii, 50.0) for x in X]).reshape(4, -1)
    ref_E = np.array([co.forcefield_energy(x, radii, radii * radii, 50.0) for x in X])
    assert (np.abs(ref_g).max(axis=1) > 0).all()
    assert relerr(out["grad"][0].reshape(4, -1), ref_g) < FAST_TOL
    assert relerr(out["grad"][1].reshape(4, -1), 0.3 * ref_g[::-1]) < FAST_TOL
    assert abs(out["energies"][0, 1] - ref_E.sum()) <= 1e-12 * ref_E.sum()
    assert abs(out["energies"][1, 1] - ref_E.sum()) <= 1e-12 * ref_E.sum()
    # the two nonbonded paths agree with each other, with backbone fused in
    monkeypatch.setenv("EHIC_NB", "pairs")
    m2 = Model(4, radii, k_nb=50.0, k_bb=500.0, max_replicas=2)
    monkeypatch.setenv("EHIC_NB", "cells")
    m3 = Model(4, radii, k_nb=50.0, k_bb=500.0, max_replicas=2)
    a = m2.evaluate(Xr, None, None, np.array([1.0, 0.3]), terms=TERM_ALL, energies=True)
    b = m3.evaluate(Xr, None, None, np.array([1.0, 0.3]), terms=TERM_ALL, energies=True)
    assert relerr(a["grad"], b["grad"]) < 1e-13
    np.testing.assert_allclose(a["energies"], b["energies"], rtol=1e-12)
