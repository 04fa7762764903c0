"""Parity at BASELINE.json's full sizes (config D: 2000 beads x 100 structures, dense list, M = 1 997 001; and a
10 000-bead sparse case of config E's shape, reduced in structures): directly against the oracle -- the reference's
own compiled Cython (oracle/_ref) when it was built, else its bit-identical C restatement -- at N = 2000,
M = 1 997 001 with S = 25 (and S = 100 when the host has the memory for the reference's 2 x S x M scratch), plus
size-independent properties of the posterior evaluated on the GPU."""
import numpy as np
import pytest

from helpers import random_walk, relerr

pytestmark = pytest.mark.gpu


def _config_d(S=100, N=2000, seed=0):
    rng = np.random.default_rng(seed)
    i, j = np.triu_indices(N, 2)
    data = np.stack([i, j, rng.poisson(2.0, size=len(i))], 1).astype(np.int64)
    p = rng.permutation(len(data))                       # parse_data leaves the rows sorted by count, not by (i, j)
    return dict(S=S, N=N, data=np.ascontiguousarray(data[p]), cds=np.full(len(i), 1.25), radii=np.full(N, 0.5),
                X=random_walk(rng, S, N, 0.9), rng=rng)


@pytest.fixture(scope="module")
def cfg_d():
    return _config_d()


def _model(pb, monkeypatch=None, path=None, exact=False, R=1, fp32=False):
    from ensemble_hic_b200.engine import Model
    import os
    if path is None:
        os.environ.pop("EHIC_PATH", None)
    else:
        os.environ["EHIC_PATH"] = path
    try:
        return Model(pb["S"], pb["radii"], pb["data"], pb["cds"], alpha=10.0, k_nb=500.0, k_bb=500.0,
                     exact=exact, max_replicas=R, fp32=fp32)
    finally:
        os.environ.pop("EHIC_PATH", None)


def _oracle(pb, S, nonbonded):
    from oracle.oracle import OraclePosterior, load_ref
    N = pb["N"]
    ul = pb["radii"][:-1] + pb["radii"][1:]
    return OraclePosterior(S, pb["data"], pb["cds"], 10.0, pb["radii"], 500.0, 500.0, [np.zeros(N - 1)], [ul], [0, N],
                           kernels=load_ref(), nonbonded=nonbonded)


def _host_can_hold(n_bytes):
    try:
        import psutil
        return psutil.virtual_memory().available > 2 * n_bytes
    except Exception:
        return False


@pytest.mark.parametrize("S", [25, 100])
def test_config_d_against_the_reference_cython(cfg_d, S):
    """headline configuration, Model.evaluate against the reference's arithmetic on identical inputs
    (likelihoods_c.pyx:27-75, forward_models_c.pyx:14-32, forcefield_c.pyx:11-59, backbone_prior_c.pyx:11-44):
    gradient 1e-10 max-norm relative, log-posterior 1e-10, mock data element-wise; the exact mode bit for bit"""
    from ensemble_hic_b200 import _lib
    pb = dict(cfg_d, S=S, X=np.ascontiguousarray(cfg_d["X"][:S]))
    M = len(pb["data"])
    if S > 25 and not _host_can_hold(2 * S * M * 8):
        pytest.skip("host memory too small for the reference's 2 x S x M scratch arrays at S = %d" % S)
    op = _oracle(pb, S, "n2")
    x, norm, lam, bet = pb["X"].ravel(), 3.0, 0.7, 0.4
    g_ref = op.gradient(x, norm, lam, bet)
    md_ref = op.mock_data(x, norm)
    U = op.energies(x, norm)
    lp_ref = -lam * U[0] - bet * U[1] - U[2] - U[3]
    for exact in (False, True):
        m = _model(pb, exact=exact)
        assert m.info()["path"] == ("rows" if exact else "tiles")
        out = m.evaluate(x, norm, lam, bet, energies=True, md=True)
        assert relerr(out["grad"][0], g_ref) < 1e-10
        # element-wise; for pairs many contact distances apart the reference's own 0.5 (1 + t/q) cancels, so both
        # sides carry an ABSOLUTE rounding error of ~1e-16 * norm per structure (conditioning caveat)
        np.testing.assert_allclose(out["md"][0], md_ref, rtol=1e-10, atol=4e-16 * norm * S)
        E = out["energies"][0]
        lp = -lam * E[0] - bet * E[1] - E[2] - E[3]
        assert abs(lp - lp_ref) <= 1e-10 * abs(lp_ref)
        np.testing.assert_allclose(E[:3], U[:3], rtol=1e-10)
        if exact:          # reference operation order and IEEE sqrt / div: the Cython entry points bit for bit
            assert np.array_equal(out["md"][0], md_ref)
            gl = m.evaluate(x, norm, 1.0, 1.0, terms=_lib.TERM_LIKELIHOOD)["grad"][0]
            assert np.array_equal(gl, op.gradient(x, norm, 1.0, 1.0, terms=("lik",)))
        m.close()


def test_config_d_paths_and_modes_agree(cfg_d):
    """tile path == lane path == row path == exact (reference-order, IEEE) arithmetic at full size;
    the optional FP32 mode stays within its 1e-5"""
    pb = cfg_d
    tiles, lanes, rows, exact = _model(pb), _model(pb, path="lanes"), _model(pb, path="rows"), _model(pb, exact=True)
    fp32 = _model(pb, fp32=True)
    assert tiles.info()["path"] == "tiles" and lanes.info()["path"] == "lanes" and rows.info()["path"] == "rows"
    out = {}
    for name, m in (("tiles", tiles), ("lanes", lanes), ("rows", rows), ("exact", exact), ("fp32", fp32)):
        out[name] = m.evaluate(pb["X"], 3.0, 0.7, 0.4, energies=True, md=True)
        m.close()
    for a in ("tiles", "lanes", "rows"):
        assert relerr(out[a]["grad"], out["exact"]["grad"]) < 1e-12
        assert relerr(out[a]["md"], out["exact"]["md"]) < 1e-12
        np.testing.assert_allclose(out[a]["energies"], out["exact"]["energies"], rtol=1e-11)
    assert relerr(out["fp32"]["grad"], out["exact"]["grad"]) < 1e-5
    np.testing.assert_allclose(out["fp32"]["md"], out["exact"]["md"], rtol=2e-5)
    np.testing.assert_allclose(out["fp32"]["energies"], out["exact"]["energies"], rtol=1e-5)


def test_config_d_invariances_and_directional_derivative(cfg_d):
    pb = cfg_d
    m = _model(pb, R=2)
    X = pb["X"]
    rng = np.random.default_rng(1)
    base = m.evaluate(X, 3.0, 0.7, 0.4, energies=True)
    g = base["grad"][0].reshape(pb["S"], pb["N"], 3)
    # every term is a function of bead-bead distances: no net force / torque on any structure
    assert np.abs(g.sum(1)).max() < 1e-9 * np.abs(g).max() * pb["N"] ** 0.5
    torque = np.cross(X, g).sum(1)
    assert np.abs(torque).max() < 1e-8 * np.abs(np.cross(X, g)).max() * pb["N"] ** 0.5
    # rigid motion of every structure leaves the energies unchanged; the gradient co-rotates
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    Xr = X @ q.T + rng.normal(size=3)
    rot = m.evaluate(Xr, 3.0, 0.7, 0.4, energies=True)
    np.testing.assert_allclose(rot["energies"], base["energies"], rtol=1e-10)
    assert relerr(rot["grad"][0].reshape(g.shape), g @ q.T) < 1e-9
    # permuting the ensemble members permutes the gradient and keeps the energies
    perm = rng.permutation(pb["S"])
    pe = m.evaluate(X[perm], 3.0, 0.7, 0.4, energies=True)
    np.testing.assert_allclose(pe["energies"], base["energies"], rtol=1e-11)
    assert relerr(pe["grad"][0].reshape(g.shape), g[perm]) < 1e-11
    # directional derivative of -log p along a random direction == <grad, v>
    v = rng.normal(size=X.shape); v /= np.linalg.norm(v)
    h = 1e-5
    two = m.evaluate(np.stack([X + h * v, X - h * v]), 3.0, 0.7, 0.4, grad=False, energies=True)["energies"]
    U = lambda E: 0.7 * E[0] + 0.4 * E[1] + E[2] + E[3]
    fd = (U(two[0]) - U(two[1])) / (2 * h)
    an = float(np.sum(base["grad"][0] * v.ravel()))
    # finite-difference noise floor: energies carry ~1e-15 relative rounding, divided by 2h
    assert abs(fd - an) <= 1e-5 * abs(an) + 1e-14 * abs(U(two[0])) / h
    # tempering is linear in (lammda, beta)
    g00 = m.evaluate(X, 3.0, 0.0, 0.0)["grad"][0]; g10 = m.evaluate(X, 3.0, 1.0, 0.0)["grad"][0]
    g01 = m.evaluate(X, 3.0, 0.0, 1.0)["grad"][0]
    assert relerr(base["grad"][0], g00 + 0.7 * (g10 - g00) + 0.4 * (g01 - g00)) < 1e-11
    m.close()


def test_config_d_trajectory_is_reversible(cfg_d):
    """leapfrog is time-reversible: integrating forward, flipping the momenta and integrating again
    returns to the start (checks the whole-trajectory path at full size)"""
    import torch
    pb = cfg_d
    m = _model(pb)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(2)
    x0 = torch.tensor(pb["X"].reshape(1, -1), device=dev)
    p0 = torch.tensor(rng.normal(size=x0.shape), device=dev)
    one = lambda v: torch.tensor([v], dtype=torch.float64, device=dev)
    x, p, H = x0.clone(), p0.clone(), torch.zeros(1, 4, dtype=torch.float64, device=dev)
    m.hmc_trajectory_device(1, 5, x, p, one(3.0), one(0.7), one(0.4), one(1e-4), H)
    H1 = H.clone(); p.neg_()
    m.hmc_trajectory_device(1, 5, x, p, one(3.0), one(0.7), one(0.4), one(1e-4), H)
    torch.cuda.synchronize()
    assert float((x - x0).abs().max()) < 1e-9 * float(x0.abs().max())
    assert float((p + p0).abs().max()) < 1e-7 * float(p0.abs().max())
    Hh, H1h = H.cpu().numpy()[0], H1.cpu().numpy()[0]
    assert abs((Hh[2] + Hh[3]) - (H1h[0] + H1h[1])) <= 1e-9 * abs(H1h[0])      # back at the initial Hamiltonian
    assert abs((H1h[2] + H1h[3]) - (H1h[0] + H1h[1])) < 1e-2 * abs(H1h[1])     # energy error small at eps = 1e-4
    m.close()


@pytest.mark.parametrize("S,path", [(8, "rows"), (32, "lanes")])
def test_config_e_shape_sparse_list(S, path):
    """10 000 beads, sparse list drawn with P(i, j) ~ 1/|i-j|, 8 / 32 structures"""
    from ensemble_hic_b200.engine import Model
    rng = np.random.default_rng(3)
    N, M = 10000, 400000
    sep = np.minimum((np.exp(rng.uniform(np.log(2), np.log(N - 1), size=3 * M))).astype(np.int64), N - 1)
    i = rng.integers(0, N, size=3 * M); j = i + sep
    ok = j < N
    key = np.unique(i[ok] * N + j[ok])[:M]
    data = np.stack([key // N, key % N, rng.poisson(3.0, size=len(key))], 1).astype(np.int64)
    data = np.ascontiguousarray(data[rng.permutation(len(data))])
    radii = np.full(N, 0.5); cds = np.full(len(data), 1.25)
    X = random_walk(rng, S, N, 0.9)
    fast = Model(S, radii, data, cds, alpha=10.0, k_nb=500.0, k_bb=500.0)
    exact = Model(S, radii, data, cds, alpha=10.0, k_nb=500.0, k_bb=500.0, exact=True)
    assert fast.info()["path"] == path                       # sparse: row path, or lane path when S fills the warps
    a = fast.evaluate(X, 3.0, 0.7, 0.4, energies=True, md=True)
    b = exact.evaluate(X, 3.0, 0.7, 0.4, energies=True, md=True)
    assert relerr(a["grad"], b["grad"]) < 1e-12 and relerr(a["md"], b["md"]) < 1e-12
    np.testing.assert_allclose(a["energies"], b["energies"], rtol=1e-11)
    g = a["grad"][0].reshape(S, N, 3)
    assert np.abs(g.sum(1)).max() < 1e-9 * np.abs(g).max() * N ** 0.5
    fast.close(); exact.close()
    if S == 8:
        # ... and against the oracle with the reference's production nonbonded path: cell-grid neighbour list +
        # PROLSQ (c/nblist.c:181-303, c/forcefield.c:8-109, c/prolsq.c:8-38: the reference's own C when oracle/_ref/libref_nbl.so is there), likelihood through the
        # reference's Cython
        op = _oracle(dict(N=N, radii=radii, data=data, cds=cds), S, "nbl")
        x = X.ravel()
        assert relerr(a["grad"][0], op.gradient(x, 3.0, 0.7, 0.4)) < 1e-10
        np.testing.assert_allclose(a["md"][0], op.mock_data(x, 3.0), rtol=1e-10, atol=4e-16 * 3.0 * S)
        U = op.energies(x, 3.0)
        np.testing.assert_allclose(a["energies"][0][:3], U[:3], rtol=1e-10)
        # no bead was dropped from a full grid cell (the reference only prints a warning; the restatement counts them)
        from oracle.oracle import OracleNBLForceField
        chk = OracleNBLForceField(op.co, radii, 500.0)
        for xk in X:
            chk.energy(xk)
            assert chk.nbl.n_dropped() == 0
