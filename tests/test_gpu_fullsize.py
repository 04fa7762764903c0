"""Parity at BASELINE.json's full sizes, where the CPU oracle would take minutes: size-independent
properties of the posterior evaluated on the GPU (config D: 2000 beads x 100 structures, dense
list, M = 1 997 001; and a 10 000-bead sparse case of config E's shape, reduced in structures)."""
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
