"""Seeded synthetic problems shared by the CPU and GPU tests."""
import numpy as np


def random_walk(rng, S, N, step):
    """S chains of N beads, consecutive beads `step` apart in a random direction."""
    d = rng.normal(size=(S, N, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    return np.cumsum(d * step, axis=1)


def dense_pairs(N, min_sep=1):
    i, j = np.triu_indices(N, min_sep + 1)
    return i, j


def make_problem(seed=0, S=3, N=40, density=1.0, radius=0.5, min_sep=1, shuffle=True,
                 uniform_radii=True, cd_factor=1.25, alpha=10.0, step=None):
    rng = np.random.default_rng(seed)
    radii = np.full(N, radius) if uniform_radii else rng.uniform(0.7 * radius, 1.3 * radius, N)
    X = random_walk(rng, S, N, (step if step is not None else 1.8 * radius))
    i, j = dense_pairs(N, min_sep)
    if density < 1.0:
        keep = rng.random(len(i)) < density
        i, j = i[keep], j[keep]
    cds = (radii[i] + radii[j]) * cd_factor
    # counts ~ Poisson(norm_true * sum_k s(d))
    d = np.linalg.norm(X[:, i] - X[:, j], axis=2)
    t = alpha * (cds[None] - d)
    s = 0.5 * (t / np.sqrt(1 + t * t) + 1)
    counts = rng.poisson(3.0 * s.sum(0))
    data = np.stack([i, j, counts], 1).astype(np.int64)
    if shuffle:
        p = rng.permutation(len(data))
        data, cds = data[p], cds[p]
    return dict(S=S, N=N, X=X, radii=radii, data=np.ascontiguousarray(data), cds=np.ascontiguousarray(cds),
                alpha=alpha, norm=3.3, rng=rng)


def relerr(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))
