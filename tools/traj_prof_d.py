"""Per-kernel-class device time of one leapfrog step of a config-D trajectory (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import random_walk
from ensemble_hic_b200.engine import Model
dev = torch.device("cuda:0")
N, S, R, L = 2000, 100, 8, int(os.environ.get("SB_L", 6))
rng = np.random.default_rng(0)
i, j = np.triu_indices(N, 2)
data = np.ascontiguousarray(np.stack([i, j, rng.poisson(2.0, size=len(i))], 1).astype(np.int64))
m = Model(S, np.full(N, 0.5), data, np.full(len(i), 1.25), alpha=10.0, k_nb=500.0, k_bb=500.0, max_replicas=R)
X = np.stack([random_walk(rng, S, N, 0.9) for _ in range(R)]).reshape(R, -1)
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
x = t(X); p = t(rng.normal(size=X.shape)); norm = t(np.full(R, 3.0))
lam = t(np.linspace(0.1, 1, R)); bet = t(np.linspace(0.1, 1, R)); eps = t(np.full(R, float(os.environ.get("SB_EPS", 1e-4)))); H = torch.zeros(R, 4, dtype=torch.float64, device=dev)
run = lambda: m.hmc_trajectory_device(R, L, x, p, norm, lam, bet, eps, H)
run(); torch.cuda.synchronize()
kt = m.kernel_times(run, repeats=3)
print(m.info()["path"], {k: round(v / (L + 1), 4) for k, v in kt.items()}, "ms per leapfrog step")
