"""Runs a few evaluations of config D (or QB_* overrides) for profiling under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import random_walk
from ensemble_hic_b200.engine import Model

N = int(os.environ.get("QB_N", 2000)); S = int(os.environ.get("QB_S", 100)); R = int(os.environ.get("QB_R", 2))
reps = int(os.environ.get("QB_REPS", 3))
rng = np.random.default_rng(0)
i, j = np.triu_indices(N, 2); M = len(i)
data = np.stack([i, j, rng.poisson(2.0, size=M)], 1).astype(np.int64)
m = Model(S, np.full(N, 0.5), data, np.full(M, 1.25), alpha=10.0, k_nb=500.0, k_bb=500.0, max_replicas=R)
dev = torch.device("cuda:0")
x = torch.tensor(np.stack([random_walk(rng, S, N, 0.9) for _ in range(R)]), device=dev)
norm = torch.full((R,), 3.0, dtype=torch.float64, device=dev)
g = torch.empty_like(x)
for _ in range(reps):
    m.evaluate_device(R, x, norm, None, None, grad=g)
torch.cuda.synchronize()
print("done", float(g.abs().max()))
