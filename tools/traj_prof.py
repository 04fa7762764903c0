"""Per-kernel-class device time of one HMC trajectory (development aid): SB_NAME, SB_R, SB_L."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ensemble_hic_b200.engine import Model
dev = torch.device("cuda:0")
name = os.environ.get("SB_NAME", "nora2012"); R = int(os.environ.get("SB_R", 302)); L = int(os.environ.get("SB_L", 20))
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
S, N = int(g["S"]), int(g["N"])
m = Model(S, g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]), max_replicas=R)
rng = np.random.default_rng(0)
X = np.stack([g["X"] + 0.01 * rng.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
x = t(X); p = t(rng.normal(size=X.shape)); norm = t(np.full(R, float(g["norm"])))
lam = t(np.linspace(0.1, 1, R)); bet = t(np.linspace(0.1, 1, R)); eps = t(np.full(R, float(os.environ.get("SB_EPS", 1e-4)))); H = torch.zeros(R, 4, dtype=torch.float64, device=dev)
run = lambda: m.hmc_trajectory_device(R, L, x, p, norm, lam, bet, eps, H)
run(); torch.cuda.synchronize()
kt = m.kernel_times(run, repeats=3)
print(m.info()["path"], {k: round(v / (L + 1), 4) for k, v in kt.items()}, "ms per leapfrog step")
