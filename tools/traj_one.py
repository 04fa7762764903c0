"""One (or SB_N) HMC trajectory launch(es) of a golden-fixture shape: the command profiled with ncu (development aid).
SB_NAME (nora2012), SB_R (38), SB_L (10), SB_N (2); EHIC_TRAJ / EHIC_TRAJ_CLUSTER select the path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ensemble_hic_b200.engine import Model
dev = torch.device("cuda:0")
name = os.environ.get("SB_NAME", "nora2012"); R = int(os.environ.get("SB_R", 38)); L = int(os.environ.get("SB_L", 10))
n = int(os.environ.get("SB_N", 2)); eps0 = float(os.environ.get("SB_EPS", 1e-3))
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
S, N = int(g["S"]), int(g["N"])
sphere = (float(g["sphere"][0]), float(g["sphere"][1]), np.zeros(3)) if len(g["sphere"]) else None
m = Model(S, g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]), sphere=sphere, max_replicas=R)
rng = np.random.default_rng(0)
X = np.stack([g["X"] + 0.01 * rng.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
x = t(X); p = t(rng.normal(size=X.shape)); norm = t(np.full(R, float(g["norm"])))
lam = t(np.linspace(0.1, 1, R)); eps = t(np.full(R, eps0)); H = torch.zeros(R, 4, dtype=torch.float64, device=dev)
print(m.info()["path"], m.trajectory_plan(R, L))
for q in range(n):
    p.normal_()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.hmc_trajectory_device(R, L, x, p, norm, lam, lam, eps, H)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("trajectory %d: %.3f ms, %.1f us per leapfrog step" % (q, dt * 1e3, dt * 1e6 / (L + 1)))
