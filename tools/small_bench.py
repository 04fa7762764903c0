"""Latency-bound configurations (hairpin_s, proteins, nora2012 3kb shapes from the golden fixtures):
time of one HMC trajectory (L leapfrog steps) for R batched replicas (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ensemble_hic_b200.engine import Model, launch_count

dev = torch.device("cuda:0")
L = int(os.environ.get("SB_L", 100))
for name, R in (("hairpin_s", 2), ("proteins", 7), ("proteins", 54), ("nora2012", 38), ("nora2012", 302)):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    S, N = int(g["S"]), int(g["N"])
    sphere = (g["sphere"][0], g["sphere"][1], np.zeros(3)) if len(g["sphere"]) else None
    m = Model(S, g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]),
              sphere=sphere, max_replicas=R)
    rng = np.random.default_rng(0)
    X = np.stack([g["X"] + 0.01 * rng.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
    x = t(X); p = t(rng.normal(size=X.shape)); norm = t(np.full(R, float(g["norm"])))
    lam = t(np.linspace(0.1, 1, R)); bet = t(np.linspace(0.1, 1, R)); eps = t(np.full(R, 1e-4)); H = torch.zeros(R, 4, dtype=torch.float64, device=dev)
    def run():
        m.hmc_trajectory_device(R, L, x, p, norm, lam, bet, eps, H)
    run(); torch.cuda.synchronize()
    n = 5
    l0 = launch_count(); t0 = time.perf_counter()
    for _ in range(n): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print("%-10s N=%4d S=%3d M=%6d R=%3d tiles=%s: trajectory L=%d %.3f ms  (%.1f us/leapfrog step, %d launches) -> %.0f replica-grad-evals/s"
          % (name, N, S, len(g["data"]), R, m.info()["path"] == "tiles", L, dt * 1e3, dt * 1e6 / (L + 1), (launch_count() - l0) // n, R * (L + 1) / dt), flush=True)
    m.close()
