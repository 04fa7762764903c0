"""The program compute-sanitizer is pointed at (memcheck / racecheck / synccheck, one tool per gpurun call):
smoke() plus one small evaluation and one short HMC trajectory through every kernel family --

  tiles   3 x 140, dense list   forward_tile / backward_tile (mbarrier + cp.async.bulk rings), combine, prior, Verlet
  tiles32 the same in FP32 mode
  lanes   20 x 90, sparse list  forward_lane / backward_lane
  rows    3 x 60, exact mode    forward_kernel / gather_kernel (cp.async ring)
  small   2 x 23                trajectory_small_kernel (whole trajectory in shared memory)
  cluster 9 x 70                trajectory_cluster_kernel (DSMEM partial sums, cluster barriers)
  cell    1 x 4200              cell_nonbonded_kernel

Results are compared with nothing here (the parity tests do that); the point is the tool's report.
Usage: compute-sanitizer --tool racecheck python tools/sanitize_target.py [family ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import make_problem
from ensemble_hic_b200.engine import Model, launch_count

dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)


def run(name, env, S, N, density, R=2, L=4, exact=False, fp32=False, k_nb=50.0, evaluate=True, trajectory=True):
    for k in ("EHIC_PATH", "EHIC_TRAJ", "EHIC_TRAJ_CLUSTER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pb = make_problem(seed=S + N, S=S, N=N, density=density, radius=1.0, cd_factor=2.0, min_sep=2, step=1.8)
    m = Model(S, pb["radii"], pb["data"], pb["cds"], alpha=pb["alpha"], k_nb=k_nb, k_bb=500.0,
              sphere=(10.0, 5.0, np.zeros(3)), max_replicas=R, exact=exact, fp32=fp32)
    rng = np.random.default_rng(1)
    X = np.stack([pb["X"] + 0.05 * rng.normal(size=pb["X"].shape) for _ in range(R)]).reshape(R, -1)
    norm, lam, bet = np.full(R, 2.5), np.linspace(0.5, 1.0, R), np.linspace(0.2, 1.0, R)
    l0 = launch_count()
    if evaluate:
        out = m.evaluate(X, norm, lam, bet, energies=True, md=True)
        assert np.isfinite(out["grad"]).all()
    if trajectory:
        x_t, p_t, H = t(X), t(rng.normal(size=X.shape)), torch.zeros(R, 4, dtype=torch.float64, device=dev)
        for _ in range(2):
            m.hmc_trajectory_device(R, L, x_t, p_t, t(norm), t(lam), t(bet), t(np.full(R, 1e-3)), H)
        torch.cuda.synchronize()
        assert torch.isfinite(x_t).all()
    print("%-8s path=%s plan=%s launches=%d" % (name, m.info()["path"], m.trajectory_plan(R, L)["mode"] if trajectory else "-",
                                                launch_count() - l0), flush=True)
    m.close()


FAMILIES = {
    "tiles": lambda: run("tiles", {"EHIC_PATH": "tiles", "EHIC_TRAJ": "multi"}, 3, 140, 1.0),
    "tiles32": lambda: run("tiles32", {"EHIC_PATH": "tiles", "EHIC_TRAJ": "multi"}, 3, 140, 1.0, fp32=True),
    "lanes": lambda: run("lanes", {"EHIC_PATH": "lanes", "EHIC_TRAJ": "multi"}, 20, 90, 0.3),
    "rows": lambda: run("rows", {"EHIC_PATH": "rows", "EHIC_TRAJ": "multi"}, 3, 60, 0.5, exact=True),
    "small": lambda: run("small", {}, 2, 23, 1.0),
    "cluster": lambda: run("cluster", {"EHIC_TRAJ": "cluster", "EHIC_TRAJ_CLUSTER": "4"}, 9, 70, 0.35, evaluate=False),
    "cell": lambda: run("cell", {}, 1, 4200, 0.0005, R=1, trajectory=False),
}

if __name__ == "__main__":
    which = sys.argv[1:] or ["smoke"] + list(FAMILIES)
    for w in which:
        if w == "smoke":
            import __graft_entry__
            __graft_entry__.smoke()
        else:
            FAMILIES[w]()
    print("sanitize_target done")
