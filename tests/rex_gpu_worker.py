"""Worker for tests/test_gpu_rex_ranks.py: the replica-exchange driver on the CUDA engine (EngineBackend), one process
per rank.  All ranks share cuda:0 -- NCCL refuses two ranks on one device, so the collective runs over gloo (host
staging in rex._allgather_rows); everything else is the production path."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_problem                                         # noqa: E402
from ensemble_hic_b200.engine import Model                               # noqa: E402
from ensemble_hic_b200.rex import EngineBackend, ReplicaExchange, partition   # noqa: E402


def main():
    out_dir, T, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    # hairpin-sized: 23 beads x 2 structures -> whole trajectories in trajectory_small_kernel, one CTA per replica, so a
    # replica's result does not depend on which other replicas share its launch
    pb = make_problem(seed=3, S=2, N=23, density=1.0, radius=1.0, cd_factor=2.0, min_sep=2, step=1.8)
    b = partition(T, world)
    R = int(b[rank + 1] - b[rank])
    m = Model(pb["S"], pb["radii"], pb["data"], pb["cds"], alpha=pb["alpha"], k_nb=50.0, k_bb=500.0,
              sphere=(10.0, 5.0, np.zeros(3)), max_replicas=max(R, 1), device=0)
    assert m.trajectory_plan(R, 10)["mode"] == "one_cta"
    rng = np.random.RandomState(99)
    x_all = np.stack([pb["X"].ravel() + 0.05 * rng.normal(size=pb["X"].size) for _ in range(T)])
    sched = {"lammda": np.linspace(0.97, 1.0, T), "beta": np.linspace(0.9, 1.0, T)}      # close temperatures: exchanges do get accepted
    counts = pb["data"][:, 2]
    rex = ReplicaExchange(EngineBackend(m), sched, x_all[b[rank]:b[rank + 1]], np.full(R, 3.0),
                          dict(timestep=1.5e-2, trajectory_length=10, adaption_limit=10 ** 6),
                          norm_prior=(float(np.mean(counts[counts > 0])) / pb["S"], 1.0 / pb["S"]), counts_sum=float(counts.sum()),
                          output_folder=None, seed=5, rank=rank, world_size=world, dist=dist, momenta="numpy")
    if dist is not None:
        dist.barrier()
    rex.run(iters, swap_interval=3, offset=0)
    torch.cuda.synchronize()
    res = dict(rank=rank, lo=rex.lo, hi=rex.hi, temp_of_slot=rex.temp_of_slot.tolist(), timestep_T=rex.timestep_T.tolist(),
               n_acc_T=rex.n_acc_T.tolist(), re_acc=rex.re_acc.tolist(), re_try=rex.re_try.tolist(),
               x=rex.x.cpu().numpy().tolist(), norm=rex.norm.cpu().numpy().tolist(),
               n_collectives=rex.n_collectives, n_swap_rounds=rex.n_swap_rounds, consistency=rex.consistency())
    with open(os.path.join(out_dir, "result_rank%d.json" % rank), "w") as f:
        json.dump(res, f)
    m.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
