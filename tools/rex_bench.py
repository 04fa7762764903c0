"""Replica-exchange sweeps/s for one of the golden-fixture shapes (default nora2012: 308 beads x 30 structures,
M = 10145, 302-replica ladder) on the GPUs of one box, next to the reference's CPU gradient rate
(oracle/_ref Cython kernels + the neighbour-list C restatement, one process per host core).
Development aid; bench.py measures config D.

  python tools/rex_bench.py                                  # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/rex_bench.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def cpu_worker(args):
    name, seed = args
    from oracle.oracle import OraclePosterior, build_c_oracle
    build_c_oracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    S, N = int(g["S"]), int(g["N"])
    ul = g["radii"][:-1] + g["radii"][1:]
    op = OraclePosterior(S, g["data"], g["cds"], float(g["alpha"]), g["radii"], float(g["k_nb"]), float(g["k_bb"]),
                         [np.zeros(N - 1)], [ul], [0, N], nonbonded="nbl")
    x = g["X"] + 0.01 * np.random.default_rng(seed).normal(size=g["X"].shape)
    op.gradient(x, float(g["norm"]), 1.0, 1.0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 5.0:
        op.gradient(x, float(g["norm"]), 1.0, 1.0); n += 1
    return n / (time.perf_counter() - t0)


def main():
    import torch
    import torch.distributed as dist
    from ensemble_hic_b200.engine import Model
    from ensemble_hic_b200.rex import EngineBackend, ReplicaExchange
    name = os.environ.get("RB_NAME", "nora2012")
    n_total = int(os.environ.get("RB_REPLICAS", 302)); n_it = int(os.environ.get("RB_ITER", 10))
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    S, N = int(g["S"]), int(g["N"])
    from ensemble_hic_b200.rex import partition
    bounds = partition(n_total, world)
    R = int(bounds[rank + 1] - bounds[rank])
    m = Model(S, g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]),
              max_replicas=R, device=local)
    lad = np.linspace(0.0, 1.0, n_total) ** 2
    rng = np.random.default_rng(rank)
    X = np.stack([g["X"] + 0.01 * rng.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
    counts = g["data"][:, 2]
    rex = ReplicaExchange(EngineBackend(m), {"lammda": lad, "beta": lad}, X, np.full(R, float(g["norm"])),
                          dict(timestep=1e-3, trajectory_length=100, adaption_limit=0),
                          norm_prior=(float(np.mean(counts[counts > 0])) / S, 1.0 / S), counts_sum=float(counts.sum()),
                          output_folder=None, seed=0, rank=rank, world_size=world, dist=dist if world > 1 else None,
                          momenta="torch")
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    rex.run(1, swap_interval=5, offset=1)
    barrier(); t0 = time.perf_counter()
    rex.run(n_it, swap_interval=5, offset=1)
    barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        acc = float(np.mean(rex.n_acc_T / np.maximum(rex.n_hmc_T, 1)))
        print("%s: %d replicas on %d GPU(s), path %s: %.2f RE sweeps/s (%d iterations, L=100, swap every 5th; HMC acceptance %.2f)"
              % (name, n_total, world, m.info()["path"], n_it / dt, n_it, acc), flush=True)
        if os.environ.get("RB_CPU", "1") != "0":
            import multiprocessing as mp
            cores = len(os.sched_getaffinity(0))
            with mp.get_context("spawn").Pool(cores) as pool:
                rates = pool.map(cpu_worker, [(name, c) for c in range(cores)])
            tot = float(np.sum(rates))
            # one sweep of the ladder = n_total trajectories of 101 gradients (4 of 5 iterations)
            print("reference CPU kernels: %.0f posterior gradients/s on %d cores (%.1f per core) -> %.4f sweeps/s for %d replicas"
                  % (tot, cores, tot / cores, tot / (n_total * 101 * 0.8), n_total), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
