"""Worker for tests/test_rex_gloo.py: runs the replica-exchange driver on a CPU test double of
the device backend (a toy quadratic model -- host logic only, not a kernel fallback)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ensemble_hic_b200.rex import ReplicaExchange, partition   # noqa: E402


class ToyBackend(object):
    """U_L = |x|^2/2 * norm, U_nb = |x - 1|^2/2, U_rest = 0; exact leapfrog in torch (CPU)."""
    torch = torch

    def __init__(self, n_dof):
        self.n_dof = n_dof

    def tensor(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))

    def _U(self, x, norm):
        return 0.5 * norm * (x * x).sum(1), 0.5 * ((x - 1.0) ** 2).sum(1)

    def energies(self, x, norm):
        UL, Unb = self._U(x, norm)
        z = torch.zeros_like(UL)
        return torch.stack([UL, Unb, z, z], 1)

    def _grad(self, x, norm, lam, bet):
        return lam[:, None] * norm[:, None] * x + bet[:, None] * (x - 1.0)

    def trajectory(self, x, p, norm, lam, bet, eps, n_steps):
        def H(x, p):
            UL, Unb = self._U(x, norm)
            return lam * UL + bet * Unb, 0.5 * (p * p).sum(1)
        E0, K0 = H(x, p)
        e = eps[:, None]
        p -= 0.5 * e * self._grad(x, norm, lam, bet)
        for _ in range(n_steps - 1):
            x += e * p
            p -= e * self._grad(x, norm, lam, bet)
        x += e * p
        p -= 0.5 * e * self._grad(x, norm, lam, bet)
        E1, K1 = H(x, p)
        return torch.stack([E0, K0, E1, K1], 1)

    def mock_sums(self, x):
        return (x * x).sum(1) + 1.0

    def select(self, mask, src, dst):
        m = mask.bool()
        dst[m] = src[m]


def main():
    out_dir, T, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    n_dof = 6
    b = partition(T, world)
    rng = np.random.RandomState(99)
    x_all = rng.normal(size=(T, n_dof)); norm_all = np.linspace(1.0, 2.0, T)
    sched = {"lammda": np.linspace(0.0, 1.0, T), "beta": np.linspace(0.1, 1.0, T)}
    rex = ReplicaExchange(ToyBackend(n_dof), sched, x_all[b[rank]:b[rank + 1]], norm_all[b[rank]:b[rank + 1]],
                          dict(timestep=0.2, trajectory_length=5, adaption_limit=20), norm_prior=(0.5, 0.5),
                          counts_sum=10.0, output_folder=out_dir, seed=5, rank=rank, world_size=world, dist=dist)
    for sub in ("samples", "statistics", "works"):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    if dist is not None:
        dist.barrier()
    rex.run(iters, swap_interval=3, status_interval=1000, dump_interval=10, dump_step=2, statistics_update_interval=5)
    res = dict(rank=rank, lo=rex.lo, hi=rex.hi, temp_of_slot=rex.temp_of_slot.tolist(), timestep_T=rex.timestep_T.tolist(),
               n_acc_T=rex.n_acc_T.tolist(), re_acc=rex.re_acc.tolist(), re_try=rex.re_try.tolist(),
               x=rex.x.numpy().tolist(), norm=rex.norm.numpy().tolist(),
               n_collectives=rex.n_collectives, n_swap_rounds=rex.n_swap_rounds, consistency=rex.consistency())
    with open(os.path.join(out_dir, "result_rank%d.json" % rank), "w") as f:
        json.dump(res, f)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
