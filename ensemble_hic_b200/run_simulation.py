"""Replica-exchange simulation from a config file -- the B200 counterpart of the reference's
scripts/run_simulation.py.

    reference:  mpirun -n <n_replicas + 1> python run_simulation.py config.cfg
    here:       python -m torch.distributed.run --nproc-per-node <n_gpus> --master-addr 127.0.0.1 \
                    -m ensemble_hic_b200.run_simulation config.cfg [--n-replicas T] [--seed S]
                (or plain ``python -m ensemble_hic_b200.run_simulation config.cfg`` on one GPU)

One process per GPU; the ladder of T temperatures is partitioned over the ranks and all local
replicas advance in one batched device call per trajectory (see :mod:`.rex`).  T is taken from
a pickled schedule's length, else ``--n-replicas``, else ``[replica] n_replicas``.
Outputs (config.cfg, schedule.pickle, samples/, statistics/, works/) land in
``[general] output_folder`` with the reference's file names.
"""
import argparse
import os
import sys

import numpy as np


def main(argv=None):
    import torch
    from . import setup_functions as sf
    from .rex import EngineBackend, ReplicaExchange, copy_run_metadata, load_continuation, partition

    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--n-replicas", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--n-samples", type=int, default=None, help="override [replica] n_samples")
    ap.add_argument("--momenta", default="numpy", choices=["numpy", "torch"])
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--continue", dest="n_cont", type=int, default=0, metavar="N",
                    help="continue a finished run (N-th continuation): start from the last dumped states and "
                         "timesteps, like scripts/continue_simulation of the reference")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    settings = sf.parse_config_file(args.config)
    out = settings["general"]["output_folder"]
    out = out if out.endswith("/") else out + "/"
    re_params = settings["replica"]
    if re_params["schedule"] not in ("linear", "exponential") and re_params["schedule"][-3:] != ".py":
        schedule = sf.load_schedule(re_params, 0)
        T = len(schedule["lammda"])
    else:
        T = args.n_replicas or int(re_params["n_replicas"])
        schedule = sf.load_schedule(re_params, T)
    bounds = partition(T, world)
    R = int(bounds[rank + 1] - bounds[rank])

    posterior = sf.make_posterior(settings, max_replicas=max(R, 1), device=local)
    posterior.sync_engine()
    np.random.seed(args.seed + 7919 * rank)
    x0 = np.stack([sf.setup_initial_state(settings["initial_state"], posterior).variables["structures"] for _ in range(R)])
    L = posterior.likelihoods["ensemble_contacts"]
    counts = np.asarray(L.error_model.data)
    sample_norm = "norm" in posterior.variables
    if sample_norm:
        norm0 = float(settings["initial_state"]["norm"])
        prior = posterior.priors["norm_prior"]
        norm_prior = (prior["shape"].value, prior["rate"].value)
    else:
        # run_simulation.py:93-94: a norm that is not sampled is fixed at max(count) / n_structures
        norm0 = float(np.max(counts)) / float(settings["general"]["n_structures"])
        posterior["norm"].set(norm0)
        norm_prior = None
    hmc = dict(timestep=float(settings["structures_hmc"]["timestep"]),
               trajectory_length=int(settings["structures_hmc"]["trajectory_length"]),
               adaption_limit=int(settings["structures_hmc"]["adaption_limit"]))
    offset, stats_folder = 0, None
    if args.n_cont > 0:
        # every rank reads the same files; rank 0 also writes init_states.npy etc.
        if rank == 0:
            cont = load_continuation(out, T, int(re_params["samples_dump_interval"]), args.n_cont)
        if dist is not None:
            dist.barrier()
        if rank != 0:
            cont = load_continuation(out, T, int(re_params["samples_dump_interval"]), args.n_cont)
        offset, stats_folder = cont["offset"], cont["cont_folder"] + "statistics"
        x0 = cont["structures"][bounds[rank]:bounds[rank + 1]]
        if cont["norms"] is not None and sample_norm:
            norm0 = cont["norms"][bounds[rank]:bounds[rank + 1]]
        if cont["timesteps"] is not None:
            hmc["timestep"] = cont["timesteps"]
    elif rank == 0:
        master = sf.setup_default_re_master(T, out)
        copy_run_metadata(args.config, schedule, out)
    if dist is not None:
        dist.barrier()
    rex = ReplicaExchange(EngineBackend(posterior.engine), schedule, x0, norm0, hmc, norm_prior=norm_prior,
                          counts_sum=float(counts.sum()), output_folder=out, seed=args.seed + 104729 * args.n_cont,
                          rank=rank, world_size=world, dist=dist, momenta=args.momenta, stats_folder=stats_folder)
    n = (args.n_samples if args.n_samples is not None else int(re_params["n_samples"])) + 1
    rex.run(n, swap_interval=int(re_params["swap_interval"]), status_interval=int(re_params["print_status_interval"]),
            dump_interval=int(re_params["samples_dump_interval"]), dump_step=int(re_params["samples_dump_step"]),
            statistics_update_interval=int(re_params["statistics_update_interval"]), offset=offset,
            verbose=not args.quiet)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return rex


if __name__ == "__main__":
    main()
    sys.exit(0)
