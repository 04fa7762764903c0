#!/usr/bin/env python
"""bench.py -- posterior gradient evaluations / s of the ensemble_hic hot path on B200.

Contract (see the task statement): ``python bench.py --gpus N --steps K --warmup W``;
for N > 1 it is launched under ``torch.distributed.run`` (one rank per GPU, NCCL).
Prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[3], SURVEY.md 8d "D"): synthetic chain, 2000 beads x 100
structures, dense contact list |i-j| > 1 (M = 1 997 001), 64 replicas over 8 GPUs
= 8 replicas per GPU (weak scaling: per-GPU work is fixed).  One *step* = one posterior
gradient evaluation (likelihood + nonbonded + backbone, the config has no sphere) of every
replica resident on the GPU, batched in one C-ABI call.  value = replicas * steps / time.

Beside the headline every other configuration BASELINE.json names is measured in the same run and reported under
``other_workloads`` at every N (each with the roofline fraction of its dominant kernel; at N = 1 also with the
reference's CPU kernels timed beside it):
  nora2012  308 beads x 30 structures, sparse list M = 10 145, the reference's full 302-temperature ladder
            (config_files/nora2012/30structures_3kb.cfg), L = 100, swap every 5th iteration: RE sweeps/s, STRONG scaling
            (the ladder is partitioned over the N GPUs)
  proteins  56 beads x 2 structures, M = 1 485, 54-temperature ladder (config_files/proteins/2structures.cfg): same
  E         10 000 beads x 200 structures, sparse list M = 1e7 (P(i,j) ~ 1/|i-j|), cell-list nonbonded, 16 replicas
            per GPU: posterior gradient evaluations/s, weak scaling

``--impl reference`` times the reference's own CPU kernels (oracle/_ref: the unmodified
Cython of /root/reference compiled for CPython 3, and its neighbour-list / PROLSQ C sources
compiled unmodified as plain C, oracle/build_ref.build_nbl) on a bounded sample of the same
workload with all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BEADS, N_STRUCT, REPLICAS_PER_GPU = 2000, 100, 8
RADIUS, CD_FACTOR, ALPHA, K_NB, K_BB = 0.5, 1.25, 10.0, 500.0, 500.0
METRIC = "posterior_grad_evals_per_s"
UNIT = "grad evals/s"

# ALGORITHMIC flop per unit (sqrt and div count 1 flop each)
FLOP_FWD_PER_PAIR = 22.0        # forward model, per (structure, data point)
FLOP_BWD_PER_PAIR = 23.0        # likelihood gradient, per (structure, data point)
FLOP_NB_PER_TEST = 9.0          # nonbonded, per tested pair i<j


def problem(seed, n_replicas, N=N_BEADS, S=N_STRUCT):
    """Seeded synthetic inputs of config D (counts are drawn on the fly from the forward model)."""
    rng = np.random.default_rng(seed)
    i, j = np.triu_indices(N, 2)
    radii = np.full(N, RADIUS)
    cds = (radii[i] + radii[j]) * CD_FACTOR                # = 1.25 for r = 0.5
    d = rng.normal(size=(n_replicas, S, N, 3))
    d /= np.linalg.norm(d, axis=3, keepdims=True)
    X = np.cumsum(d * (1.8 * RADIUS), axis=2)               # random-walk chains, step 1.8 r
    return dict(N=N, S=S, i=i, j=j, radii=radii, cds=cds, X=X, rng=rng)


def ladder(n_total):
    """lammda = beta, geometric-ish 1e-6 .. 1 (setup_functions.expspace shape)."""
    a, n = 0.1, np.arange(1, n_total + 1)
    g = (1.0 - 1e-6) / (np.exp(a * (n_total - 1.0)) - 1.0) * (np.exp(a * (n - 1.0)) - 1.0) + 1e-6
    return g


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bench_config(world):
    """the `config` object shared by both arms (own and --impl reference)"""
    M = N_BEADS * (N_BEADS - 1) // 2 - (N_BEADS - 1)
    R = REPLICAS_PER_GPU
    return {"workload": "D: synthetic chain %d beads x %d structures, dense contact list M=%d, %d replicas per GPU "
                        "(%d total), posterior gradient (likelihood+nonbonded+backbone)" % (N_BEADS, N_STRUCT, M, R, R * world),
            "replicas_per_gpu": R, "n_beads": N_BEADS, "n_structures": N_STRUCT, "n_data": M,
            "coordinates": "seeded random-walk chains, step 1.8 r (SURVEY 8d ii)",
            "l2": "working set per step (contact tiles 43 MB + weights 17 MB/replica + coordinates + reaction partials "
                  "77 MB/replica) exceeds the 126 MB L2; no explicit flush"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline

def _cpu_worker(args):
    seed, S_sample = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.oracle import OraclePosterior, load_ref
    pb = problem(seed, 1, S=S_sample)
    M = len(pb["i"])
    counts = pb["rng"].poisson(2.0, size=M)
    data = np.ascontiguousarray(np.stack([pb["i"], pb["j"], counts], 1).astype(np.int64))
    ul = pb["radii"][:-1] + pb["radii"][1:]
    op = OraclePosterior(S_sample, data, pb["cds"], ALPHA, pb["radii"], K_NB, K_BB,
                         [np.zeros(pb["N"] - 1)], [ul], [0, pb["N"]], kernels=load_ref(), nonbonded="nbl")
    x = pb["X"][0].ravel()
    t0 = time.perf_counter()
    # the reference's call sequence per posterior gradient: calculate_gradient;
    # per structure nonbonded energy (unused) + gradient, each rebuilding the list; backbone
    for xk in x.reshape(S_sample, -1, 3):
        op.nb_energy(xk)
    g = op.gradient(x, 3.0, 1.0, 1.0)
    dt = time.perf_counter() - t0
    return dt, float(np.abs(g).max())


def cpu_reference_rate(budget_s=20.0, cores=None):
    """grad evals/s of the reference CPU kernels on the host cores, on a bounded sample."""
    import multiprocessing as mp
    from oracle import build_ref
    from oracle.oracle import build_c_oracle, load_ref
    build_c_oracle()
    build_ref.build(verbose=False)
    build_ref.build_nbl(verbose=False)
    kind = "reference" if load_ref() is not None else "port"
    cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    try:
        import psutil
        free_gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        free_gb = 32.0
    # the reference allocates 2 x S x M doubles of scratch per call (likelihoods_c.pyx:36-43)
    S_sample = N_STRUCT
    per_s = 0.11                                    # ~ seconds of one core per structure (probe)
    while S_sample > 5 and (S_sample * per_s > budget_s or cores * S_sample * 0.04 > 0.6 * free_gb):
        S_sample //= 2
    ctx = mp.get_context("spawn")      # the parent may hold a CUDA context: never fork it
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(1000 + c, S_sample) for c in range(cores)])
    wall = max(r[0] for r in res)
    # cost is linear in the number of structures: scale the sample to a full S = 100 gradient
    rate = cores / (wall * N_STRUCT / S_sample)
    sample = ("one posterior gradient per core at N=%d, M=%d, S_sample=%d of %d structures "
              "(cost linear in S, scaled by %d/%d); %d worker processes; %s kernels%s"
              % (N_BEADS, len(np.triu_indices(N_BEADS, 2)[0]), S_sample, N_STRUCT, N_STRUCT, S_sample, cores,
                 "reference Cython (oracle/_ref)" if kind == "reference" else "C oracle port",
                 ", neighbour-list forcefield: the reference's own c/nblist.c + c/prolsq.c (oracle/_ref/libref_nbl.so)"
                 if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_nbl.so")) and kind == "reference"
                 else ", neighbour-list forcefield via the bit-identical C restatement"))
    return rate, cores, kind, sample, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rates, last = [], None
    for s in range(args.warmup + args.steps):
        r = cpu_reference_rate(budget_s=max(3.0, 100.0 / max(1, args.warmup + args.steps)))
        if s >= args.warmup:
            rates.append(r[0])
        last = r
    rate = float(np.mean(rates))
    _, cores, kind, sample, wall = last
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.gpus),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# the other named configurations (BASELINE.json configs[1], [2], [4])

GOLD = os.path.join(ROOT, "tests", "golden")
E_N, E_S, E_M, E_REPLICAS_PER_GPU = 10000, 200, 10 ** 7, 16


def _golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sphere = (float(g["sphere"][0]), float(g["sphere"][1]), np.zeros(3)) if len(g["sphere"]) else None
    return g, sphere


def _ladder_cpu_worker(args):
    """reference CPU kernels on one core: posterior gradients/s for one replica of a golden-fixture shape"""
    name, seed, budget = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.oracle import OraclePosterior, load_ref
    g, sphere = _golden(name)
    S, N = int(g["S"]), int(g["N"])
    ul = g["radii"][:-1] + g["radii"][1:]
    op = OraclePosterior(S, g["data"], g["cds"], float(g["alpha"]), g["radii"], float(g["k_nb"]), float(g["k_bb"]),
                         [np.zeros(N - 1)], [ul], [0, N], sphere=sphere, kernels=load_ref(), nonbonded="nbl")
    x = (g["X"] + 0.01 * np.random.default_rng(seed).normal(size=g["X"].shape)).ravel()
    nrm = float(g["norm"])

    def one():      # the reference's call sequence per posterior gradient, incl. the unused energy pass
        for xk in x.reshape(S, -1, 3):
            op.nb_energy(xk)
        op.gradient(x, nrm, 1.0, 1.0)
    one()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        one(); n += 1
    return n / (time.perf_counter() - t0)


def _e_cpu_worker(args):
    seed, S_sample, M_sample = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.oracle import OraclePosterior, load_ref
    rng = np.random.default_rng(seed)
    data = e_contact_list(rng, E_N, M_sample)
    radii = np.full(E_N, RADIUS)
    d = rng.normal(size=(S_sample, E_N, 3)); d /= np.linalg.norm(d, axis=2, keepdims=True)
    X = np.cumsum(d * (1.8 * RADIUS), axis=1)
    ul = radii[:-1] + radii[1:]
    op = OraclePosterior(S_sample, data, np.full(len(data), 2 * RADIUS * CD_FACTOR), ALPHA, radii, K_NB, K_BB,
                         [np.zeros(E_N - 1)], [ul], [0, E_N], kernels=load_ref(), nonbonded="nbl")
    x = X.ravel()
    t0 = time.perf_counter()
    for xk in X:                               # nonbonded energy (unused by the reference) + gradient, backbone
        op.nb_energy(xk)
    op.gradient(x, 3.0, 1.0, 1.0, terms=("nb", "bb"))
    t1 = time.perf_counter()
    op.gradient(x, 3.0, 1.0, 1.0, terms=("lik",))
    return t1 - t0, time.perf_counter() - t1


def e_contact_list(rng, N, M):
    """sparse contact list of config E: M pairs drawn without replacement with P(i, j) ~ 1/|i-j|, |i-j| > 1"""
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < M:
        n = int(1.6 * (M - len(keys))) + 1000
        sep = np.minimum(np.exp(rng.uniform(np.log(2), np.log(N - 1), size=n)).astype(np.int64), N - 1)
        i = rng.integers(0, N, size=n); j = i + sep
        ok = j < N
        keys = np.unique(np.concatenate([keys, i[ok] * N + j[ok]]))
    keys = rng.permutation(keys)[:M]
    return np.ascontiguousarray(np.stack([keys // N, keys % N, rng.poisson(3.0, size=M)], 1).astype(np.int64))


def _pool_map(fn, jobs, cores):
    import multiprocessing as mp
    from oracle import build_ref
    from oracle.oracle import build_c_oracle
    build_c_oracle()
    build_ref.build(verbose=False)
    with mp.get_context("spawn").Pool(cores) as pool:      # never fork a process that holds a CUDA context
        return pool.map(fn, jobs)


def _host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()


def bench_ladder(name, ctx, n_iter, L=100, swap_interval=5):
    """RE sweeps/s of a shipped configuration's full ladder, partitioned over the GPUs (strong scaling).  One sweep =
    one iteration of the whole ladder: every replica draws one Gibbs sample (HMC trajectory of L leapfrog steps =
    L + 1 gradients + 2 energies, then the conjugate norm draw with one forward-model evaluation); every
    swap_interval-th iteration is an exchange round instead (scripts/run_simulation.py:64-69)."""
    import torch
    from ensemble_hic_b200.engine import Model, launch_count
    from ensemble_hic_b200.rex import EngineBackend, ReplicaExchange, partition
    world, rank, local, dev, dist = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["dist"]
    g, sphere = _golden(name)
    sched = np.load(os.path.join(GOLD, "schedules.npz"))
    lam_all, bet_all = sched[name + "_lammda"], sched[name + "_beta"]
    T = len(lam_all)
    bounds = partition(T, world)
    R = int(bounds[rank + 1] - bounds[rank])
    S, N, M = int(g["S"]), int(g["N"]), len(g["data"])
    kw = dict(alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]), sphere=sphere, device=local)
    # synthetic counts of the same sparse shape: the shipped (i, j) list, counts ~ Poisson(norm * fwm(x*)) at the
    # seeded structures x* of the fixture, norm chosen so that the total count equals the real data's
    m0 = Model(S, g["radii"], g["data"], g["cds"], max_replicas=1, **kw)
    md1 = m0.mock_data(g["X"], 1.0)[0]
    m0.close()
    norm_true = float(g["data"][:, 2].sum()) / float(md1.sum())
    counts = np.random.default_rng(2012).poisson(norm_true * md1)
    data = np.ascontiguousarray(np.stack([g["data"][:, 0], g["data"][:, 1], counts], 1).astype(np.int64))
    m = Model(S, g["radii"], data, g["cds"], max_replicas=max(R, 1), **kw)
    rs = np.random.default_rng(1000 + rank)
    X0 = np.stack([g["X"] + 0.01 * rs.normal(size=g["X"].shape) for _ in range(R)]).reshape(R, -1)
    eps0 = 1e-3        # (the shipped 1e-4 is the start of an adaptation that runs for the whole simulation)
    rex = ReplicaExchange(EngineBackend(m), {"lammda": lam_all, "beta": bet_all}, X0, np.full(R, norm_true),
                          dict(timestep=eps0, trajectory_length=L, adaption_limit=10 ** 8),
                          norm_prior=(float(np.mean(counts[counts > 0])) / S, 1.0 / S), counts_sum=float(counts.sum()),
                          output_folder=None, seed=0, rank=rank, world_size=world, dist=dist, momenta="torch")
    rex.run(swap_interval + 1, swap_interval=swap_interval, offset=0)         # warm-up: 5 samples + 1 exchange round
    ctx["barrier"]()
    l0, c0 = launch_count(), rex.n_collectives
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rex.run(n_iter, swap_interval=swap_interval, offset=swap_interval + 1)
    e1.record()
    ctx["barrier"]()
    ms = ctx["max_over_ranks"](e0.elapsed_time(e1))
    launches = launch_count() - l0
    n_swaps = len([it for it in range(swap_interval + 1, swap_interval + 1 + n_iter) if it % swap_interval == 0])
    n_samples = n_iter - n_swaps
    # replicated bookkeeping must agree on every rank after the run (exchanges really happened across GPUs)
    mine = rex.consistency()
    consistent = True
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, mine)
        consistent = all(c == allc[0] for c in allc)
    plan = m.trajectory_plan(max(R, 1), L)
    out = None
    if rank == 0:
        # dominant kernel: one trajectory of the local replicas, timed alone with CUDA events
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
        xw, pw = t(X0), t(rs.normal(size=X0.shape))
        lamw = t(lam_all[bounds[0]:bounds[1]]); epsw = t(np.full(R, eps0)); nrmw = t(np.full(R, norm_true))
        Hw = torch.zeros(R, 4, dtype=torch.float64, device=dev)
        run = lambda: m.hmc_trajectory_device(R, L, xw, pw, nrmw, lamw, lamw, epsw, Hw)
        run(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(3):
            run()
        a1.record(); torch.cuda.synchronize()
        traj_ms = a0.elapsed_time(a1) / 3
        contact_flop = (FLOP_FWD_PER_PAIR + FLOP_BWD_PER_PAIR) * S * M * R * (L + 1)
        roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": ctx["fp64_tf"], "traffic": None,
                "trajectory_ms": traj_ms, "ms_per_leapfrog_step": traj_ms / (L + 1), "trajectory_mode": plan["mode"],
                "ctas_per_replica": plan["ctas_per_replica"],
                "flop_model": "contact term only: 45 algorithmic flop per (structure, data point) per gradient "
                              "(SURVEY 8d); the nonbonded / backbone work is not counted"}
        if plan["mode"] == "graph":
            km = m.kernel_times(run, repeats=2)
            dom = max(("forward", "gather"), key=lambda k: km.get(k, 0.0))
            flop_dom = (FLOP_BWD_PER_PAIR if dom == "gather" else FLOP_FWD_PER_PAIR) * S * M * R * (L + 1)
            roof.update(kernel={"gather": "backward_lane_kernel" if m.info()["path"] == "lanes" else "gather_kernel",
                                "forward": "forward_lane_kernel" if m.info()["path"] == "lanes" else "forward_kernel"}[dom],
                        achieved=flop_dom / (km[dom] * 1e-3) / 1e12, kernel_ms_per_trajectory=km,
                        note="kernel durations from a profiled run of the same step sequence (one launch per kernel, "
                             "no graph); trajectory_ms is the production path (one CUDA graph)")
        else:
            roof.update(kernel="trajectory_small_kernel" if plan["mode"] == "one_cta" else "trajectory_cluster_kernel",
                        achieved=contact_flop / (traj_ms * 1e-3) / 1e12,
                        note="the whole trajectory is ONE launch of this kernel")
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["whole_trajectory"] = {"achieved": contact_flop / (traj_ms * 1e-3) / 1e12,
                                    "frac": contact_flop / (traj_ms * 1e-3) / 1e12 / roof["peak"]}
        # optional FP32 mode (sparse lists: FP32 lane kernels) on the same trajectory, beside the FP64 figure
        fp32_mode = None
        try:
            m32 = Model(S, g["radii"], data, g["cds"], max_replicas=max(R, 1), fp32=True, **kw)
            x32, p32 = t(X0), t(rs.normal(size=X0.shape))
            run32 = lambda: m32.hmc_trajectory_device(R, L, x32, p32, nrmw, lamw, lamw, epsw, Hw)
            run32(); torch.cuda.synchronize()
            a0.record()
            for _ in range(3):
                run32()
            a1.record(); torch.cuda.synchronize()
            ms32 = a0.elapsed_time(a1) / 3
            fp32_mode = {"trajectory_ms": ms32, "ms_per_leapfrog_step": ms32 / (L + 1), "path": m32.info()["path"],
                         "trajectory_mode": m32.trajectory_plan(max(R, 1), L)["mode"], "speedup_vs_f64": traj_ms / ms32,
                         "dtype": "f32 pair arithmetic of the contact term, f64 I/O and priors"}
            m32.close(); del x32, p32
        except Exception as exc:                                    # reported, never fatal for the FP64 record
            fp32_mode = {"error": str(exc)}
        sec = ms * 1e-3
        out = {"workload": "%s: %d beads x %d structures, M=%d, full ladder of %d temperatures partitioned over %d GPU(s), "
                           "L=%d, swap every %dth iteration, norm sampled" % (name, N, S, M, T, world, L, swap_interval),
               "metric": "re_sweeps_per_s", "value": n_iter / sec, "unit": "RE sweeps/s (whole ladder)",
               "scaling": "strong", "iterations": n_iter, "exchange_rounds": n_swaps, "ms_per_sweep": ms / n_iter,
               "replica_grad_evals_per_s": T * n_samples * (L + 1) / sec,
               "replica_trajectories_per_s": T * n_samples / sec,
               "path": m.info()["path"], "replicas_on_rank0": R, "gpu_launches": launches,
               "collectives": rex.n_collectives - c0, "rex_consistent": bool(consistent),
               "hmc_acceptance": float(np.mean(rex.n_acc_T / np.maximum(rex.n_hmc_T, 1))),
               "re_acceptance": float(np.sum(rex.re_acc) / max(np.sum(rex.re_try), 1)),
               "timestep_range": [float(rex.timestep_T.min()), float(rex.timestep_T.max())],
               "roofline": roof, "fp32_mode": fp32_mode}
        if world == 1 and ctx["cpu"]:
            cores = _host_cores()
            rates = _pool_map(_ladder_cpu_worker, [(name, c, 4.0) for c in range(cores)], cores)
            tot = float(np.sum(rates))
            per_sweep = T * ((L + 1) * (1.0 - 1.0 / swap_interval))       # gradients per average sweep
            out["cpu_baseline"] = {"value": tot / per_sweep, "unit": "RE sweeps/s (whole ladder)", "cores": cores,
                                   "kind": ctx["cpu_kind"], "replica_grad_evals_per_s": tot,
                                   "sample": "posterior gradients of one replica per core for 4 s (reference call sequence, "
                                             "SURVEY 3.4; neighbour-list forcefield: the reference's own C where oracle/_ref/libref_nbl.so exists, else its bit-identical restatement), converted with "
                                             "%.0f gradients per sweep of the ladder; energies, norm draw and exchange not "
                                             "charged to the CPU" % per_sweep}
        del xw, pw
    m.close()
    return out


def bench_e(ctx, steps):
    """config E: 10 000 beads x 200 structures, sparse list M = 1e7, cell-list nonbonded, 16 replicas per GPU"""
    import torch
    from ensemble_hic_b200.engine import Model, launch_count
    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    N, S, M, R = E_N, E_S, E_M, E_REPLICAS_PER_GPU
    t0 = time.perf_counter()
    data = e_contact_list(np.random.default_rng(5), N, M)                       # the same list on every rank
    m = Model(S, np.full(N, RADIUS), data, np.full(M, 2 * RADIUS * CD_FACTOR), alpha=ALPHA, k_nb=K_NB, k_bb=K_BB,
              max_replicas=R, device=local)
    setup_s = time.perf_counter() - t0
    rng = np.random.default_rng(100 + rank)
    x = torch.empty(R, S, N, 3, dtype=torch.float64, device=dev)
    for r in range(R):
        d = rng.normal(size=(S, N, 3)); d /= np.linalg.norm(d, axis=2, keepdims=True)
        x[r] = torch.from_numpy(np.cumsum(d * (1.8 * RADIUS), axis=1)).to(dev)
    norm = torch.full((R,), 3.0, dtype=torch.float64, device=dev)
    lad = torch.tensor(ladder(R * world)[rank * R:(rank + 1) * R], dtype=torch.float64, device=dev)
    g = torch.empty_like(x)
    step = lambda: m.evaluate_device(R, x, norm, lad, lad, grad=g)
    for _ in range(2):
        step()
    ctx["barrier"]()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx["barrier"]()
    ms = ctx["max_over_ranks"](e0.elapsed_time(e1)) / steps
    launches = launch_count() - l0
    out = None
    if rank == 0:
        km = m.kernel_times(step, repeats=2)
        dom = max(("forward", "gather", "prior"), key=lambda k: km.get(k, 0.0))
        flop = {"gather": FLOP_BWD_PER_PAIR * S * M * R, "forward": FLOP_FWD_PER_PAIR * S * M * R}.get(dom)
        path = m.info()["path"]
        names = {"gather": "backward_lane_kernel" if path == "lanes" else "gather_kernel",
                 "forward": "forward_lane_kernel" if path == "lanes" else "forward_rec_kernel", "prior": "prior / cell kernels"}
        ach = flop / (km[dom] * 1e-3) / 1e12 if flop else None
        flop_step = (FLOP_FWD_PER_PAIR + FLOP_BWD_PER_PAIR) * S * M * R
        out = {"workload": "E: synthetic chain %d beads x %d structures, sparse contact list M=%d (P(i,j) ~ 1/|i-j|), cell-list "
                           "nonbonded, %d replicas per GPU (%d total), posterior gradient" % (N, S, M, R, R * world),
               "metric": METRIC, "value": R * world / (ms * 1e-3), "unit": UNIT, "scaling": "weak", "ms_per_step": ms,
               "steps": steps, "path": path, "gpu_launches": launches, "setup_s": setup_s,
               "gpu_memory_gb": (torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 2 ** 30,
               "roofline": {"bound": "fp64", "kernel": names[dom], "achieved": ach, "peak": ctx["fp64_tf"], "unit": "TFLOP/s",
                            "frac": ach / ctx["fp64_tf"] if ach else None, "traffic": None, "kernel_ms": km,
                            "step": {"achieved": flop_step / (ms * 1e-3) / 1e12,
                                     "frac": flop_step / (ms * 1e-3) / 1e12 / ctx["fp64_tf"]},
                            "flop_model": "contact term: 22 (forward) / 23 (backward) algorithmic flop per (structure, data "
                                          "point) (SURVEY 8d); nonbonded cell-list work not counted"}}
        if world == 1 and ctx["cpu"]:
            cores = _host_cores()
            try:
                import psutil
                free_gb = psutil.virtual_memory().available / 2 ** 30
            except Exception:
                free_gb = 32.0
            S_s, M_s = 4, M // 4            # cost is linear in S and in M: a 1/200 sample of the pair evaluations per core
            per_worker_gb = (2 * S_s * M_s * 8 + 5 * M_s * 8 + 3 * E_N * E_N * 4) / 2 ** 30     # scratch + N x N lists of the nbl
            use = int(max(1, min(cores, 0.5 * free_gb / per_worker_gb)))
            parts = _pool_map(_e_cpu_worker, [(c, S_s, M_s) for c in range(use)], use)
            wall = max(p[0] for p in parts) * (S / S_s) + max(p[1] for p in parts) * (S / S_s) * (M / M_s)
            out["cpu_baseline"] = {"value": use / wall, "unit": UNIT, "cores": use, "kind": ctx["cpu_kind"],
                                   "sample": "one posterior gradient per core at N=%d with S_sample=%d of %d structures and "
                                             "M_sample=%d of %d data points (contact term scaled by S/S_sample x M/M_sample, priors by "
                                             "S/S_sample); reference call sequence; neighbour-list forcefield: the reference's own C where "
                                             "oracle/_ref/libref_nbl.so exists, else its bit-identical restatement" % (N, S_s, S, M_s, M)}
    m.close()
    del x, g
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# own arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    from ensemble_hic_b200 import engine
    from ensemble_hic_b200.engine import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R = REPLICAS_PER_GPU
    n_total = R * world

    pb = problem(rank, R)
    N, S = pb["N"], pb["S"]
    M = len(pb["i"])
    # counts ~ Poisson(fwm(x*)) drawn from the (parity-tested) forward kernel at replica 0's x*
    data0 = np.ascontiguousarray(np.stack([pb["i"], pb["j"], np.zeros(M, dtype=np.int64)], 1).astype(np.int64))
    m0 = Model(S, pb["radii"], data0, pb["cds"], alpha=ALPHA, k_nb=K_NB, k_bb=K_BB, max_replicas=1, device=local)
    md = m0.mock_data(pb["X"][0], 3.0)[0]
    m0.close()
    counts = np.random.default_rng(12345).poisson(md)        # same contact counts on every rank
    data = np.ascontiguousarray(np.stack([pb["i"], pb["j"], counts], 1).astype(np.int64))
    model = Model(S, pb["radii"], data, pb["cds"], alpha=ALPHA, k_nb=K_NB, k_bb=K_BB, max_replicas=R, device=local)
    lad = ladder(n_total)[rank * R:(rank + 1) * R]

    x = torch.tensor(pb["X"], dtype=torch.float64, device=dev)
    norm = torch.full((R,), max(float(counts.max()) / S, 1.0), dtype=torch.float64, device=dev)
    lam = torch.tensor(lad, dtype=torch.float64, device=dev)
    bet = torch.tensor(lad, dtype=torch.float64, device=dev)
    g = torch.empty_like(x)

    def step():
        model.evaluate_device(R, x, norm, lam, bet, grad=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from fma_peak import fma_peak
    fp64_tf, fp32_tf = fma_peak(True, 10, local)[0], fma_peak(False, 10, local)[0]
    cpu_kind = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import build_ref
        from oracle.oracle import build_c_oracle, load_ref
        build_c_oracle(); build_ref.build(verbose=False)
        cpu_kind = "reference" if load_ref() is not None else "port"
    ctx = dict(world=world, rank=rank, local=local, dev=dev, dist=dist if world > 1 else None, barrier=barrier,
               max_over_ranks=max_over_ranks, fp64_tf=fp64_tf, cpu=cpu_kind is not None, cpu_kind=cpu_kind)
    workloads = [] if args.no_extras else [w for w in args.workloads.split(",") if w]

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = engine.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = n_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public host API: pinned host buffers, H2D + D2H inside ------
    xh = torch.empty(x.shape, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    xh_np = xh.numpy()
    normh, lamh, beth = norm.cpu().numpy(), lam.cpu().numpy(), bet.cpu().numpy()
    gh = torch.empty((R, S * N * 3), dtype=torch.float64).pin_memory()
    L = model._L
    from ensemble_hic_b200 import _lib

    def step_e2e():
        _lib.check(L.ehic_evaluate_host(model._h, R, 15, xh_np.ctypes.data, normh.ctypes.data, lamh.ctypes.data,
                                        beth.ctypes.data, gh.data_ptr(), None, None))
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = n_total * n_e2e / float(dt.item())
    bytes_x = R * S * N * 3 * 8

    # ---- replica-exchange sweeps / s (second half of BASELINE.json's metric) ----------------------
    # one sweep = one iteration of the whole ladder: every replica draws one Gibbs sample (HMC
    # trajectory of L = 100 leapfrog steps = 101 gradients + 2 energies, then the norm draw with one
    # forward-model evaluation); every 5th iteration is an exchange round instead (swap_interval = 5).
    sweeps = None
    if args.sweeps > 0:
        from ensemble_hic_b200.rex import EngineBackend, ReplicaExchange
        lad_all = ladder(n_total)
        rex = ReplicaExchange(EngineBackend(model), {"lammda": lad_all, "beta": lad_all}, pb["X"].reshape(R, -1),
                              norm.cpu().numpy(), dict(timestep=1e-3, trajectory_length=100, adaption_limit=10 ** 8),
                              norm_prior=(float(np.mean(counts[counts > 0])) / S, 1.0 / S), counts_sum=float(counts.sum()),
                              output_folder=None, seed=0, rank=rank, world_size=world, dist=dist if world > 1 else None,
                              momenta="torch")
        rex.run(1, swap_interval=5, offset=1)                       # warm-up: one sample
        barrier()
        t0 = time.perf_counter()
        rex.run(args.sweeps, swap_interval=5, offset=1)            # it = 1 .. sweeps: exchange rounds at 5, 10, ...
        barrier()
        dts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dts, op=dist.ReduceOp.MAX)
        mine, consistent = rex.consistency(), True
        if world > 1:
            allc = [None] * world
            dist.all_gather_object(allc, mine)
            consistent = all(c == allc[0] for c in allc)
        sweeps = {"value": args.sweeps / float(dts.item()), "unit": "RE sweeps/s (whole ladder of %d replicas)" % n_total,
                  "iterations": args.sweeps, "trajectory_length": 100, "swap_interval": 5, "scaling": "weak",
                  "hmc_acceptance": float(np.mean(rex.n_acc_T / np.maximum(rex.n_hmc_T, 1))),
                  "re_acceptance": float(np.sum(rex.re_acc) / max(np.sum(rex.re_try), 1)),
                  "rex_consistent": bool(consistent),
                  "exchange_rounds": args.sweeps // 5,
                  "replica_trajectories_per_s": n_total * (args.sweeps - args.sweeps // 5) / float(dts.item())}
        del rex

    # ---- optional FP32 mode (north star: 1e-5 relative), same step, reported beside the FP64 headline ------
    fp32_mode = None
    if rank == 0 and world == 1 and model.info()["path"] == "tiles":
        m32 = Model(S, pb["radii"], data, pb["cds"], alpha=ALPHA, k_nb=K_NB, k_bb=K_BB, max_replicas=R, device=local, fp32=True)
        g32 = torch.empty_like(x)
        step()
        for _ in range(3):
            m32.evaluate_device(R, x, norm, lam, bet, grad=g32)
        torch.cuda.synchronize()
        n32 = max(3, min(args.steps, 20))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n32):
            m32.evaluate_device(R, x, norm, lam, bet, grad=g32)
        a1.record()
        torch.cuda.synchronize()
        ms32 = a0.elapsed_time(a1) / n32
        fp32_mode = {"value": R / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "dtype": "f32 pair arithmetic, f64 I/O",
                     "grad_max_rel_dev_from_f64": float((g32 - g).abs().max() / g.abs().max())}
        m32.close()
        del g32

    # ---- per-kernel durations (CUDA events on the launching stream, same call sequence) ------
    kern = None
    if rank == 0:
        kern = model.kernel_times(lambda: step(), repeats=max(3, min(args.steps, 10)))

    tiles = model.info()["path"] == "tiles"
    # ---- every other configuration BASELINE.json names, at every N (all ranks take part) -----------------------
    model.close()
    del x, g, xh, gh
    torch.cuda.empty_cache()
    other = {}
    if "nora2012" in workloads:
        other["nora2012"] = bench_ladder("nora2012", ctx, args.ladder_iters)
    if "proteins" in workloads:
        other["proteins"] = bench_ladder("proteins", ctx, args.ladder_iters)
    if "E" in workloads:
        other["E"] = bench_e(ctx, max(2, min(args.steps, 5)))

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        # dominant kernel: the likelihood-gradient kernel (timing class "gather": backward_tile_kernel on the
        # dense tile path, gather_kernel on the row path); algorithmic flop = 23 per (structure, data point)
        gather_ms = kern.get("gather", None) if kern else None
        flop_gather = FLOP_BWD_PER_PAIR * S * M * R
        ach = flop_gather / (gather_ms * 1e-3) / 1e12 if gather_ms else None
        flop_step = ((FLOP_FWD_PER_PAIR + FLOP_BWD_PER_PAIR) * S * M + FLOP_NB_PER_TEST * S * N * (N - 1) / 2) * R
        bytes_step = (24.0 * S * N * 2 + 24.0 * M + 16.0 * M) * R
        prof = {}
        pj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(pj):
            with open(pj) as f:
                prof = json.load(f)
        roofline = {"bound": "fp64", "kernel": "backward_tile_kernel" if tiles else "gather_kernel", "achieved": ach, "peak": fp64_tf, "unit": "TFLOP/s",
                    "frac": (ach / fp64_tf) if ach else None, "traffic": prof.get("gather_dram_bytes_per_step", prof.get("gather_dram_bytes_per_launch")),
                    "peak_source": "builder-measured peak: DFMA issue-rate micro-benchmark run in this process "
                                   "(tools/peak/fma_peak.cu); FP64/FP32 non-tensor peaks are not in MEASURED_PEAKS.json (SURVEY 8d)",
                    "traffic_source": "ncu --set full capture of this kernel at commit %s (profiles/traffic.json), not measured in "
                                      "this run" % prof.get("commit", "?"),
                    "algorithmic_flop_per_launch": flop_gather, "launches_per_step": prof.get("launches_per_step", 1),
                    "note": "per step of the whole resident batch: when the batch runs as several part launches, achieved = "
                            "algorithmic flop of the batch / summed duration of the kernel's launches, traffic = sum over them",
                    "kernel_ms": kern,
                    "issue_ceiling": ({"frac": FLOP_BWD_PER_PAIR / (2.0 * 29.84),
                                       "note": "backward_tile_kernel issues 29 FP64 instructions + 0.84 reduction adds per (structure, "
                                               "data point) for 23 algorithmic flop and the peak counts an FMA as 2 flop: frac cannot "
                                               "exceed this value (DESIGN.md 6.1); frac / ceiling = share of the FP64 issue slots used"}
                                      if tiles else None),
                    "step": {"achieved": flop_step / (ms / args.steps * 1e-3) / 1e12, "peak": fp64_tf,
                             "frac": flop_step / (ms / args.steps * 1e-3) / 1e12 / fp64_tf,
                             "algorithmic_flop_per_step": flop_step},
                    "fp32_peak_tflops": fp32_tf}
        hbm_ach = bytes_step / (ms / args.steps * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": hbm_ach / peaks["hbm_gbs"], "peak_source": peaks_kind + " (MEASURED_PEAKS.json)",
                        "note": "compulsory bytes of a whole step / step time; the path is FP64-pipe bound, not HBM bound",
                        "traffic": prof.get("step_dram_bytes")}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, kind, sample, _ = cpu_reference_rate(budget_s=12.0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(world),
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_x + 3 * R * 8,
                        "d2h_bytes_per_step": bytes_x,
                        "api": "ehic_evaluate_host (C ABI) with pinned host buffers, batched over the GPU's replicas"},
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "re_sweeps": sweeps,
                "fp32_mode": fp32_mode, "other_workloads": other or None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other named configurations (nora2012, proteins, E)")
    ap.add_argument("--workloads", default="nora2012,proteins,E", help="other named configurations measured beside D")
    ap.add_argument("--ladder-iters", type=int, default=20, help="RE iterations timed per ladder workload")
    ap.add_argument("--sweeps", type=int, default=10, help="RE iterations timed for the sweeps/s figure (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
