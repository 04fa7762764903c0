"""Batched multi-GPU replica exchange: one process per GPU, the replica ladder partitioned
contiguously over the ranks, every local replica advanced by ONE batched device call per HMC
trajectory, exchanges decided from ENERGIES ONLY.

Reference being replaced: one MPI process per replica + a master
(run_simulation.py:17-21, 49-108) that ships whole states between partner processes and
re-evaluates the full posterior on the foreign state (replica.py:12-14).  Here tempering is
linear in two untempered terms,

    E_t(state) = lammda_t * U_L(state) + beta_t * U_nb(state) + U_rest(state),

so the work of exchanging temperatures t and t+1 held by slots a and b is

    w = (lammda_t - lammda_t+1) * (U_L(b) - U_L(a)) + (beta_t - beta_t+1) * (U_nb(b) - U_nb(a))

(U_rest cancels).  Each exchange round all-gathers (U_L, U_nb) of every slot -- 16 bytes per
replica over NCCL -- and every rank takes the same accept/reject decisions from a shared seeded
uniform stream, then permutes the temperature -> slot map.  Coordinates never move; what
follows the *temperature* is (lammda, beta, HMC timestep and its acceptance statistics, and the
sample file), because the reference's outputs are per temperature
(analysis_functions.py:66-69, setup_continue.py:53-54).

The rexfw package (ExchangeMaster, Replica, Slave, statistics writers) is absent from the
reference tree; swap schedule (alternating neighbour pairs every ``swap_interval``-th step),
acceptance rule min(1, exp(-w)) and the file layout are restated from the reference's call
sites and from its continuation / analysis scripts: "parity unpinned" (SURVEY.md 8c).
"""
import os
import pickle
import shutil

import numpy as np

from ._binf import BinfState


def partition(n_replicas, world_size):
    """contiguous blocks of the ladder: rank r holds slots [bounds[r], bounds[r+1])"""
    base, extra = divmod(n_replicas, world_size)
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(int)


def swap_pairs(n_replicas, round_index):
    """neighbouring temperature pairs of one exchange round: (0,1),(2,3).. on even rounds,
    (1,2),(3,4).. on odd rounds (rexfw create_default_RE_params)."""
    start = round_index % 2
    return [(t, t + 1) for t in range(start, n_replicas - 1, 2)]


def decide_swaps(pairs, slot_of_temp, U, lammda, beta, uniforms):
    """Accept/reject every pair from the gathered energies.  U[slot] = (U_L, U_nb).
    Returns (list of accepted flags, list of works)."""
    acc, works = [], []
    for (t0, t1), u in zip(pairs, uniforms):
        a, b = slot_of_temp[t0], slot_of_temp[t1]
        w = (lammda[t0] - lammda[t1]) * (U[b, 0] - U[a, 0]) + (beta[t0] - beta[t1]) * (U[b, 1] - U[a, 1])
        ok = bool(np.isfinite(w) and np.log(max(u, 1e-300)) < -w)
        acc.append(ok); works.append(float(w))
    return acc, works


class EngineBackend(object):
    """Device side of the driver: the resident CUDA model (product path)."""

    def __init__(self, model, device=None):
        import torch
        self.torch = torch
        self.m = model
        self.device = torch.device("cuda", model.device) if device is None else device
        self.n_dof = model.S * model.N * 3

    def tensor(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)

    def upload(self, a):
        """host -> device without a stream synchronisation: pinned staging + asynchronous copy (the pinned block
        is recycled by torch's caching host allocator only after the copy has completed)"""
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        return t.to(self.device, non_blocking=True)

    def energies(self, x, norm):
        R = x.shape[0]
        E = self.torch.empty(R, 4, dtype=self.torch.float64, device=self.device)
        with self.torch.cuda.device(self.device):
            self.m.evaluate_device(R, x, norm, None, None, grad=None, energies=E)
        return E

    def trajectory(self, x, p, norm, lam, beta, eps, n_steps):
        R = x.shape[0]
        if getattr(self, "_H", None) is None or self._H.shape[0] != R:      # persistent: part of the graph signature
            self._H = self.torch.empty(R, 4, dtype=self.torch.float64, device=self.device)
        H = self._H
        with self.torch.cuda.device(self.device):
            self.m.hmc_trajectory_device(R, n_steps, x, p, norm, lam, beta, eps, H)
        return H

    def mock_sums(self, x):
        """sum_m fwm(x, norm=1)_m per replica (npsamplers.py:110-116)"""
        R = x.shape[0]
        if getattr(self, "_md", None) is None or self._md.shape[0] != R:
            self._one = self.torch.ones(R, dtype=self.torch.float64, device=self.device)
            self._md = self.torch.empty(R, self.m.M, dtype=self.torch.float64, device=self.device)
        with self.torch.cuda.device(self.device):
            self.m.mock_data_device(R, x, self._one, self._md)
        return self._md.sum(dim=1)

    def select(self, mask, src, dst):
        from .engine import select_device
        with self.torch.cuda.device(self.device):
            select_device(src.shape[0], src.shape[1], mask, src, dst)


class ReplicaExchange(object):
    """The whole simulation loop of run_simulation.py for the slots of this rank.

    backend      EngineBackend (or a test double with the same methods)
    schedule     {'lammda': [T], 'beta': [T]} for the full ladder
    x0, norm0    initial states of the LOCAL slots: [R_local, n_dof], [R_local]
    hmc          dict(timestep, trajectory_length, adaption_limit)
    norm_prior   (shape, rate) of the Gamma prior if norm is sampled, else None
    counts_sum   sum of the contact counts (shape of the conditional Gamma)

    Between two exchange rounds a rank needs nothing from the others and nothing from its own device:
    the Metropolis test, the step-size adaptation of the temperatures it holds and the conjugate Gamma
    draw run on the device (the random numbers they consume -- one uniform and one standard Gamma
    variate per slot, whose shape lammda * sum(counts) + a does not depend on the state -- are drawn on
    the host ahead of the trajectory and uploaded asynchronously).  The only collective is ONE
    all-gather per exchange round (reference schedule: every ``swap_interval``-th iteration,
    scripts/run_simulation.py:64-69) of four doubles per slot: (U_L, U_nb) for the swap decisions plus
    the slot's current step size and its acceptance count since the last round, which is how the
    replicated per-temperature bookkeeping (mcmc_stats.txt) stays identical on every rank.
    """

    def __init__(self, backend, schedule, x0, norm0, hmc, norm_prior=None, counts_sum=0.0,
                 output_folder=None, seed=0, rank=0, world_size=1, dist=None, momenta="numpy",
                 stats_folder=None):
        self.be = backend
        self.lam_T = np.asarray(schedule["lammda"], dtype=np.float64)
        self.bet_T = np.asarray(schedule["beta"], dtype=np.float64)
        self.T = len(self.lam_T)
        self.rank, self.world, self.dist = rank, world_size, dist
        self.bounds = partition(self.T, world_size)
        self.lo, self.hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.R = self.hi - self.lo
        assert np.shape(x0)[0] == self.R, "x0 must hold the local slots"
        self.x = backend.tensor(np.array(x0, dtype=np.float64))
        self.norm = backend.tensor(np.array(np.broadcast_to(norm0, (self.R,)), dtype=np.float64))
        self.L = int(hmc["trajectory_length"])
        self.adaption_limit = int(hmc.get("adaption_limit", 0))
        self.up, self.down = float(hmc.get("adaption_uprate", 1.05)), float(hmc.get("adaption_downrate", 0.95))
        # per-TEMPERATURE quantities (identical on every rank after every exchange round / sync_stats())
        ts = np.asarray(hmc["timestep"], dtype=np.float64)      # scalar, or one timestep per temperature (continuation)
        self.timestep_T = np.full(self.T, float(ts)) if ts.ndim == 0 else ts.astype(np.float64).copy()
        assert len(self.timestep_T) == self.T
        self.n_acc_T = np.zeros(self.T); self.n_hmc_T = np.zeros(self.T)
        self.re_acc = np.zeros(max(self.T - 1, 1)); self.re_try = np.zeros(max(self.T - 1, 1))
        self.temp_of_slot = np.arange(self.T)
        self.slot_of_temp = np.arange(self.T)
        self.norm_prior, self.counts_sum = norm_prior, float(counts_sum)
        self.out = output_folder
        self.stats_out = stats_folder if stats_folder is not None else (
            None if output_folder is None else os.path.join(output_folder, "statistics"))
        self.momenta = momenta
        self.rng_swap = np.random.RandomState(seed)                       # shared by all ranks
        self.rng_slot = [np.random.RandomState(seed + 1 + s) for s in range(self.lo, self.hi)]
        self.n_swap_rounds = 0
        self.step = 0
        self._buffer = {}          # temperature -> [(step, BinfState)]
        self._win_count = 0        # iterations since the last dump (for the dump_step thinning)
        self.works = []
        self.n_collectives = 0
        self._upload = getattr(backend, "upload", backend.tensor)
        # device-resident per-slot state: persistent tensors (their pointers are part of the trajectory graph's
        # signature), updated in place
        z = np.zeros(self.R)
        self._lam_d, self._bet_d, self._eps_d = backend.tensor(z), backend.tensor(z), backend.tensor(z)
        self._acc_d = backend.tensor(z)                                   # accepted trajectories since the last sync
        self._x_work = backend.torch.empty_like(self.x)
        self._p_work = backend.torch.empty_like(self.x)
        self.last_accept = None
        self._push_temps()

    # ---- helpers ---------------------------------------------------------------------------------
    def _local_temps(self):
        return self.temp_of_slot[self.lo:self.hi]

    def _push_temps(self):
        """(lammda, beta, step size) of the temperatures now held by the local slots -> device"""
        temps = self._local_temps()
        packed = self._upload(np.stack([self.lam_T[temps], self.bet_T[temps], self.timestep_T[temps]]))
        self._lam_d.copy_(packed[0]); self._bet_d.copy_(packed[1]); self._eps_d.copy_(packed[2])

    def _allgather_rows(self, local, width):
        """gather [R_local, width] rows of every rank into a [T, width] numpy array (synchronises)"""
        torch = self.be.torch if hasattr(self.be, "torch") else None
        self.n_collectives += 1
        if self.world == 1 or self.dist is None:
            return local.detach().cpu().numpy() if hasattr(local, "detach") else np.asarray(local)
        rmax = int(np.max(np.diff(self.bounds)))
        pad = torch.zeros(rmax, width, dtype=local.dtype, device=local.device)
        pad[:self.R] = local
        if pad.is_cuda and self.dist.get_backend() == "gloo":      # gloo gathers host tensors (tests: two ranks on one GPU)
            pad = pad.cpu()
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        rows = [out[r][:self.bounds[r + 1] - self.bounds[r]] for r in range(self.world)]
        return torch.cat(rows).cpu().numpy()

    def _fold_stats(self, eps_all, acc_all):
        """per-slot step sizes and acceptance counts since the last sync -> per-temperature tables"""
        t = self.temp_of_slot
        self.timestep_T[t] = eps_all
        self.n_acc_T[t] += acc_all
        self._acc_d.zero_()

    def sync_stats(self):
        """bring timestep_T / n_acc_T up to date on every rank (a collective; called by run() when statistics
        are written or printed and at its end -- exchange rounds do the same as a by-product)"""
        torch = self.be.torch
        G = self._allgather_rows(torch.stack([self._eps_d, self._acc_d], 1), 2)
        self._fold_stats(G[:, 0], G[:, 1])

    # ---- one Gibbs sample for every local slot ------------------------------------------------------
    def sample(self):
        be, torch, temps = self.be, self.be.torch, self._local_temps()
        # host side: everything random, none of it depends on the device.  Per-slot streams, drawn in the
        # order momenta, uniform, gamma.
        if self.momenta == "numpy":
            self._p_work.copy_(self._upload(np.stack([rs.normal(size=be.n_dof) for rs in self.rng_slot])))
        else:
            self._p_work.normal_()
        u = np.array([rs.random_sample() for rs in self.rng_slot])
        self.n_hmc_T += 1
        adapt = self.n_hmc_T[temps] < self.adaption_limit
        host = np.zeros((5, self.R))
        host[0] = np.log(np.maximum(u, 1e-300))
        host[1] = np.where(adapt, self.up, 1.0); host[2] = np.where(adapt, self.down, 1.0)
        if self.norm_prior is not None:
            shape = self.lam_T[temps] * self.counts_sum + self.norm_prior[0]
            host[3] = [rs.gamma(sh) for rs, sh in zip(self.rng_slot, shape)]
        d = self._upload(host)
        # device side: trajectory, Metropolis test, step-size adaptation, conjugate Gamma draw
        x_new, p = self._x_work, self._p_work
        x_new.copy_(self.x)
        H = be.trajectory(x_new, p, self.norm, self._lam_d, self._bet_d, self._eps_d, self.L)
        dH = (H[:, 2] + H[:, 3]) - (H[:, 0] + H[:, 1])
        acc = torch.isfinite(dH) & (d[0] < -dH)
        be.select(acc.to(dtype=torch.int32), x_new, self.x)
        self._acc_d += acc.to(dtype=torch.float64)
        self._eps_d *= torch.where(acc, d[1], d[2])
        self.last_accept = acc
        if self.norm_prior is not None:       # norm = Gamma(shape) / rate (npsamplers.py:42-116)
            rate = self._lam_d * be.mock_sums(self.x) + self.norm_prior[1]
            v = d[3] / rate
            self.norm.copy_(torch.where(v != 0.0, v, torch.full_like(v, 1e-10)))
        return acc

    # ---- one exchange round ----------------------------------------------------------------------------
    def exchange(self):
        torch = self.be.torch
        E = self.be.energies(self.x, self.norm)
        G = self._allgather_rows(torch.stack([E[:, 0], E[:, 1], self._eps_d, self._acc_d], 1), 4)
        self._fold_stats(G[:, 2], G[:, 3])
        U = G[:, :2]                                                  # [T slots, (U_L, U_nb)]
        pairs = swap_pairs(self.T, self.n_swap_rounds)
        uniforms = self.rng_swap.random_sample(len(pairs))
        acc, works = decide_swaps(pairs, self.slot_of_temp, U, self.lam_T, self.bet_T, uniforms)
        for (t0, t1), ok in zip(pairs, acc):
            self.re_try[t0] += 1
            if ok:
                self.re_acc[t0] += 1
                a, b = self.slot_of_temp[t0], self.slot_of_temp[t1]
                self.slot_of_temp[t0], self.slot_of_temp[t1] = b, a
                self.temp_of_slot[a], self.temp_of_slot[b] = t1, t0
        if self.out is not None:
            self.works.append((self.step, pairs, works, acc))
        self.n_swap_rounds += 1
        self._push_temps()
        return acc

    def consistency(self):
        """what must be identical on every rank (replicated bookkeeping): compared by bench.py / tests through
        all_gather_object"""
        return dict(temp_of_slot=self.temp_of_slot.tolist(), timestep_T=self.timestep_T.tolist(),
                    n_acc_T=self.n_acc_T.tolist(), re_acc=self.re_acc.tolist(), re_try=self.re_try.tolist(),
                    n_swap_rounds=self.n_swap_rounds)

    # ---- output -------------------------------------------------------------------------------------------
    def _record(self, dump_step):
        """Every iteration (Gibbs sample or exchange round: rexfw is absent from the reference tree, so which of
        the two it stores is unpinned; analysis_functions.load_sr_samples makes no assumption on the number of
        states per file) adds the current state to the dump window; a dump keeps every dump_step-th of them.
        Only those are copied off the device, and none at all without an output folder."""
        keep = self._win_count % max(int(dump_step), 1) == 0
        self._win_count += 1
        if self.out is None or not keep:
            return
        xs = self.x.detach().cpu().numpy()
        ns = self.norm.detach().cpu().numpy()
        for q, t in enumerate(self._local_temps()):
            v = {"structures": xs[q].copy()}
            if self.norm_prior is not None:
                v["norm"] = float(ns[q])
            self._buffer.setdefault(int(t), []).append((self.step, BinfState(v)))

    def _dump(self, first, last):
        """samples/samples_replica{t+1}_{first}-{last}.pickle: list of states (already thinned by dump_step).
        Records of temperatures whose home rank is another rank travel there now (I/O, not hot path)."""
        self._win_count = 0
        if self.out is None:
            self._buffer = {}
            return
        home = lambda t: int(np.searchsorted(self.bounds, t, side="right") - 1)
        mine, away = {}, {}
        for t, recs in self._buffer.items():
            (mine if home(t) == self.rank else away).setdefault(t, []).extend(recs)
        if self.world > 1 and self.dist is not None:
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, away)
            for other in gathered:
                for t, recs in other.items():
                    if home(t) == self.rank:
                        mine.setdefault(t, []).extend(recs)
        for t, recs in mine.items():
            recs.sort(key=lambda r: r[0])
            states = [s for (st, s) in recs]
            fn = os.path.join(self.out, "samples", "samples_replica%d_%d-%d.pickle" % (t + 1, first, last))
            with open(fn, "wb") as f:
                pickle.dump(states, f, protocol=2)
        self._buffer = {}

    def _write_stats(self):
        if self.out is None or self.rank != 0:
            return
        p_acc = self.n_acc_T / np.maximum(self.n_hmc_T, 1)
        row = [str(self.step)]
        for t in range(self.T):
            row += ["%.6f" % p_acc[t], "%.6e" % self.timestep_T[t]]
        with open(os.path.join(self.stats_out, "mcmc_stats.txt"), "a") as f:
            f.write("\t".join(row) + "\n")
        re_rate = self.re_acc / np.maximum(self.re_try, 1)
        with open(os.path.join(self.stats_out, "re_stats.txt"), "a") as f:
            f.write("\t".join([str(self.step)] + ["%.6f" % v for v in re_rate[:max(self.T - 1, 0)]]) + "\n")
        if self.works:
            with open(os.path.join(self.out, "works", "works_%d.pickle" % self.step), "wb") as f:
                pickle.dump(self.works, f, protocol=2)
            self.works = []

    def status_line(self):
        p_acc = self.n_acc_T / np.maximum(self.n_hmc_T, 1)
        re_rate = self.re_acc / np.maximum(self.re_try, 1)
        return ("step %d  HMC p_acc %s  stepsize %s  RE p_acc %s"
                % (self.step, np.array2string(p_acc, precision=2), np.array2string(self.timestep_T, precision=2),
                   np.array2string(re_rate[:max(self.T - 1, 0)], precision=2)))

    # ---- the loop (rexfw ExchangeMaster.run as driven by run_simulation.py:65-70) ----------------------------
    def run(self, n_iterations, swap_interval=5, status_interval=100, dump_interval=250, dump_step=5,
            statistics_update_interval=100, offset=0, verbose=False):
        first = offset
        for it in range(offset, offset + int(n_iterations)):
            self.step = it
            swapped = it % int(swap_interval) == 0 and it > 0
            if swapped:
                self.exchange()
            else:
                self.sample()
            self._record(dump_step)
            # a continued run starts at it == offset > 0: nothing has been sampled yet, nothing is written
            want_stats = self.out is not None and it % int(statistics_update_interval) == 0 and it > offset
            want_status = verbose and it % int(status_interval) == 0 and it > offset
            if (want_stats or want_status) and not swapped:
                self.sync_stats()
            if want_stats:
                self._write_stats()
            if want_status and self.rank == 0:
                print(self.status_line(), flush=True)
            if it % int(dump_interval) == 0 and it > offset:
                self._dump(first, it)
                first = it
        self.sync_stats()
        return self


class ExchangeMaster(object):
    """Object returned by setup_default_re_master: remembers the ladder size and output folder;
    ``run`` mirrors ``master.run(n, swap_interval=..., ...)`` of run_simulation.py:65-70 once a
    :class:`ReplicaExchange` has been bound."""

    def __init__(self, n_replicas, sim_path):
        self.n_replicas = n_replicas
        self.sim_path = sim_path if sim_path.endswith("/") else sim_path + "/"
        self.rex = None

    def bind(self, rex):
        self.rex = rex
        return self

    def run(self, n_iterations, swap_interval=5, status_interval=100, dump_interval=250, dump_step=5,
            statistics_update_interval=100, offset=0, verbose=False):
        if self.rex is None:
            raise RuntimeError("bind() a ReplicaExchange first (see run_simulation.main)")
        return self.rex.run(n_iterations, swap_interval, status_interval, dump_interval, dump_step,
                            statistics_update_interval, offset, verbose)

    def terminate_replicas(self):
        pass


def load_continuation(output_folder, n_replicas, dump_interval, n_cont=1):
    """What scripts/continue_simulation/setup_continue.py extracts from a finished run: the offset
    (first sample index without a dump file), the last dumped state of every temperature, and the
    HMC timesteps from the last row of mcmc_stats.txt (columns 2::2).  Also writes the reference's
    artefacts init_states.npy / init_norms.npy / timesteps.npy into init_continue{n_cont}/."""
    out = output_folder if output_folder.endswith("/") else output_folder + "/"
    fname = out + "samples/samples_replica{}_{}-{}.pickle"
    offset = 0
    while os.path.exists(fname.format(1, offset, offset + dump_interval)):
        offset += dump_interval
    if offset == 0:
        raise IOError("no dumped samples under %ssamples/: nothing to continue from" % out)
    states = []
    for t in range(1, n_replicas + 1):
        with open(fname.format(t, offset - dump_interval, offset), "rb") as f:
            states.append(pickle.load(f)[-1])
    structures = np.array([s.variables["structures"] for s in states])
    norms = np.array([s.variables["norm"] for s in states]) if "norm" in states[0].variables else None
    stats_dir = out + ("statistics/" if n_cont == 1 else "init_continue%d/statistics/" % (n_cont - 1))
    timesteps = None
    if os.path.exists(stats_dir + "mcmc_stats.txt"):
        timesteps = np.atleast_2d(np.loadtxt(stats_dir + "mcmc_stats.txt"))[-1, 2::2]
    cont = out + "init_continue%d/" % n_cont
    os.makedirs(cont + "statistics", exist_ok=True)
    np.save(cont + "init_states.npy", structures)
    if norms is not None:
        np.save(cont + "init_norms.npy", norms)
    if timesteps is not None:
        np.save(cont + "timesteps.npy", timesteps)
    return dict(offset=offset, structures=structures, norms=norms, timesteps=timesteps, cont_folder=cont)


def copy_run_metadata(config_file, schedule, output_folder):
    """config.cfg and schedule.pickle next to the samples (run_simulation.py:56-59)"""
    os.makedirs(output_folder, exist_ok=True)
    shutil.copy2(config_file, os.path.join(output_folder, "config.cfg"))
    with open(os.path.join(output_folder, "schedule.pickle"), "wb") as f:
        pickle.dump({k: np.asarray(v) for k, v in schedule.items()}, f, protocol=2)
