"""Batch re-evaluation of stored samples.

The reference's density-of-states analysis (analysis_functions.py:379-490, ``calculate_DOS``) walks the
stored samples of every replica and evaluates, one sample and one component at a time,
``-L.log_prob(**x.variables)`` and ``-P.log_prob(structures=...)`` -- 10^4..10^5 posterior evaluations in
log_prob-only mode -- before handing the resulting energy table to WHAM (which stays on the CPU and is
not part of this package).  Here the same table comes from the resident GPU model, ``max_replicas``
samples per call, with the energies-only entry point (forward + prior kernels, no gradient):

    energies, sels = replica_energies("config.cfg", n_samples, subsamples_fraction, burnin)
    q = reweighting_matrix(energies, schedule)          # [n_replicas, n_replicas * n_sel]: the WHAM input

Columns follow analysis_functions.py:452-456: ``[-log L(x, norm), E_nb(x)]`` (untempered).
"""
import os
import pickle

import numpy as np

from . import _lib


def load_sr_samples(samples_folder, replica_id, n_samples, dump_interval, burnin, interval=1):
    """Samples of one replica from the per-interval pickles written by a replica-exchange run (same
    arguments and selection rule as analysis_functions.py:38-73; replica_id is 1-based)."""
    samples = []
    for first in range((burnin // dump_interval) * dump_interval, n_samples - dump_interval, dump_interval):
        fn = os.path.join(samples_folder, "samples_replica%d_%d-%d.pickle" % (replica_id, first, first + dump_interval))
        with open(fn, "rb") as f:
            samples += list(pickle.load(f))
    if burnin < dump_interval:
        start = burnin
    elif burnin > dump_interval:
        start = burnin % dump_interval
    else:
        start = 0
    out = np.empty(len(samples[start::interval]), dtype=object)
    out[:] = samples[start::interval]
    return out


def batch_energies(engine, structures, norms, batch=None):
    """Untempered energies ``[-log L, E_nonbonded, E_backbone, E_sphere]`` of n samples.

    structures: [n, S*N*3]; norms: [n] (or a scalar).  One C-ABI call per ``batch`` samples
    (default: the model's max_replicas)."""
    x = np.ascontiguousarray(structures, dtype=np.float64)
    per = engine.S * engine.N * 3
    x = x.reshape(-1, per)
    n = x.shape[0]
    nrm = np.ascontiguousarray(np.broadcast_to(np.asarray(norms, dtype=np.float64), (n,)))
    batch = int(batch or engine.max_replicas)
    if batch < 1 or batch > engine.max_replicas:
        raise ValueError("batch must be in [1, max_replicas]")
    out = np.empty((n, _lib.N_ENERGIES))
    for a in range(0, n, batch):
        b = min(n, a + batch)
        out[a:b] = engine.evaluate(x[a:b], nrm[a:b], grad=False, energies=True)["energies"]
    return out


def sample_energies(posterior, samples, use_likelihood=True, use_nonbonded=True, batch=None):
    """``[[-L.log_prob(**x.variables), -P.log_prob(structures=x.variables['structures'])] for x in samples]``
    (analysis_functions.py:452-456) through the posterior's resident GPU model.  A column whose term is
    switched off is zero, as in the reference when 'lammda' / 'beta' is not part of the schedule."""
    engine = getattr(posterior, "engine", None)
    if engine is None:
        raise ValueError("the posterior has no resident GPU model (build it with setup_functions.make_posterior)")
    n = len(samples)
    if n == 0:
        return np.zeros((0, 2))
    xs = np.stack([np.asarray(s.variables["structures"], dtype=np.float64).ravel() for s in samples])
    if "norm" in samples[0].variables:
        norms = np.array([float(s.variables["norm"]) for s in samples])
    else:
        norms = np.full(n, float(posterior["norm"].value))
    E = batch_energies(engine, xs, norms, batch)
    out = np.zeros((n, 2))
    if use_likelihood:
        out[:, 0] = E[:, _lib.E_LIKELIHOOD]
    if use_nonbonded:
        out[:, 1] = E[:, _lib.E_NONBONDED]
    return out


def replica_energies(config_file, n_samples, subsamples_fraction, burnin, max_replicas=64, rng=None):
    """The energy table of ``calculate_DOS`` (analysis_functions.py:416-458): for every replica a random
    ``1 / subsamples_fraction`` of its stored samples after ``burnin``; returns
    (energies [n_replicas, n_sel, 2], sels [n_replicas, n_sel])."""
    from .setup_functions import make_posterior, parse_config_file
    settings = parse_config_file(config_file)
    n_replicas = int(settings["replica"]["n_replicas"])
    n_samples = min(int(n_samples), int(settings["replica"]["n_samples"]))
    dump_interval = int(settings["replica"]["samples_dump_interval"])
    out = settings["general"]["output_folder"]
    out = out if out.endswith("/") else out + "/"
    with open(out + "schedule.pickle", "rb") as f:
        schedule = pickle.load(f)
    posterior = make_posterior(settings, max_replicas=max_replicas)
    rng = rng or np.random
    energies, sels = [], []
    for r in range(n_replicas):
        samples = load_sr_samples(out + "samples/", r + 1, n_samples + 1, dump_interval, burnin)
        sel = rng.choice(len(samples), int(len(samples) / float(subsamples_fraction)), replace=False)
        sels.append(sel)
        energies.append(sample_energies(posterior, samples[sel], "lammda" in schedule, "beta" in schedule))
    return np.array(energies), np.array(sels)


def reweighting_matrix(energies, schedule):
    """q[r, s] = lammda_r * (-log L)_s + beta_r * E_nb,s over the flattened samples: what WHAM reweights with
    (analysis_functions.py:460-464)."""
    flat = np.asarray(energies).reshape(-1, 2)
    n_rep = len(np.asarray(schedule["lammda"] if "lammda" in schedule else schedule["beta"]))
    lam = np.asarray(schedule["lammda"], dtype=np.float64) if "lammda" in schedule else np.zeros(n_rep)
    bet = np.asarray(schedule["beta"], dtype=np.float64) if "beta" in schedule else np.zeros(n_rep)
    return lam[:, None] * flat[None, :, 0] + bet[:, None] * flat[None, :, 1]
