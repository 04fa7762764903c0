#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted; the GPU box has only the committed .npz files).

For each shipped configuration of the reference (hairpin_s toy, proteins 1pga_1shf, nora2012
3 kb) the reference's own filtering rules (setup_functions.parse_data / make_likelihood /
make_priors semantics, restated in ensemble_hic_b200.setup_functions and cross-checked here
against an independent inline restatement) produce the kernel inputs; the reference's OWN
compiled Cython (oracle/_ref, built from /root/reference by oracle/build_ref.py) produces the
expected outputs on seeded coordinates.  The neighbour-list forcefield values come from the
plain-C restatement (the reference's Python-2 C extension cannot be built on CPython 3).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import build_ref                      # noqa: E402
from oracle.oracle import COracle, OraclePosterior, load_ref   # noqa: E402

CONFIGS = {
    # name: (cfg, data file, n_structures override or None, seed)
    "hairpin_s": ("scripts/hairpin_s/config.cfg", "data/hairpin_s/hairpin_s_littlenoise.txt", None, 1),
    "proteins": ("config_files/proteins/2structures.cfg", "data/proteins/1pga_1shf/1pga_1shf_fwm_poisson.txt", None, 2),
    "nora2012": ("config_files/nora2012/30structures_3kb.cfg", "data/nora2012/bothdomains.txt", None, 3),
}


def parse_cfg(path):
    import configparser
    c = configparser.ConfigParser(interpolation=None)
    c.read(path)
    return {s: dict(c.items(s)) for s in c.sections()}


def filter_data(path, flt):
    """independent inline restatement of setup_functions.py:562-574"""
    data = np.loadtxt(path, dtype=int)
    if flt["include_zero_counts"] == "False":
        data = data[data[:, 2] > 0]
    data = data[np.argsort(data[:, 2])]
    data = data[int(float(flt["disregard_lowest"]) * len(data)):]
    return data[np.abs(data[:, 0] - data[:, 1]) > int(flt["ignore_sequential_neighbors"])]


def compact_walk(rng, S, N, step, box):
    """seeded chains: random walk with reflecting pull towards the origin (realistic contacts)"""
    X = np.zeros((S, N, 3))
    for k in range(S):
        p = rng.normal(scale=0.3 * box, size=3)
        for i in range(N):
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            q = p + step * d
            if np.linalg.norm(q) > box:
                q = p - step * d
            X[k, i] = p = q
    return X


def main():
    assert build_ref.build(verbose=False), "oracle/_ref could not be built"
    ref = load_ref()
    co = COracle()
    summary = {}
    for name, (cfg_rel, data_rel, S_over, seed) in CONFIGS.items():
        cfg = parse_cfg(os.path.join(REF, cfg_rel))
        g = cfg["general"]
        N, S = int(g["n_beads"]), int(S_over or g["n_structures"])
        data = np.ascontiguousarray(filter_data(os.path.join(REF, data_rel), cfg["data_filtering"]).astype(np.int64))
        r = float(cfg["nonbonded_prior"]["bead_radii"])
        radii = np.ones(N) * r
        cds = (radii[data[:, 0]] + radii[data[:, 1]]) * float(cfg["forward_model"]["contact_distance_factor"])
        alpha = float(cfg["forward_model"]["alpha"])
        k_nb = float(cfg["nonbonded_prior"]["force_constant"])
        k_bb = float(cfg["backbone_prior"]["force_constant"])
        sphere = None
        if cfg["sphere_prior"]["active"] == "True":
            R = cfg["sphere_prior"]["radius"]
            R = 2 * radii.mean() * N ** (1 / 3.0) if R == "auto" else float(R)
            sphere = (R, float(cfg["sphere_prior"]["force_constant"]), np.zeros(3))
        shape, rate = cfg["norm_prior"]["shape"], cfg["norm_prior"]["rate"]
        if shape == rate == "auto":
            rate = 1.0 / S
            shape = float(np.mean(data[:, 2][data[:, 2] > 0]) / float(S))
        gamma = (float(shape), float(rate))
        rng = np.random.default_rng(seed)
        box = 1.5 * r * N ** (1 / 3.0) * 1.2
        X = compact_walk(rng, S, N, 1.9 * r, box)
        norm = float(data[:, 2].max()) / S
        lam, bet = 0.7, 0.4
        ul = radii[:-1] + radii[1:]
        op = OraclePosterior(S, data, cds, alpha, radii, k_nb, k_bb, [np.zeros(N - 1)], [ul], [0, N],
                             sphere=sphere, gamma=gamma, kernels=ref, nonbonded="nbl")
        md = np.asarray(ref["ensemble_contacts_evaluate"](X, norm, cds, alpha, data))
        glik = np.asarray(ref["calculate_gradient"](X, alpha, norm, cds, data))
        ffE = np.array([ref["forcefield_energy"](np.ascontiguousarray(x), radii, radii * radii, k_nb) for x in X])
        ffG = np.concatenate([np.asarray(ref["forcefield_gradient"](np.ascontiguousarray(x), radii, radii * radii, k_nb)) for x in X])
        bbG = np.concatenate([np.asarray(ref["backbone_prior_gradient"](np.ascontiguousarray(x).ravel(), np.zeros(N - 1), ul, k_bb)) for x in X])
        spG = np.zeros(0)
        if sphere is not None:
            spG = np.concatenate([np.asarray(ref["sphere_prior_gradient"](np.ascontiguousarray(x), np.zeros(3), sphere[0], sphere[1],
                                                                            np.arange(N), radii, radii * radii)) for x in X])
        energies = np.array(op.energies(X, norm))
        logp = op.log_prob(X, norm, lam, bet)
        grad = op.gradient(X, norm, lam, bet)
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, N=N, S=S, data=data, cds=cds, radii=radii, alpha=alpha, k_nb=k_nb, k_bb=k_bb,
                            sphere=np.array([sphere[0], sphere[1]]) if sphere else np.zeros(0),
                            gamma=np.array(gamma), X=X, norm=norm, lammda=lam, beta=bet,
                            md=md, grad_lik=glik, ff_energy=ffE, ff_grad=ffG, bb_grad=bbG, sphere_grad=spG,
                            energies=energies, log_prob=logp, grad=grad)
        summary[name] = dict(N=N, S=S, M=int(len(data)), zero_counts=int((data[:, 2] == 0).sum()),
                             max_count=int(data[:, 2].max()), bytes=os.path.getsize(out),
                             n_nb_contacts_struct0=int((np.abs(ffG.reshape(S, N, 3)[0]).sum(1) > 0).sum()))
        print(name, summary[name])
    # the reference's own known-answer vectors (tests/cases/*.py), restated as data
    kat = {
        "forward_models.py:18-46": {
            "structures": [[[0, 0, 0], [0, 0, 1], [0, 0, 2]], [[0, 0, 0], [0, 0, 3], [0, 0, 3]]],
            "contact_distances": [2.0, 1.0], "data_points": [[0, 1, 42], [0, 2, 53]],
            "smooth_steepness": 1.0, "norm": 2.0, "md0": 2.0},
        "backbone_prior.py:38-55": {
            "ll": [[0.0, 0.5], [1.0, 0.0, 0.0]], "ul": [[0.5, 1.0], [2.0, 2.0, 1.0]], "k_bb": 2.0,
            "mol_ranges": [0, 3, 7],
            "structure0_z": [0.0, 1.0, 1.25, 4.0, 4.5, 6.5, 8.5], "structure1_z": [2.0, 4.0, 6.0, 7.0, 8.0, 9.0, 10.0],
            "log_prob_mol0": -0.5 * 2.0 * (0.0625 + 0.25),
            "grad_mol1_z": [2.0 * 0.5, -2.0 * 0.5, -2.0 * 1.0, 2.0 * 1.0]},
        "error_models.py:14-20": {"data": [3, 4], "mock_data_log": [3.0, 4.0]},
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kat, f, indent=1)
    with open(os.path.join(HERE, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
