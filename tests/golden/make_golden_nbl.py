#!/usr/bin/env python
"""Generates tests/golden/nbl_reference.npz: neighbour lists and PROLSQ energies / gradients produced by the REFERENCE'S
OWN C code (ensemble_hic/c/nblist.c, forcefield.c, prolsq.c compiled unmodified into oracle/_ref/libref_nbl.so by
oracle/build_ref.build_nbl) on seeded inputs.  Run in the build container, where /root/reference is mounted; the GPU box
has only the committed .npz.  The cases: uniform radii (every shipped configuration), mixed radii, a compact globule
(many contacts per bead), a chain folded onto a grid boundary, and the seeded structures of the nora2012 fixture."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build_ref                      # noqa: E402
from oracle.oracle import RefNBL                  # noqa: E402
from helpers import random_walk                   # noqa: E402


def cases():
    rng = np.random.default_rng(2016)
    out = []
    out.append(("chain_uniform", random_walk(rng, 1, 120, 0.8)[0], np.full(120, 0.5), 11.4))
    out.append(("chain_mixed", random_walk(rng, 1, 300, 0.9)[0], rng.uniform(0.35, 0.65, 300), 50.0))
    x = rng.normal(size=(200, 3)); x *= 2.2 * rng.uniform(size=(200, 1)) ** (1.0 / 3.0) / np.linalg.norm(x, axis=1, keepdims=True)
    out.append(("globule", x, np.full(200, 0.5), 500.0))
    out.append(("far_from_origin", random_walk(rng, 1, 80, 0.7)[0] + np.array([1000.0, -500.0, 250.0]), np.full(80, 0.5), 3.0))
    g = np.load(os.path.join(HERE, "nora2012.npz"))
    out.append(("nora2012_structure0", g["X"][0].copy(), g["radii"].copy(), float(g["k_nb"])))
    return out


def main():
    assert build_ref.build_nbl(verbose=True), "needs /root/reference"
    ref = RefNBL.load()
    store = {"names": np.array([c[0] for c in cases()])}
    for name, x, radii, k in cases():
        x = np.ascontiguousarray(x, dtype=np.float64)
        cell = float(np.add.outer(radii, radii).max() + 1e-3)                 # forcefields.py:199-201
        tot, pairs = ref.pairs(x, cell, 90, len(x))
        E, F, Eg = ref.energy_gradient(x, radii, k)
        store.update({name + "_x": x, name + "_radii": radii, name + "_k": np.float64(k), name + "_cell": np.float64(cell),
                      name + "_total": np.int64(tot), name + "_pairs": pairs, name + "_energy": np.float64(E),
                      name + "_gradient": F, name + "_energy_from_gradient": np.float64(Eg)})
        print("%-22s n=%4d contacts=%5d E=%.12g" % (name, len(x), tot, E))
    np.savez_compressed(os.path.join(HERE, "nbl_reference.npz"), **store)


if __name__ == "__main__":
    main()
