#!/usr/bin/env python
"""Copies the replica ladders of the reference's shipped configurations into tests/golden/schedules.npz (run in the
build container, where /root/reference is mounted; the GPU box has only the committed file).  bench.py runs the
nora2012 and proteins workloads on exactly these ladders:
  config_files/nora2012/30structures_3kb_schedule.pickle   302 temperatures, lammda = beta, 1e-6 .. 1
  config_files/proteins/2structures_schedule.pickle         54 temperatures, lammda = beta, 1e-6 .. 1
"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/config_files"
SRC = {"nora2012": "nora2012/30structures_3kb_schedule.pickle", "proteins": "proteins/2structures_schedule.pickle"}

out = {}
for name, rel in SRC.items():
    with open(os.path.join(REF, rel), "rb") as f:
        d = pickle.load(f, encoding="latin1")
    out[name + "_lammda"] = np.asarray(d["lammda"], dtype=np.float64)
    out[name + "_beta"] = np.asarray(d["beta"], dtype=np.float64)
    print(name, len(d["lammda"]), d["lammda"][0], d["lammda"][-1])
np.savez_compressed(os.path.join(HERE, "schedules.npz"), **out)
