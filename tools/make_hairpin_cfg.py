"""Writes a runnable hairpin_s-like config (data file from the golden fixture) into argv[1]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_host_logic import CFG
out = sys.argv[1]; n_rep = sys.argv[2] if len(sys.argv) > 2 else "2"
os.makedirs(out, exist_ok=True)
g = np.load(os.path.join(ROOT, "tests", "golden", "hairpin_s.npz"))
rows = [tuple(r) for r in g["data"]]
open(os.path.join(out, "data.txt"), "w").write("\n".join("%d\t%d\t%d" % r for r in rows) + "\n")
cfg = (CFG % {"data": os.path.join(out, "data.txt")}).replace("/tmp/hairpin_s/", os.path.join(out, "sim") + "/")
cfg = cfg.replace("n_replicas = 2", "n_replicas = " + n_rep).replace("timestep = 1e-4", "timestep = 1e-3")
open(os.path.join(out, "config.cfg"), "w").write(cfg)
print(os.path.join(out, "config.cfg"))
