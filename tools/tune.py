"""Sweeps (EHIC_KT, EHIC_GW) for the gather kernel on config D (development aid)."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import random_walk
from ensemble_hic_b200.engine import Model

N = int(os.environ.get("QB_N", 2000)); S = int(os.environ.get("QB_S", 100)); R = int(os.environ.get("QB_R", 8))
rng = np.random.default_rng(0)
i, j = np.triu_indices(N, 2); M = len(i)
data = np.stack([i, j, rng.poisson(2.0, size=M)], 1).astype(np.int64)
cds = np.full(M, 1.25); radii = np.full(N, 0.5)
dev = torch.device("cuda:0")
x = torch.tensor(np.stack([random_walk(rng, S, N, 0.9) for _ in range(R)]), device=dev)
norm = torch.full((R,), 3.0, dtype=torch.float64, device=dev)
g = torch.empty_like(x)
combos = [c.split(":") for c in os.environ.get("COMBOS", "4:4,4:7,2:4,2:8,2:10,1:10,1:16").split(",")]
ref = None
for kt, gw in combos:
    os.environ["EHIC_KT"] = kt; os.environ["EHIC_GW"] = gw
    m = Model(S, radii, data, cds, alpha=10.0, k_nb=500.0, k_bb=500.0, max_replicas=R)
    def run(terms):
        m.evaluate_device(R, x, norm, None, None, grad=g, terms=terms)
    for terms in (1,):
        run(terms); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(terms)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gc = g.clone()
        if ref is None: ref = gc
        err = float((gc - ref).abs().max() / ref.abs().max())
        print("KT=%s GW=%s lik-gradient R=%d: %.3f ms (%.3f ms/replica) relerr-vs-first %.1e" % (kt, gw, R, ms, ms / R, err), flush=True)
    m.close()
