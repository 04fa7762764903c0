"""Config E: 10 000 beads x 200 structures, sparse list M ~ 1e7 drawn with P(i,j) ~ 1/|i-j|.
Times one posterior gradient for R replicas (development aid; bench.py measures config D)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import random_walk
from ensemble_hic_b200.engine import Model

N = int(os.environ.get("E_N", 10000)); S = int(os.environ.get("E_S", 200)); M = int(float(os.environ.get("E_M", 1e7))); R = int(os.environ.get("E_R", 2))
rng = np.random.default_rng(0)
t0 = time.time()
keys = np.zeros(0, dtype=np.int64)
while len(keys) < M:
    n = int(1.6 * (M - len(keys))) + 1000
    sep = np.minimum(np.exp(rng.uniform(np.log(2), np.log(N - 1), size=n)).astype(np.int64), N - 1)
    i = rng.integers(0, N, size=n); j = i + sep
    ok = j < N
    keys = np.unique(np.concatenate([keys, i[ok] * N + j[ok]]))
keys = rng.permutation(keys)[:M]
data = np.ascontiguousarray(np.stack([keys // N, keys % N, rng.poisson(3.0, size=M)], 1).astype(np.int64))
print("list built %.1fs: M=%d density %.3f" % (time.time() - t0, M, M / (N * (N - 1) / 2)), flush=True)
t0 = time.time()
m = Model(S, np.full(N, 0.5), data, np.full(M, 1.25), alpha=10.0, k_nb=500.0, k_bb=500.0, max_replicas=R)
print("model %.1fs" % (time.time() - t0), m.info(), flush=True)
dev = torch.device("cuda:0")
x = torch.tensor(np.stack([random_walk(rng, S, N, 0.9) for _ in range(R)]), device=dev)
norm = torch.full((R,), 3.0, dtype=torch.float64, device=dev); g = torch.empty_like(x)
run = lambda: m.evaluate_device(R, x, norm, None, None, grad=g)
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("config E gradient R=%d: %.2f ms -> %.1f grad evals/s; kernels %s" % (R, ms, R / ms * 1e3, m.kernel_times(run, 3)), flush=True)
print("GPU memory in use: %.1f GB" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30))
