"""Ad-hoc timing of the device entry points (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import random_walk
from ensemble_hic_b200.engine import Model, launch_count
from fma_peak import fma_peak

def main():
    N = int(os.environ.get("QB_N", 2000)); S = int(os.environ.get("QB_S", 100)); R = int(os.environ.get("QB_R", 2))
    rng = np.random.default_rng(0)
    i, j = np.triu_indices(N, 2)
    M = len(i)
    radii = np.full(N, 0.5)
    if os.environ.get("QB_COMPACT"):
        # compact globules: reflecting random walk inside a ball holding ~1 bead per unit volume
        rad = 0.62 * N ** (1 / 3.0)
        X = np.zeros((R, S, N, 3))
        for r_ in range(R):
            for k in range(S):
                p = np.zeros(3)
                steps = rng.normal(size=(N, 3)); steps *= 0.9 / np.linalg.norm(steps, axis=1, keepdims=True)
                for bi in range(N):
                    q = p + steps[bi]
                    if np.linalg.norm(q) > rad: q = p - steps[bi]
                    X[r_, k, bi] = p = q
    else:
        X = np.stack([random_walk(rng, S, N, 0.9) for _ in range(R)])
    cds = np.full(M, 1.25)
    d = np.linalg.norm(X[0][:, i] - X[0][:, j], axis=2) if N <= 500 else None
    counts = rng.poisson(2.0, size=M)
    data = np.stack([i, j, counts], 1).astype(np.int64)
    t0 = time.time()
    m = Model(S, radii, data, cds, alpha=10.0, k_nb=500.0, k_bb=500.0, max_replicas=R, fp32=bool(os.environ.get("QB_FP32")))
    print("model create %.2fs" % (time.time() - t0), m.info(), flush=True)
    for fp64 in (True, False):
        print("fma peak fp64=%s: %.2f TFLOP/s (%.3f ms/launch)" % ((fp64,) + fma_peak(fp64, 10)), flush=True)
    dev = torch.device("cuda:0")
    x = torch.tensor(X, device=dev); norm = torch.full((R,), 3.0, dtype=torch.float64, device=dev)
    lam = torch.ones(R, dtype=torch.float64, device=dev); bet = torch.ones(R, dtype=torch.float64, device=dev)
    g = torch.empty_like(x); E = torch.zeros(R, 4, dtype=torch.float64, device=dev)
    md = torch.empty(R, M, dtype=torch.float64, device=dev)
    def timeit(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for terms, name in ((15, "all"), (1, "lik"), (2, "nb"), (12, "bb+sph")):
        ms = timeit(lambda: m.evaluate_device(R, x, norm, lam, bet, grad=g, terms=terms))
        print("gradient[%s] R=%d: %.3f ms  -> %.1f grad evals/s" % (name, R, ms, R / ms * 1e3), flush=True)
    ms = timeit(lambda: m.mock_data_device(R, x, norm, md))
    print("mock_data R=%d: %.3f ms" % (R, ms), flush=True)
    ms = timeit(lambda: m.evaluate_device(R, x, norm, lam, bet, energies=E))
    print("energies R=%d: %.3f ms" % (R, ms), flush=True)
    print("launches", launch_count())

if __name__ == "__main__":
    main()
