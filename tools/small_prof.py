import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ensemble_hic_b200.engine import Model
dev = torch.device("cuda:0")
name = os.environ.get("SB_NAME", "nora2012"); R = int(os.environ.get("SB_R", 38))
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
S, N = int(g["S"]), int(g["N"])
m = Model(S, g["radii"], g["data"], g["cds"], alpha=float(g["alpha"]), k_nb=float(g["k_nb"]), k_bb=float(g["k_bb"]), max_replicas=R)
print(m.info())
deg = np.bincount(g["data"][:, 0], minlength=N) + np.bincount(g["data"][:, 1], minlength=N)
print("row lengths: mean %.1f max %d; slice widths %s" % (deg.mean(), deg.max(), [int(deg[s:s+32].max()) for s in range(0, N, 32)]))
rng = np.random.default_rng(0)
X = np.stack([g["X"] for _ in range(R)]).reshape(R, -1)
t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device=dev)
x = t(X); norm = t(np.full(R, float(g["norm"]))); gr = torch.empty_like(x)
print(m.kernel_times(lambda: m.evaluate_device(R, x, norm, None, None, grad=gr), repeats=10))
