import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from helpers import make_problem, relerr
from oracle.oracle import COracle
from ensemble_hic_b200.engine import Model
co = COracle()
for (S, N, dens) in ((5, 70, 1.0), (3, 300, 1.0), (9, 100, 0.2)):
    pb = make_problem(seed=S + N, S=S, N=N, density=dens, step=0.7)
    m = Model(S, pb["radii"], pb["data"], pb["cds"], alpha=10.0, k_nb=50.0, k_bb=500.0)
    out = m.evaluate(pb["X"], 3.3, terms=1, md=True)
    g = co.calculate_gradient(pb["X"], 10.0, 3.3, pb["cds"], pb["data"])
    md = co.ensemble_contacts_evaluate(pb["X"], 3.3, pb["cds"], 10.0, pb["data"])
    nb = m.evaluate(pb["X"], 3.3, terms=2)["grad"][0]
    gn = np.concatenate([co.forcefield_gradient(x, pb["radii"], pb["radii"] ** 2, 50.0) for x in pb["X"]])
    print("S=%d N=%d tiles=%s: grad relerr %.2e, md max elementwise rel %.2e, nb grad relerr %.2e"
          % (S, N, m.info()["path"] == "tiles", relerr(out["grad"][0], g), np.max(np.abs(out["md"][0] - md) / md), relerr(nb, gn)))
