"""The multi-rank replica exchange ON THE CUDA ENGINE: two ranks (three and two temperatures, both on cuda:0) must
reproduce the single-rank run of the same ladder -- identical exchange decisions, acceptance counts and adapted step
sizes, states and norms per slot equal to rounding.  Complements tests/test_rex_gloo.py (host logic on a toy backend) and the `rex_consistent`
check of bench.py (NCCL, one rank per GPU)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(out, world, T=5, iters=31):
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK="0")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "rex_gpu_worker.py"), out, str(T), str(iters)], env=e))
    for p in procs:
        assert p.wait(timeout=600) == 0
    return [json.load(open(os.path.join(out, "result_rank%d.json" % r))) for r in range(world)]


def test_two_ranks_on_the_cuda_engine_reproduce_one_rank(tmp_path):
    one = _run(str(tmp_path / "w1"), 1)[0]
    two = _run(str(tmp_path / "w2"), 2)
    assert [two[0]["lo"], two[0]["hi"], two[1]["lo"], two[1]["hi"]] == [0, 3, 3, 5]
    for r in two:
        for k in ("temp_of_slot", "timestep_T", "n_acc_T", "re_acc", "re_try"):
            assert r[k] == one[k], k
        assert r["n_swap_rounds"] == one["n_swap_rounds"] == 10 and r["n_collectives"] == one["n_collectives"]
    assert two[0]["consistency"] == two[1]["consistency"] == one["consistency"]
    x2 = np.concatenate([np.array(r["x"]) for r in two]); n2 = np.concatenate([np.array(r["norm"]) for r in two])
    # states: the per-replica kernels do not depend on the batch, torch's row sums (mock-count totals of the norm draw) may
    # pick another summation order for another batch shape -> last-bit differences of the norms, carried along
    np.testing.assert_allclose(x2, np.array(one["x"]), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(n2, np.array(one["norm"]), rtol=1e-10)
    # the run did something: trajectories were accepted and rejected, exchanges were tried and some accepted
    assert 0 < sum(one["n_acc_T"]) and sum(one["re_try"]) > 0 and sum(one["re_acc"]) > 0
    assert np.all(np.isfinite(one["x"])) and len(set(one["timestep_T"])) > 1
