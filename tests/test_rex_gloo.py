"""The multi-rank path of the replica-exchange driver on CPU: world_size 2 over gloo must
reproduce the single-process run exactly (same swap decisions, same per-temperature step sizes,
same states per slot), and write the reference's per-temperature output files."""
import glob
import json
import os
import pickle
import socket
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(out, world, T=5, iters=31):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="1")
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "rex_worker.py"), out, str(T), str(iters)], env=e))
    for p in procs:
        assert p.wait(timeout=300) == 0
    return [json.load(open(os.path.join(out, "result_rank%d.json" % r))) for r in range(world)]


def test_two_ranks_reproduce_one_rank(tmp_path):
    one = _run(str(tmp_path / "w1"), 1)[0]
    two = _run(str(tmp_path / "w2"), 2)
    assert [two[0]["lo"], two[0]["hi"], two[1]["lo"], two[1]["hi"]] == [0, 3, 3, 5]
    for r in two:                                        # replicated bookkeeping is identical everywhere
        for k in ("temp_of_slot", "timestep_T", "n_acc_T", "re_acc", "re_try"):
            assert r[k] == one[k], k
    # the only collective between the ranks is ONE all-gather per exchange round (it = 3, 6, ..., 30), plus one
    # per statistics row that does not fall on an exchange round (it = 5, 10, 20, 25) and one when run() returns
    for r in two + [one]:
        assert r["n_swap_rounds"] == 10 and r["n_collectives"] == 10 + 4 + 1
    assert two[0]["consistency"] == two[1]["consistency"] == one["consistency"]
    x2 = np.concatenate([np.array(r["x"]) for r in two]); n2 = np.concatenate([np.array(r["norm"]) for r in two])
    assert np.array_equal(x2, np.array(one["x"])) and np.array_equal(n2, np.array(one["norm"]))
    assert sum(one["re_try"]) > 0 and sorted(one["temp_of_slot"]) == list(range(5))
    assert one["temp_of_slot"] != list(range(5)) or sum(one["re_acc"]) == 0
    # output layout: per-temperature sample files with the reference's names, stats tables
    for out in ("w1", "w2"):
        d = str(tmp_path / out)
        files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(d, "samples", "*.pickle")))
        assert files == sorted("samples_replica%d_%d-%d.pickle" % (t, a, b) for t in range(1, 6)
                               for (a, b) in ((0, 10), (10, 20), (20, 30)))
        stats = np.loadtxt(os.path.join(d, "statistics", "mcmc_stats.txt"))
        assert stats.shape[1] == 1 + 2 * 5                      # step, then (p_acc, stepsize) per replica
        assert np.array_equal(stats[-1, 2::2], np.loadtxt(os.path.join(d, "statistics", "mcmc_stats.txt"))[-1, 2::2])
        assert os.path.exists(os.path.join(d, "statistics", "re_stats.txt"))
    # same samples whichever way the ladder was partitioned
    for f in glob.glob(os.path.join(str(tmp_path / "w1"), "samples", "*.pickle")):
        a = pickle.load(open(f, "rb")); b = pickle.load(open(f.replace("w1", "w2"), "rb"))
        assert len(a) == len(b) and len(a) in (5, 6)      # records thinned by dump_step = 2
        for sa, sb in zip(a, b):
            assert np.array_equal(sa.variables["structures"], sb.variables["structures"])
            assert sa.variables["norm"] == sb.variables["norm"]


def test_continuation_artifacts(tmp_path):
    """scripts/continue_simulation/setup_continue.py semantics: offset, last states, timesteps"""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from ensemble_hic_b200.rex import load_continuation
    out = str(tmp_path / "w1")
    one = _run(out, 1)[0]
    c = load_continuation(out, 5, 10)
    assert c["offset"] == 30
    last = [pickle.load(open(os.path.join(out, "samples", "samples_replica%d_20-30.pickle" % t), "rb"))[-1] for t in range(1, 6)]
    assert np.array_equal(c["structures"], np.array([s.variables["structures"] for s in last]))
    assert np.array_equal(c["norms"], np.array([s.variables["norm"] for s in last]))
    stats = np.loadtxt(os.path.join(out, "statistics", "mcmc_stats.txt"))
    assert np.array_equal(c["timesteps"], stats[-1, 2::2]) and np.allclose(c["timesteps"], one["timestep_T"], rtol=1e-5)
    for f in ("init_states.npy", "init_norms.npy", "timesteps.npy"):
        assert os.path.exists(os.path.join(out, "init_continue1", f))
    assert np.array_equal(np.load(os.path.join(out, "init_continue1", "init_states.npy")), c["structures"])
    import pytest
    with pytest.raises(IOError):
        load_continuation(str(tmp_path / "nothing"), 5, 10)
