"""N > 1 host logic on CPU: world_size 2, gloo backend, rendezvous on 127.0.0.1 (what bench.py --gpus N does around the
GPU solver: per-rank windows, barrier, max-over-ranks timing, whole-job counts; and the landmark-block partition of one
large window)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from limo_b200 import parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_world_size_2_gloo(tmp_path, oracle):
    out = tmp_path / "rank0.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(out.read_text())
    assert d["world"] == 2 and d["n_windows"] == 4
    # the step time is the slowest rank's, not rank 0's own
    assert d["ms_max"] == max(d["all_ms"]) and d["all_ms"][1] >= 5.0
    # whole-job aggregate equals the serial computation over both ranks' windows
    expect = 0.0
    for rank in range(2):
        for s in parallel.window_seeds(2, rank):
            expect += oracle.solve_window(synth.make_window(1, seed=s, n_kf=4, n_lm=60, n_obs=200)).solves[-1].final_cost
    assert np.isclose(d["cost_sum"], expect, rtol=1e-12)
    assert set(parallel.window_seeds(2, 0)).isdisjoint(parallel.window_seeds(2, 1))
    # the landmark shards tile the large window exactly and their per-keyframe contributions add up
    assert [d["n_obs_total"], d["n_lm_total"]] == list(d["big"]) and d["ok_slice"]
    assert d["counts_sum"] == d["counts_full"]
    assert d["packed_ok"]  # sums and the per-rank max slots through one sum all-reduce


def test_landmark_ranges_balance():
    win = synth.make_window(2, seed=3, n_kf=10, n_lm=500, n_obs=4000)
    for world in (1, 2, 3, 8):
        rng = parallel.landmark_ranges(win.lm_obs_ptr, world)
        assert rng[0][0] == 0 and rng[-1][1] == win.n_lm
        assert all(a[1] == b[0] for a, b in zip(rng, rng[1:]))
        obs = [int(win.lm_obs_ptr[j1] - win.lm_obs_ptr[j0]) for j0, j1 in rng]
        assert sum(obs) == win.n_obs and max(obs) - min(obs) <= 2 * np.diff(win.lm_obs_ptr).max()
        subs = [parallel.shard_window(win, r, world)[0] for r in range(world)]
        assert sum(s.n_lm for s in subs) == win.n_lm and all(s.n_kf == win.n_kf for s in subs)
