"""The three ways kba_batch_solve can issue the passes of a solve (KBA_GRAPH: 2 = one CUDA graph with a conditional WHILE node,
1 = flat graph of four passes, 0 = kernel by kernel) run the same kernels with the same arguments: results must be bit-identical."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_graph_modes_are_bit_identical(tmp_path):
    outs = {}
    for mode in ("0", "1", "2"):
        path = str(tmp_path / ("mode%s.npz" % mode))
        env = dict(os.environ, KBA_GRAPH=mode, KBA_GRAPH_VERBOSE="1")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "graph_mode_worker.py"), path], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "not available" not in r.stderr, r.stderr[-2000:]          # no silent fallback to the stream path
        if mode != "0":
            assert "solve graph built (mode %s" % mode in r.stderr, r.stderr[-2000:]
        outs[mode] = dict(np.load(path))
    ref = outs["0"]
    for mode in ("1", "2"):
        for k, v in ref.items():
            if k == "launches":
                continue
            assert np.array_equal(v, outs[mode][k], equal_nan=True), (mode, k)
    # every mode reports the kernels it ran (the graph modes count them from the device's pass counter)
    assert all(int(o["launches"][0]) > 100 for o in outs.values())
