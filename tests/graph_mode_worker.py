"""Child of tests/test_graph_modes.py: solves a fixed set of windows under the KBA_GRAPH mode of its environment (the library
reads it once per process) and stores everything the caller gets back."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from limo_b200 import capi, synth  # noqa: E402


def main(out_path):
    h = capi.Handle(0)
    wins = [synth.make_window(1, seed=31), synth.make_window(3, seed=41, n_kf=8, n_lm=300, n_obs=1800, gp_frac=0.2),
            synth.make_window(2, n_kf=10, n_lm=300, n_obs=2500, seed=5)]
    out = {}
    single = [h.solve_window(w) for w in wins]
    batch = h.batch(wins)
    for rep in range(2):        # the second solve reuses the cached graph
        batch.upload()
        batch.solve()
        res = batch.download()
    batch.close()
    for tag, results, ws in (("single", single, wins), ("batch", res, wins)):
        for i, (r, w) in enumerate(zip(results, ws)):
            out["%s%d_pose" % (tag, i)] = r.kf_pose
            out["%s%d_plane" % (tag, i)] = r.kf_plane
            out["%s%d_lm" % (tag, i)] = r.lm_pos[:w.n_lm]
            out["%s%d_rej" % (tag, i)] = r.lm_rejected[:w.n_lm]
            out["%s%d_cost" % (tag, i)] = np.array([r.c.final_cost] + [s.final_cost for s in r.solves[:r.c.num_solves]])
            out["%s%d_iter" % (tag, i)] = np.array([s.num_iterations for s in r.solves[:r.c.num_solves]] + [r.c.status])
    cnt = h.counters()
    out["launches"] = np.array([cnt.launches_total])
    h.close()
    np.savez(out_path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
