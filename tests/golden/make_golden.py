#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the CPU oracle (oracle/, itself pinned to the reference's own tests by
tests/test_oracle_reference_tests.py).  The reference cannot be built or imported in this image (Ceres / Eigen / catkin
are absent), so these vectors are NOT reference outputs: they freeze the pinned oracle on small seeded inputs so that
  * the CPU suite notices any drift of the oracle (tests/test_golden.py::test_oracle_matches_golden), and
  * the GPU suite has committed vectors to hold the CUDA path to (tests/test_golden.py::test_cuda_matches_golden).

  python tests/golden/make_golden.py        # rewrites the .npz files next to this script
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from limo_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

# name -> (config, kwargs of synth.make_window); all small enough for one-thread oracle solves in well under a second
SOLVE_CASES = {
    "config1_seed7": (1, dict(seed=7)),
    "config2_slice": (2, dict(n_kf=10, n_lm=300, n_obs=2400, seed=5)),
    "config3_small": (3, dict(n_kf=8, n_lm=300, n_obs=1800, gp_frac=0.2, seed=41)),
}
EVAL_CASES = {
    "config1": (1, dict()),
    "config2_slice": (2, dict(n_kf=12, n_lm=400, n_obs=3000)),
}


def solve_record(win):
    r = orc.solve_window(win, num_threads=1)
    return dict(
        kf_pose=r.kf_pose.copy(), lm_pos=r.lm_pos[:win.n_lm].copy(), lm_rejected=r.lm_rejected[:win.n_lm].copy(),
        kf_plane=r.kf_plane.copy(),
        initial_cost=np.array([s.initial_cost for s in r.solves]), final_cost=np.array([s.final_cost for s in r.solves]),
        num_iterations=np.array([s.num_iterations for s in r.solves]),
        num_successful_steps=np.array([s.num_successful_steps for s in r.solves]),
        termination=np.array([s.termination for s in r.solves]), num_landmarks=np.array([s.num_landmarks for s in r.solves]))


def main():
    orc.lib()
    for name, (cfg, kw) in SOLVE_CASES.items():
        np.savez_compressed(os.path.join(HERE, "solve_%s.npz" % name), **solve_record(synth.make_window(cfg, **kw)))
    for name, (cfg, kw) in EVAL_CASES.items():
        r, jp, jl, cost, failed = orc.evaluate(synth.make_window(cfg, **kw))
        # the Jacobians of the small slice are kept in full, the larger one as every 7th observation (size)
        step = 1 if name == "config1" else 7
        np.savez_compressed(os.path.join(HERE, "eval_%s.npz" % name), r=r[::step], jp=jp[::step], jl=jl[::step],
                            cost=np.array([cost]), failed=np.array([failed]), step=np.array([step]))
    cloud, T, K, feats = synth.make_lidar_scene()
    d = orc.lidar_depth(cloud, T, K, feats)
    np.savez_compressed(os.path.join(HERE, "lidar_scene.npz"), depth=d)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-28s %7d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
