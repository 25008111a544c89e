#!/usr/bin/env python
"""adjustPoseOnly() (reference cpp:820-888: one free pose against constant landmarks + the speed prior, called on every
frame) as a batch through the same kernels: throughput of B motion-only problems per kba_batch_solve, and the CPU oracle
next to it.  SURVEY 8(f) row 1."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from limo_b200 import capi, synth, geometry as g
from limo_b200.capi_types import Window
from oracle import oracle as orc


def motion_only_window(seed):
    win, truth = synth.make_window(2, n_kf=12, n_lm=1200, n_obs=12000, seed=seed, return_truth=True)
    k = win.n_kf - 1
    sel = win.obs_kf == k
    lm_of_obs = np.repeat(np.arange(win.n_lm), np.diff(win.lm_obs_ptr))
    lms = lm_of_obs[sel]
    Tb, Tb2 = g.pose_to_iso(truth["kf_pose"][k - 1]), g.pose_to_iso(truth["kf_pose"][k - 2])
    return Window(kf_pose=win.kf_pose[k:k + 1], kf_fixed=[0], cam_intr=win.cam_intr, cam_pose=win.cam_pose,
                  lm_pos=truth["lm_pos"][lms], lm_weight=np.ones(len(lms)), lm_obs_ptr=np.arange(len(lms) + 1),
                  obs_kf=np.zeros(len(lms), dtype=np.int32), obs_u=win.obs_u[sel], obs_v=win.obs_v[sel],
                  obs_d=win.obs_d[sel], landmarks_fixed=True, speed_kf=0, speed_weight=0.7, speed_dt=0.1,
                  speed_v_before=(Tb @ g.iso_inv(Tb2))[:3, 3] / 0.1, speed_T_origin_before=g.iso_to_pose(g.iso_inv(Tb)))


def main():
    base = [motion_only_window(100 + i) for i in range(8)]
    h = capi.Handle(0)
    opt = capi.default_options()
    out = {"workload": "adjustPoseOnly: 1 free pose, %d observations, speed prior" % base[0].n_obs}
    for B in (1, 64, 1024):
        batch = h.batch([base[i % 8] for i in range(B)])
        for _ in range(2):
            batch.solve(opt)
        t = time.perf_counter()
        n = 5
        for _ in range(n):
            batch.solve(opt)
        dt = (time.perf_counter() - t) / n
        res = batch.download()
        out["batch_%d" % B] = {"ms_per_batch": 1e3 * dt, "problems_per_s": B / dt,
                               "iterations": int(sum(s.num_iterations for s in res[0].solves))}
        batch.close()
    orc.lib()
    t = time.perf_counter()
    for i in range(8):
        rc = orc.solve_window(base[i], opt)
    out["cpu_oracle_problems_per_s_1_thread"] = 8 / (time.perf_counter() - t)
    rg = h.solve_window(base[0], opt)
    out["max_dt_vs_oracle"] = float(np.linalg.norm(rg.kf_pose[:, 4:] - orc.solve_window(base[0], opt).kf_pose[:, 4:], axis=1).max())
    print(json.dumps(out))
    h.close()


if __name__ == "__main__":
    main()
