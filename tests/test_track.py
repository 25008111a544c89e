"""The persistent, device-resident window (kba_track_*) against the rebuild-per-call path.

A 40-keyframe synthetic drive is replayed the way the reference's node drives the adjuster: push a keyframe, solve the
sliding window, the solve updates poses and landmarks in place, push the next keyframe ...  Path A keeps everything on the
device (kba_track_push_keyframe once per keyframe, kba_track_solve with the lists of active keyframe / selected landmark slots);
path B builds the whole kba_window on the host for every solve (what solve() did in round 1 and what the reference does with
its ceres::Problem, bundle_adjuster_keyframes.cpp:635-637).  The two must agree bit for bit, and path A must send a small
fraction of path B's bytes per solve."""
import numpy as np
import pytest

from limo_b200 import synth
from limo_b200.capi_types import Window

pytestmark = pytest.mark.gpu

W = 10          # sliding window length
N_KF = 40


def _drive():
    win = synth.make_window(2, n_kf=N_KF, n_lm=2500, n_obs=30000, seed=123)
    lm_of_obs = np.repeat(np.arange(win.n_lm), np.diff(win.lm_obs_ptr))
    per_kf = []
    for k in range(N_KF):
        sel = np.nonzero(win.obs_kf == k)[0]          # landmark-major order = ascending landmark id inside a keyframe
        per_kf.append((lm_of_obs[sel].astype(np.int32), win.obs_u[sel], win.obs_v[sel], win.obs_d[sel]))
    return win, per_kf


def _window_lists(per_kf, first, last):
    """selected landmarks (>= 2 observations inside the window) and the host-built CSR of path B"""
    count = {}
    for k in range(first, last + 1):
        for j in per_kf[k][0]:
            count[j] = count.get(j, 0) + 1
    lm_sel = np.array(sorted(j for j, c in count.items() if c >= 2), dtype=np.int32)
    index = {j: i for i, j in enumerate(lm_sel)}
    rows = [[] for _ in lm_sel]
    for k in range(first, last + 1):
        lm, u, v, d = per_kf[k]
        for a in range(len(lm)):
            i = index.get(int(lm[a]))
            if i is not None:
                rows[i].append((k - first, u[a], v[a], d[a]))
    ptr = np.zeros(len(lm_sel) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    flat = [x for r in rows for x in r]
    return lm_sel, ptr, np.array([x[0] for x in flat], dtype=np.int32), np.array([x[1] for x in flat], dtype=np.float32), \
        np.array([x[2] for x in flat], dtype=np.float32), np.array([x[3] for x in flat], dtype=np.float32)


def _scale(poses, n_depth):
    from limo_b200 import geometry as g
    T10 = g.pose_to_iso(poses[1]) @ g.iso_inv(g.pose_to_iso(poses[0]))
    return dict(scale_kf0=0, scale_kf1=1, scale_weight=1000.0 / max(n_depth, 1), scale_value=float(np.linalg.norm(T10[:3, 3])))


def test_thirty_pushes_and_solves_equal_the_rebuild_path():
    from limo_b200 import capi
    win, per_kf = _drive()
    h = capi.Handle(0)
    track = capi.Track(h, win.cam_intr, win.cam_pose, max_keyframes=N_KF, max_landmarks=win.n_lm, max_measurements=win.n_obs,
                       win_keyframes=W, win_landmarks=win.n_lm, win_observations=win.n_obs)
    # state both paths evolve in place (the rebuild path's copy lives on the host)
    poses_b, lms_b = win.kf_pose.copy(), win.lm_pos.copy()
    track.set_landmarks(np.arange(win.n_lm, dtype=np.int32), pos=win.lm_pos, weight=win.lm_weight)
    for k in range(W):
        track.push_keyframe(k, win.kf_pose[k], *per_kf[k])
    rebuild_h2d, n_solves = [], 0
    for last in range(W - 1, W - 1 + 30):
        first = last - W + 1
        if last >= W:
            track.push_keyframe(last, win.kf_pose[last], *per_kf[last])
        lm_sel, ptr, okf, ou, ov, od = _window_lists(per_kf, first, last)
        fixed = np.zeros(W, dtype=np.uint8); fixed[0] = 1
        sc = _scale(poses_b[first:last + 1], int((od > 0).sum()))
        # ---- path B: the whole window from host arrays
        wb = Window(poses_b[first:last + 1], fixed, win.cam_intr, win.cam_pose, lms_b[lm_sel], win.lm_weight[lm_sel], ptr, okf, ou, ov, od, **sc)
        batch = h.batch([wb])
        batch.solve(capi.default_options())
        rb = batch.download()[0]
        rebuild_h2d.append(batch.transfer_bytes()[0])
        batch.close()
        # ---- path A: lists only
        ra = track.solve(np.arange(first, last + 1), fixed, lm_sel, **sc)
        assert ra.c.status == 0 and rb.c.status == 0
        assert [s.num_iterations for s in ra.solves] == [s.num_iterations for s in rb.solves], "solve %d" % n_solves
        assert np.array_equal(ra.kf_pose, rb.kf_pose), "solve %d" % n_solves
        assert np.array_equal(ra.lm_pos[:len(lm_sel)], rb.lm_pos[:len(lm_sel)])
        assert np.array_equal(ra.lm_rejected[:len(lm_sel)], rb.lm_rejected[:len(lm_sel)])
        assert ra.c.final_cost == rb.c.final_cost
        poses_b[first:last + 1] = rb.kf_pose          # in-place semantics of solve() (cpp:554-557, 592-593)
        lms_b[lm_sel] = rb.lm_pos[:len(lm_sel)]
        n_solves += 1
    h2d_solve, d2h_solve, h2d_push = track.transfer_bytes()
    assert n_solves == 30
    # bytes: a solve of the persistent window sends < 10 % of the rebuild path's upload; the pushes add the measurements once
    assert h2d_solve < 0.1 * np.mean(rebuild_h2d), (h2d_solve, np.mean(rebuild_h2d))
    # ... and everything together (every keyframe's measurements and every landmark's initial value once + 30 selections) is a
    # fraction of what the rebuild path uploads for the same 30 solves
    assert h2d_push + 30 * h2d_solve < 0.25 * np.sum(rebuild_h2d), (h2d_push, h2d_solve, np.sum(rebuild_h2d))
    print("persistent window: %d B per solve (rebuild path %.0f B), %d B for all pushes" % (h2d_solve, np.mean(rebuild_h2d), h2d_push))
    track.close()
    h.close()


def test_track_capacity_errors():
    from limo_b200 import capi
    win, per_kf = _drive()
    h = capi.Handle(0)
    with pytest.raises(capi.KbaError, match="error 4"):   # 40 keyframes do not fit the fused path
        capi.Track(h, win.cam_intr, win.cam_pose, 64, 100, 1000, win_keyframes=40, win_landmarks=100, win_observations=1000)
    track = capi.Track(h, win.cam_intr, win.cam_pose, 8, win.n_lm, 20000, win_keyframes=4, win_landmarks=win.n_lm, win_observations=20000)
    track.push_keyframe(0, win.kf_pose[0], *per_kf[0])
    with pytest.raises(capi.KbaError, match="slot in use"):
        track.push_keyframe(0, win.kf_pose[0], *per_kf[0])
    with pytest.raises(capi.KbaError, match="error 3"):   # NotEnoughKeyframes
        track.solve([0], [1], [0, 1])
    with pytest.raises(capi.KbaError, match="not pushed"):
        track.solve([0, 1, 2], [1, 0, 0], [0, 1])
    track.close()
    h.close()
