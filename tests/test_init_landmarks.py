"""push() landmark initialisation for a whole window: the oracle's batch restatement against the
host mirror of the reference (limo_b200/adjuster.py, itself held to the reference's Triangulator / LandmarkCreator
tests), and the CUDA kernel against the oracle."""
import numpy as np
import pytest

from limo_b200 import geometry as g
from limo_b200 import synth


def _numpy_reference(win):
    """straight numpy restatement of cpp:332-355 (back-projection) and triangulator.hpp:51-75"""
    out = np.zeros((win.n_lm, 3)); flags = np.zeros(win.n_lm, dtype=np.uint8)
    Tc = [g.pose_to_iso(p) for p in win.cam_pose]
    Tk = [g.pose_to_iso(p) for p in win.kf_pose]
    for j in range(win.n_lm):
        o0, o1 = win.lm_obs_ptr[j], win.lm_obs_ptr[j + 1]
        obs = range(o0, o1)
        cam = lambda o: 0 if win.obs_cam is None else int(win.obs_cam[o])
        p = None
        for o in obs:
            if win.obs_d[o] >= 0:
                f, cx, cy = win.cam_intr[cam(o)]
                z = float(win.obs_d[o])
                pc = np.array([(float(win.obs_u[o]) - cx) * z / f, (float(win.obs_v[o]) - cy) * z / f, z, 1.0])
                p = (np.linalg.inv(Tc[cam(o)] @ Tk[win.obs_kf[o]]) @ pc)[:3]
                break
        if p is None and o1 - o0 >= 2:
            A = np.zeros((3, 3)); b = np.zeros(3)
            for o in obs:
                f, cx, cy = win.cam_intr[cam(o)]
                r = np.array([(float(win.obs_u[o]) - cx) / f, (float(win.obs_v[o]) - cy) / f, 1.0])
                r /= np.linalg.norm(r)
                T = np.linalg.inv(Tc[cam(o)] @ Tk[win.obs_kf[o]])
                rr = T[:3, :3] @ r
                M = np.eye(3) - np.outer(rr, rr)
                A += M; b += M @ T[:3, 3]
            p = np.linalg.solve(A, b)
        if p is None:
            continue
        front = all((Tc[cam(o)] @ Tk[win.obs_kf[o]] @ np.append(p, 1.0))[2] >= 0 for o in obs)
        out[j] = p; flags[j] = 1 | (int(front) << 1)
    return out, flags


def _windows():
    from tests import edge_windows as ew
    yield "config1", synth.make_window(1, seed=5)                       # mono, no depth: triangulation only
    yield "config2-slice", synth.make_window(2, n_kf=10, n_lm=300, n_obs=2400, seed=6)  # 40 % lidar depths
    yield "ragged", ew.ragged()                                          # empty rows and single observations


def test_oracle_init_landmarks_matches_numpy(oracle):
    for name, win in _windows():
        pos, flags = oracle.init_landmarks(win)
        ref, rflags = _numpy_reference(win)
        assert np.array_equal(flags, rflags), name
        created = (flags & 1) == 1
        assert np.allclose(pos[created], ref[created], rtol=1e-9, atol=1e-9), name
        if name == "ragged":
            assert (~created).sum() == 6   # 4 landmarks without observations, 2 with a single depth-less one ... or fewer
    # a landmark pushed behind a camera fails the cheirality bit
    win = synth.make_window(1, seed=5)
    u = win.obs_u.copy(); v = win.obs_v.copy()
    o0, o1 = win.lm_obs_ptr[0], win.lm_obs_ptr[1]
    u[o0:o1] = u[o0:o1][::-1]                 # scramble one track: its rays now intersect behind some camera or far off
    from tests.edge_windows import _rebuild
    bad = _rebuild(win, obs_u=u)
    pos, flags = oracle.init_landmarks(bad)
    ref, rflags = _numpy_reference(bad)
    assert np.array_equal(flags, rflags)


@pytest.mark.gpu
def test_cuda_init_landmarks_matches_oracle(oracle):
    from limo_b200 import capi
    h = capi.Handle(0)
    for name, win in list(_windows()) + [("config2", synth.make_window(2))]:
        pos, flags, ms = h.init_landmarks(win)
        ref, rflags = oracle.init_landmarks(win)
        assert np.array_equal(flags, rflags), name
        created = (flags & 1) == 1
        scale = np.maximum(1.0, np.abs(ref[created]))
        assert (np.abs(pos[created] - ref[created]) / scale).max() <= 1e-9, name
        assert np.array_equal(pos[~created], np.zeros_like(pos[~created])), name
    h.close()
