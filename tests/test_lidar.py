"""Lidar depth extraction (BASELINE config 4).  There is no reference code or test data for this step (it lives in the
un-vendored mono_lidar_depth repository), so the CPU restatement follows mono_lidar_fusion_parameters.yaml and is checked
for self-consistency on scenes with a known answer; the CUDA kernels are then held to that restatement exactly."""
import numpy as np
import pytest

from limo_b200 import geometry as g
from limo_b200 import synth


def _wall_scene(depth=10.0, tilt=0.0):
    """dense points on a plane in front of the camera (lidar frame == camera frame): z = depth + tilt * x"""
    xs, ys = np.meshgrid(np.linspace(-6, 6, 500), np.linspace(-2, 2, 160))
    z = depth + tilt * xs
    cloud = np.stack([xs.ravel(), ys.ravel(), z.ravel(), np.zeros(xs.size)], axis=1).astype(np.float32)
    return cloud, np.array([1.0, 0, 0, 0, 0, 0, 0]), np.array([synth.F, synth.CX, synth.CY])


def test_oracle_recovers_plane_depth(oracle):
    rng = np.random.default_rng(3)
    for tilt in (0.0, 0.3):
        cloud, T, K = _wall_scene(10.0, tilt)
        feats = np.stack([rng.uniform(300, 900, 200), rng.uniform(100, 280, 200)], axis=1).astype(np.float32)
        d = oracle.lidar_depth(cloud, T, K, feats)
        assert (d > 0).mean() > 0.9
        x_over_z = (feats[:, 0] - synth.CX) / synth.F
        expected = 10.0 / (1.0 - tilt * x_over_z)  # ray / plane intersection
        ok = d > 0
        assert np.allclose(d[ok], expected[ok], rtol=2e-4)


def test_oracle_gates(oracle):
    cloud, T, K = _wall_scene(10.0)
    feats = np.array([[600.0, 180.0]], dtype=np.float32)
    assert oracle.lidar_depth(cloud, T, K, feats)[0] > 0
    far = cloud.copy(); far[:, 2] += 200.0                       # beyond treshold_depth_max = 100 m
    assert oracle.lidar_depth(far, T, K, feats)[0] == -1.0
    behind = cloud.copy(); behind[:, 2] *= -1.0                  # do_use_cut_behind_camera
    assert oracle.lidar_depth(behind, T, K, feats)[0] == -1.0
    assert oracle.lidar_depth(cloud[:2], T, K, feats)[0] == -1.0  # fewer than 3 points
    # foreground wins: a second, farther wall behind a near one seen in the same rectangle
    near = cloud[(cloud[:, 0] > -0.05) & (cloud[:, 0] < 0.05)].copy(); near[:, 2] = 6.0
    mixed = np.concatenate([cloud, near])
    d = oracle.lidar_depth(mixed, T, K, np.array([[synth.CX, 180.0]], dtype=np.float32))[0]
    assert d == pytest.approx(6.0, rel=1e-3)


@pytest.mark.gpu
def test_lidar_kernel_matches_oracle(oracle):
    from limo_b200 import capi
    h = capi.Handle(0)
    cloud, T, K, feats = synth.make_lidar_scene()
    assert 100000 < len(cloud) <= 120000 and len(feats) == 2000
    for opt_mod in (None, dict(rect_width=24.0, rect_height=30.0)):
        og, oc = capi.lidar_default_options(), oracle.lidar_default_options()
        for k, v in (opt_mod or {}).items():
            setattr(og, k, v); setattr(oc, k, v)
        dg, ms = h.lidar_depth(cloud, T, K, feats, og)
        dc = oracle.lidar_depth(cloud, T, K, feats, oc)
        assert np.array_equal(dg > 0, dc > 0)
        assert np.array_equal(dg, dc)          # same single-precision operation order: bit-identical
        assert (dg > 0).sum() > (20 if opt_mod is None else 400)
    cloud, T, K = _wall_scene(12.0, 0.2)
    feats = np.stack([np.linspace(300, 900, 64), np.full(64, 180.0)], axis=1).astype(np.float32)
    dg, _ = h.lidar_depth(cloud, T, K, feats)
    assert np.array_equal(dg, oracle.lidar_depth(cloud, T, K, feats))
    h.close()
