"""Finite-difference checks of the oracle's analytic local Jacobians.

The increments go through the oracle's own Plus operators (Ceres quaternion Plus, FixScaleVectorPlus), i.e. the check
is d r(Plus(x, delta)) / d delta at delta = 0 -- exactly what AutoDiffCostFunction x LocalParameterization yields in
the reference.
"""
import numpy as np
import pytest

from limo_b200 import geometry as g

H = 1e-6


def rand_pose(rng, scale=1.0):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    R = g.angle_axis(rng.uniform(-1, 1), ax)
    return np.concatenate([g.rot_to_quat(R), rng.normal(size=3) * scale])


def fd(fun, n, plus):
    """central differences of fun(plus(delta)) for delta in R^n"""
    cols = []
    for i in range(n):
        d = np.zeros(n); d[i] = H
        cols.append((fun(plus(d)) - fun(plus(-d))) / (2 * H))
    return np.stack(cols, axis=-1)


@pytest.mark.parametrize("seed", range(5))
def test_reprojection_and_depth_jacobians(oracle, seed):
    rng = np.random.default_rng(seed)
    pose, cam = rand_pose(rng), rand_pose(rng, 0.3)
    intr = [718.856, 607.1928, 185.2157]
    # a point in front of the camera
    pc = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(4, 40)])
    p = g.apply(g.iso_inv(g.pose_to_iso(cam) @ g.pose_to_iso(pose)), pc)
    ok, r0, jp, jl = oracle.reprojection(pose, cam, intr, p, 600.0, 180.0)
    assert ok
    f_pose = lambda q: oracle.reprojection(q, cam, intr, p, 600.0, 180.0, jac=False)[1]
    f_pt = lambda x: oracle.reprojection(pose, cam, intr, x, 600.0, 180.0, jac=False)[1]
    Jp_fd = fd(f_pose, 6, lambda d: oracle.pose_plus(pose, d))
    Jl_fd = fd(f_pt, 3, lambda d: p + d)
    assert np.allclose(jp, Jp_fd, rtol=1e-6, atol=1e-5 * np.abs(jp).max())
    assert np.allclose(jl, Jl_fd, rtol=1e-6, atol=1e-5 * np.abs(jl).max())
    r, jpd, jld = oracle.depth(pose, cam, p, 7.5)
    assert r[0] == pytest.approx(pc[2] - 7.5, abs=1e-9)
    Jp_fd = fd(lambda q: oracle.depth(q, cam, p, 7.5)[0], 6, lambda d: oracle.pose_plus(pose, d))
    Jl_fd = fd(lambda x: oracle.depth(pose, cam, x, 7.5)[0], 3, lambda d: p + d)
    assert np.allclose(jpd, Jp_fd, rtol=1e-6, atol=1e-6)
    assert np.allclose(jld, Jl_fd, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("seed", range(3))
def test_ground_plane_and_regulariser_jacobians(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    pose0, pose1 = rand_pose(rng, 3.0), rand_pose(rng, 3.0)
    n = rng.normal(size=3); n /= np.linalg.norm(n)
    p = rng.normal(size=3) * 5
    # ground plane height
    r, jp, jd, jdist, jl = oracle.gp_height(pose0, n, 0.31, p)
    assert np.allclose(jp, fd(lambda q: oracle.gp_height(q, n, 0.31, p)[0], 6, lambda d: oracle.pose_plus(pose0, d)), atol=1e-6)
    assert np.allclose(jd, fd(lambda m: oracle.gp_height(pose0, m, 0.31, p)[0], 3, lambda d: oracle.dir_plus(n, d)), atol=1e-6)
    assert np.allclose(jl, fd(lambda x: oracle.gp_height(pose0, n, 0.31, x)[0], 3, lambda d: p + d), atol=1e-6)
    assert jdist[0, 0] == 1.0
    # ground plane motion
    r, j0, j1, jdir = oracle.gp_motion(pose0, pose1, n)
    assert np.allclose(j0, fd(lambda q: oracle.gp_motion(q, pose1, n)[0], 6, lambda d: oracle.pose_plus(pose0, d)), atol=1e-6)
    assert np.allclose(j1, fd(lambda q: oracle.gp_motion(pose0, q, n)[0], 6, lambda d: oracle.pose_plus(pose1, d)), atol=1e-6)
    assert np.allclose(jdir, fd(lambda m: oracle.gp_motion(pose0, pose1, m)[0], 3, lambda d: oracle.dir_plus(n, d)), atol=1e-6)
    # scale regulariser: value = |(T1 T0^-1).t| - s0
    r, j1, j0 = oracle.scale_reg(pose1, pose0, 0.7)
    T = g.pose_to_iso(pose1) @ g.iso_inv(g.pose_to_iso(pose0))
    assert r[0] == pytest.approx(np.linalg.norm(T[:3, 3]) - 0.7, abs=1e-12)
    assert np.allclose(j1, fd(lambda q: oracle.scale_reg(q, pose0, 0.7)[0], 6, lambda d: oracle.pose_plus(pose1, d)), atol=1e-6)
    assert np.allclose(j0, fd(lambda q: oracle.scale_reg(pose1, q, 0.7)[0], 6, lambda d: oracle.pose_plus(pose0, d)), atol=1e-6)
    # speed prior
    Tob = rand_pose(rng, 2.0)
    vb = rng.normal(size=3)
    r, jp = oracle.speed_reg(pose0, Tob, 0.1, vb)
    Tc = g.pose_to_iso(pose0) @ g.pose_to_iso(Tob)
    assert np.allclose(r, Tc[:3, 3] / 0.1 - vb, atol=1e-12)
    assert np.allclose(jp, fd(lambda q: oracle.speed_reg(q, Tob, 0.1, vb)[0], 6, lambda d: oracle.pose_plus(pose0, d)), atol=1e-5)


def test_pose_plus_matches_ceres_quaternion_plus(oracle):
    """q+ = q_delta (x) q with q_delta = [cos|d|, sin|d|/|d| d]; translation additive."""
    rng = np.random.default_rng(7)
    pose = rand_pose(rng)
    d = rng.normal(size=6) * 0.1
    out = oracle.pose_plus(pose, d)
    nd = np.linalg.norm(d[:3])
    qd = np.concatenate([[np.cos(nd)], np.sin(nd) / nd * d[:3]])
    R = g.quat_to_rot(qd) @ g.quat_to_rot(pose[:4])
    assert np.allclose(g.quat_to_rot(out[:4]), R, atol=1e-14)
    assert np.allclose(out[4:], pose[4:] + d[3:])
    assert np.array_equal(oracle.pose_plus(pose, np.zeros(6)), pose)
    n = np.array([0.1, -0.2, 0.97]); n /= np.linalg.norm(n)
    out = oracle.dir_plus(n, [0.01, 0.02, -0.03])
    assert np.linalg.norm(out) == pytest.approx(1.0, abs=1e-15)
