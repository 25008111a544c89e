"""Pins the CPU oracle against every known-answer / geometric test the reference holds for the hot path.

Each test names the reference gtest it restates (keyframe_bundle_adjustment/test/keyframe_bundle_adjustment.cpp,
robust_optimization/test/robust_optimization.cpp).  The host logic under test (push / landmark creation / problem
assembly) is limo_b200.adjuster with the oracle injected as backend; no GPU is involved.
"""
import numpy as np
import pytest

from limo_b200 import geometry as g
from limo_b200.adjuster import Camera, Keyframe, NotEnoughKeyframesException, BundleAdjusterKeyframes
from tests import ref_scenes as rs

ID7 = [1.0, 0, 0, 0, 0, 0, 0]


# ---- CostFunctor.* known answers -------------------------------------------------------------------------
def test_costfunctor_get_error_point_ray(oracle):
    """CostFunctor.get_error_point_ray (:118-175)"""
    ok, res, _, _ = oracle.reprojection(ID7, ID7, [600.0, 200.0, 100.0], [1.0, 1.0, 10.0], 260.0, 160.0)
    assert ok and abs(res[0]) < 1e-5 and abs(res[1]) < 1e-5
    # second half: rotated pose; EulerAnglesToQuaternion (ceres, local_parameterizations.hpp) -> use an arbitrary unit quaternion
    R = g.angle_axis(0.2, [0, 0, 1.0]) @ g.angle_axis(0.05, [0, 1.0, 0]) @ g.angle_axis(0.1, [1.0, 0, 0])
    pose = np.concatenate([g.rot_to_quat(R), [0.01, -0.01, 0.01]])
    p = g.apply(g.pose_to_iso(pose), [1.0, 1.0, 10.0])
    proj = np.array([[600.0, 0, 200.0], [0, 600.0, 100.0], [0, 0, 1]]) @ (p / p[2])
    ok, res, _, _ = oracle.reprojection(pose, ID7, [600.0, 200.0, 100.0], [1.0, 1.0, 10.0], proj[0], proj[1])
    assert ok and abs(res[0]) < 1e-2 and abs(res[1]) < 1e-2


def test_costfunctor_groundplane_height(oracle):
    """CostFunctor.GroundPlaneHeightRegularization (:1346-1358): exact 0.5"""
    res, *_ = oracle.gp_height(ID7, [0, 0, 1.0], 1.0, [2.0, 3.0, -0.5])
    assert res[0] == 0.5


def test_costfunctor_groundplane_motion(oracle):
    """CostFunctor.GroundPlaneMotionRegularization (:1360-1371): exact -0.5/sqrt(1.25)"""
    res, *_ = oracle.gp_motion([1.0, 0, 0, 0, 2.0, 0, 0], [1.0, 0, 0, 0, 2.0, 1.0, 0.5], [0, 0, 1.0])
    assert res[0] == -0.5 / np.sqrt(0.0 * 0.0 + 1.0 * 1.0 + 0.5 * 0.5)


def test_reprojection_fails_close_to_camera_plane(oracle):
    """ReprojectionErrorWithQuaternions::project returns false for |z| < 0.01 (cost_functors_ceres.hpp:78-83)"""
    ok, *_ = oracle.reprojection(ID7, ID7, [600.0, 200.0, 100.0], [1.0, 1.0, 0.005], 0.0, 0.0)
    assert not ok
    ok, *_ = oracle.reprojection(ID7, ID7, [600.0, 200.0, 100.0], [1.0, 1.0, -0.02], 0.0, 0.0)
    assert ok


# ---- Triangulator ----------------------------------------------------------------------------------------
def test_triangulator_process(oracle):
    """Triangulator.process (:51-74)"""
    p = np.array([1.0, 1.0, 3.0])
    t = g.iso(t=[1.0, -1.0, 0.0])
    v1 = p / np.linalg.norm(p)
    v2 = g.apply(g.iso_inv(t), p); v2 /= np.linalg.norm(v2)
    out = oracle.triangulate_rays([np.eye(3), t[:3, :3]], [np.zeros(3), t[:3, 3]], [v1, v2])
    assert np.linalg.norm(out - p) < 1e-5


def test_triangulator_process2(oracle):
    """Triangulator.process2 (:76-117)"""
    p_gt = np.array([0.5, -1.0, 3.0])
    t0 = g.iso(t=[1.0, -0.1, 0.5])
    t1 = g.translate(t0, [0.5, -0.05, 0.25])
    v0 = g.apply(g.iso_inv(t0), p_gt); v0 /= np.linalg.norm(v0)
    v1 = g.apply(g.iso_inv(t1), p_gt); v1 /= np.linalg.norm(v1)
    out = oracle.triangulate_rays([t0[:3, :3], t1[:3, :3]], [t0[:3, 3], t1[:3, 3]], [v0, v1])
    assert np.linalg.norm(out - p_gt) < 1e-5
    from limo_b200.adjuster import triangulate_rays
    assert np.linalg.norm(triangulate_rays([(t0, v0), (t1, v1)]) - p_gt) < 1e-5


# ---- Trimmers --------------------------------------------------------------------------------------------
def _trimmer_data(seed):
    """makeData (robust_optimization/test/robust_optimization.cpp:46-73), deterministically seeded."""
    rng = np.random.default_rng(seed)
    out = np.maximum(np.minimum(rng.normal(5.0, 1.0, 10), 3.6), 100.0)       # min(val, 3.6) then max(.,100) -> 100
    inl = np.maximum(np.minimum(rng.normal(0.0, 1.0, 100), 0.0), 3.4)        # -> 3.4
    return np.concatenate([out, inl])


def test_trimmer_quantile(oracle):
    """Trimmers.TrimmerQuantile (:99-109): 110 samples, q = 0.9 -> exactly 11 rejected"""
    n, rej = oracle.trimmer_quantile(_trimmer_data(1), 0.9)
    assert n == 11 and rej.sum() == 11
    v = np.random.default_rng(2).normal(size=110)
    n, rej = oracle.trimmer_quantile(v, 0.9)
    assert n == 11 and set(np.argsort(v)[-11:]) == set(np.nonzero(rej)[0])
    n, rej = oracle.trimmer_quantile(v, 1.0)  # quantile 1.0 rejects nothing (gp group, cpp:758)
    assert n == 0 and not rej.any()


def test_trimmer_fix(oracle):
    """Trimmers.TrimmerFix (robust_optimization/test/robust_optimization.cpp:89-97): threshold 3.5 on makeData() -> the 10 outliers"""
    n, rej = oracle.trimmer_fix(_trimmer_data(1), 3.5)
    assert n == 10 and rej[:10].all() and not rej[10:].any()
    n, rej = oracle.trimmer_fix([3.5, 3.5000001, -1.0], 3.5)       # strictly greater (trimmer_fix.hpp:41)
    assert n == 1 and list(rej) == [False, True, False]


def test_solve_trimmed_scalar_problem(oracle):
    """Solve.trimmed (robust_optimization.cpp:134-156): one parameter x = 2, 90 residuals 3x and 10 constant residuals 10, ALL in
    residual group 1, TrimmerQuantile 0.9, iterations {0, 2}.  Restated with the solveTrimmed outer loop (robust_solving.cpp:140-248)
    around a scalar Levenberg-Marquardt step with ceres' Jacobi-scaled damping: one group is fewer than
    minimum_number_residual_groups, so nothing is trimmed (robust_solving.cpp:19-21), the constant residuals have no gradient,
    and the final solve drives x to 0 within the reference's 1e-3."""
    x = 2.0
    groups = {1: list(range(100))}                                   # residual index -> all in group 1
    res = lambda x_: np.array([3.0 * x_] * 90 + [10.0] * 10)
    jac = np.array([3.0] * 90 + [0.0] * 10)

    def lm(x_, iters, radius=1e4):
        for _ in range(iters):
            r = res(x_)
            g, h = jac @ r, jac @ jac
            s2 = (1.0 / (1.0 + np.sqrt(h))) ** 2                     # Jacobi scaling as in the oracle
            lam = min(max(h * s2, 1e-6), 1e32) / (radius * s2)
            cand = x_ - g / (h + lam)
            if (res(cand) ** 2).sum() < (r ** 2).sum():
                x_, radius = cand, radius * 3.0
            else:
                radius /= 2.0
        return x_

    for iters in (0, 2):                                             # trimming rounds
        c0 = (res(x) ** 2).sum()
        x1 = lm(x, iters)
        if c0 - (res(x1) ** 2).sum() <= 0.0:                         # retry rule, robust_solving.cpp:172-181
            x1 = lm(x, 3 * iters)
        x = x1
        vals = np.array([np.abs(res(x)[idx]).max() for idx in groups.values()])
        if len(vals) >= 30:                                          # minimum_number_residual_groups
            n, rej = oracle.trimmer_quantile(vals, 0.9)
            groups = {k: v for (k, v), r_ in zip(groups.items(), rej) if not r_}
    assert len(groups) == 1                                          # nothing was removed
    x = lm(x, 100)                                                   # final solve
    assert abs(x) < 1e-3


# ---- LandmarkCreator.CreateWithDepth ---------------------------------------------------------------------
def test_landmark_creation_with_depth(oracle):
    """LandmarkCreator.CreateWithDepth (:1149-1210): depth back-projection recovers the landmark to 0.01 m"""
    b, poses_gt, noisy, lms, *_ = rs.build_adjuster(oracle.OracleBackend(), (0, 0, 0), (0, 0, 0, 0), [np.eye(4)],
                                                    with_depth=True)
    assert len(b.landmarks_) == len(lms)
    for i, lm in enumerate(lms):
        assert b.landmarks_[i].has_measured_depth
        assert np.linalg.norm(b.landmarks_[i].pos - lm) < 1e-2


# ---- KeyFrameBundleAdjustment.solve ----------------------------------------------------------------------
SOLVE_CASES = [
    ((0.0, 0.0), (0.0, 0.0, 0.0, 0.0), 1e-3),
    ((0.0, 0.0), (5.0 * np.pi / 180.0, 0.2, 0.1, 0.1), 1e-3),
    ((1.5, 1.5), (5.0 * np.pi / 180.0, 0.2, 0.1, 0.1), 1e-2),
]


def _check_solve(oracle, extr, noise_lms, noise_poses, thres, with_depth):
    b, poses_gt, noisy, lms, *_ = rs.build_adjuster(oracle.OracleBackend(), noise_lms, noise_poses, extr, with_depth)
    for i, lm in enumerate(lms):  # :472-476 triangulated / back-projected landmarks close to ground truth
        assert len(b.landmarks_) == len(lms)
        assert np.linalg.norm(b.landmarks_[i].pos - lm) < 1e-1
    summary = b.solve()
    assert "Merged summaries" in summary
    for ts in sorted(b.keyframes_):  # :606-608
        assert g.is_approx(b.keyframes_[ts].getEigenPose(), poses_gt[ts], thres), (ts, b.keyframes_[ts].pose_)
    return b


@pytest.mark.parametrize("noise_lms,noise_poses,thres", SOLVE_CASES)
def test_solve_mono(oracle, noise_lms, noise_poses, thres):
    """KeyFrameBundleAdjustment.solve, mono rig (:816-830)"""
    b = _check_solve(oracle, [rs.mono_extrinsics()], noise_lms, noise_poses, thres, False)
    # with 5 landmarks the trimming loop is never entered (cpp:741): exactly one inner solve
    assert b.last_result.c.num_solves == 1
    # fewer than 10 depth residuals, no gp: fixed-scale regulariser with weight 1000 (cpp:712-716)
    assert b.last_window.scale_weight == 1000.0


@pytest.mark.parametrize("noise_lms,noise_poses,thres", SOLVE_CASES)
def test_solve_multicam(oracle, noise_lms, noise_poses, thres):
    """KeyFrameBundleAdjustment.solve, two-camera rig (:832-857)"""
    _check_solve(oracle, rs.stereo_extrinsics(), noise_lms, noise_poses, thres, False)


@pytest.mark.parametrize("noise_lms,noise_poses,thres", SOLVE_CASES)
def test_solve_depth_mono(oracle, noise_lms, noise_poses, thres):
    """KeyFrameBundleAdjustment.solve_depth, mono (:1090-1110)"""
    b = _check_solve(oracle, [np.eye(4)], noise_lms + (0.0,), noise_poses, thres, True)
    # 4 keyframes x 5 landmarks = 20 depth residuals > 10: scale weight 1000 / 20 (cpp:704-711)
    assert b.last_window.scale_weight == 1000.0 / 20.0


@pytest.mark.parametrize("noise_lms,noise_poses,thres", SOLVE_CASES)
def test_solve_depth_multicam(oracle, noise_lms, noise_poses, thres):
    """KeyFrameBundleAdjustment.solve_depth, two cameras (:1112-1144)"""
    _check_solve(oracle, rs.stereo_extrinsics(), noise_lms + (0.0,), noise_poses, thres, True)


def test_adjust_motion_only(oracle):
    """BundleAdjusterKeyframes.adjustMotionOnly (:1340-1344): tolerance 0.5"""
    b, poses_gt, noisy, lms, ts, cameras, l2c = rs.build_adjuster(
        oracle.OracleBackend(), (0, 0, 0), (0, 0, 0, 0), [np.eye(4)], with_depth=True, motion_only=True)
    kf = Keyframe(4, ts, b.keyframes_[0].cameras_[0], noisy[4])
    b.landmark_selector_.select(b.getActiveLandmarkConstPtrs(), b.getActiveKeyframeConstPtrs())
    b.adjustPoseOnly(kf)
    assert g.is_approx(kf.getEigenPose(), poses_gt[4], 0.5)
    # The reference's tolerance is loose because stamps 2,3,4 are nanoseconds: dt = 1e-9 s makes the constant-velocity
    # prior (cpp:835-853) dominate.  It was attached here:
    assert b.last_window.speed_weight == 1.0 and b.last_window.speed_dt == pytest.approx(1e-9)
    # Without the prior (rotation between the two newest keyframes above 0.03 rad) the pose is recovered exactly.
    b.keyframes_[3].pose_ = g.iso_to_pose(g.rotate(b.keyframes_[3].getEigenPose(), g.angle_axis(0.05, [0, 0, 1.0])))
    kf2 = Keyframe(4, ts, b.keyframes_[0].cameras_[0], noisy[4])
    b.adjustPoseOnly(kf2)
    assert b.last_window.speed_weight == 0.0
    assert g.is_approx(kf2.getEigenPose(), poses_gt[4], 1e-3)


def test_solve_needs_three_keyframes(oracle):
    """solve() throws NotEnoughKeyframesException below 3 keyframes (cpp:630-632)"""
    b = BundleAdjusterKeyframes(backend=oracle.OracleBackend())
    with pytest.raises(NotEnoughKeyframesException):
        b.solve()


def test_deactivate_keyframes(oracle):
    """BundleAdjusterKeyframes.deactivateKeyframes (:744-805), window bookkeeping part"""
    b, *_ = rs.build_adjuster(oracle.OracleBackend(), (0, 0), (0, 0, 0, 0), [rs.mono_extrinsics()])
    assert len(b.active_keyframe_ids_) == 5
    b.deactivateKeyframes(3, 2, 3)
    assert len(b.active_keyframe_ids_) == 3
    ordered = sorted(b.active_keyframe_ids_)
    assert b.keyframes_[ordered[0]].fixation_status_ == Keyframe.FIX_POSE
    assert b.keyframes_[ordered[1]].fixation_status_ == Keyframe.FIX_SCALE


def test_landmark_selector_cheirality():
    """LandmarkSelector.base, cheirality part (reference test :649-707): five landmarks seen from the five test poses
    through an identity-extrinsic camera (f = 600, pp = (300, 200)); the library-default rejection scheme keeps exactly
    three of them (landmark 1 at z = -20 lies behind the image plane, landmark 3 at z = 1.5 falls behind the later
    poses).  Host logic of the drop-in (A17): the Python mirror here, the C++ facade in tests/cpp/test_facade.cpp."""
    from limo_b200.adjuster import Camera, Keyframe, Landmark, LandmarkSelector
    lms = [np.array(p) for p in ((0.5, 3., 5.5), (0., 1., -20.), (1., -5., 4.), (2.0, 1., 1.5), (-2.0, -1., 10.))]
    poses = rs.get_poses(0.0, (0.0, 0.0, 0.0))
    cam = Camera(600.0, (300.0, 200.0), np.eye(4))
    ts = rs.make_tracklets(poses, lms, {0: cam}, (0.0, 0.0, 0.0), {i: [0] for i in range(len(lms))}, with_depth=True)
    landmarks = {i: Landmark(p) for i, p in enumerate(lms)}
    kfs = {i: Keyframe(i, ts, cam, pose) for i, pose in enumerate(poses)}
    selected = LandmarkSelector().select(landmarks, kfs)
    assert len(selected) == 3 and 1 not in selected
    # the batch restatement used by the device kernel agrees on which landmarks are in front of every camera
    behind = {i for i, p in enumerate(lms)
              if any((cam.getEigenPose() @ pose @ np.append(p, 1.0))[2] < 0 for pose in poses)}
    assert selected == set(range(5)) - behind
