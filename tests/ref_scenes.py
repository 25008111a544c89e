"""The synthetic scenes of the reference's own gtest file, restated.

Source: keyframe_bundle_adjustment/test/keyframe_bundle_adjustment.cpp
  getPoses :232-249, makeTracklets :288-356, makeTrackletsDepth :359-417,
  evaluate_bundle_adjustment :419-609, evaluate_bundle_adjustment_depth :860-1087.
The reference's add_noise() helpers (:180-216) construct a fresh default-seeded std::default_random_engine on
every call, so each call returns the same draw; the constants below were computed with libstdc++ (g++ 13).
"""
import numpy as np

from limo_b200 import geometry as g
from limo_b200.adjuster import (BundleAdjusterKeyframes, Camera, FeaturePoint, Keyframe, Tracklet, Tracklets)

NOISE_ANGLE_5DEG = -0.01064352254023776
NOISE_VEC3_02_01_01 = np.array([-0.024393156828319384, 0.068428994379655481, 0.0033269476420492391])
NOISE_PIX_15 = np.array([-0.18294867621239536, 1.0264349156948323])


def noise_angle(sigma):
    return 0.0 if sigma == 0 else NOISE_ANGLE_5DEG * sigma / (5.0 * np.pi / 180.0)


def noise_vec3(sig):
    return np.zeros(3) if not any(sig) else NOISE_VEC3_02_01_01 * np.array(sig) / np.array([0.2, 0.1, 0.1])


def noise_pix(sig):
    return np.zeros(2) if not any(sig[:2]) else NOISE_PIX_15 * np.array(sig[:2]) / 1.5


def get_poses(noise_ang, noise_transl):
    """getPoses :232-249"""
    z = np.array([0.0, 0.0, 1.0])
    P = [np.eye(4)]
    p = g.translate(P[0], [-1.5, 0.0, -2.0]); p = g.rotate(p, g.angle_axis(-0.05, z)); P.append(p)
    p = g.translate(P[1], np.array([-2.0, 0.0, 0.0]) + noise_vec3(noise_transl))
    p = g.rotate(p, g.angle_axis(-0.05 + noise_angle(noise_ang), z)); P.append(p)
    p = g.translate(P[2], np.array([-1.5, -0.1, 0.0]) + noise_vec3(noise_transl)); P.append(p)
    p = g.translate(P[3], np.array([-2.9, -0.0, 0.0]) + noise_vec3(noise_transl)); P.append(p)
    return P


def mono_extrinsics():
    """test :808-814"""
    p = np.eye(4)
    p = g.rotate(p, g.angle_axis(np.pi / 2.0, [1.0, 0.0, 0.0]))
    p = g.rotate(p, g.angle_axis(np.pi / 2.0, [0.0, 0.0, 1.0]))
    p = g.translate(p, [-1.5, 0.2, -1.35])
    return g.iso_inv(p)


def stereo_extrinsics():
    """test :834-840"""
    p = mono_extrinsics()
    p2 = g.translate(p, [0.0, -0.5, 0.0])
    p2 = g.rotate(p2, g.angle_axis(np.pi / 18.0, [0.0, 1.0, 0.0]))
    p2 = g.rotate(p2, g.angle_axis(np.pi / 18.0, [1.0, 0.0, 0.0]))
    return [p, p2]


LMS_SOLVE = [(10., 0.5, 5.5), (11., 1., 6.5), (14., -5., 6.), (9., 1., 5.), (16., -1., 4.)]       # :452-456
LMS_SOLVE_DEPTH = [(10., 3., 5.5), (11., 1., 6.5), (14., -5., 6.), (9., 1., 5.), (16., -1., 4.)]  # :894-898
F, PP = 600.0, (200.0, 100.0)


def make_tracklets(poses_gt, lms, cameras, noise_lms, landmark_to_cameras, with_depth):
    """makeTracklets :288-356 / makeTrackletsDepth :359-417"""
    stamps = list(range(len(poses_gt)))
    tracks = [Tracklet(i) for i in range(len(lms))]
    for pose in poses_gt:
        for i, lm in enumerate(lms):
            cam = cameras[landmark_to_cameras[i][0]]
            lm_cam = g.apply(cam.getEigenPose() @ pose, lm)
            proj = cam.getIntrinsicMatrix() @ lm_cam
            proj = proj / proj[2]
            uv = proj[:2] + noise_pix(noise_lms)
            if with_depth:
                d = lm_cam[2] + 0.0  # noise_z is always 0 in the reference's calls
                tracks[i].feature_points.append(FeaturePoint(uv[0], uv[1], d))
            else:
                tracks[i].feature_points.append(FeaturePoint(uv[0], uv[1]))
    return Tracklets(stamps, tracks)


def build_adjuster(backend, noise_lms, noise_poses, extrinsics, with_depth=False, motion_only=False):
    """Common part of evaluate_bundle_adjustment(:419-609) / _depth(:860-1087) up to the solve call."""
    poses_gt = get_poses(0.0, (0.0, 0.0, 0.0))
    noisy = get_poses(noise_poses[0], noise_poses[1:])
    for i in range(2, len(noisy)):  # :446-449 unit-normalised translation
        noisy[i] = noisy[i].copy()
        noisy[i][:3, 3] /= np.linalg.norm(noisy[i][:3, 3])
    lms = [np.array(p) for p in (LMS_SOLVE_DEPTH if with_depth else LMS_SOLVE)]
    cameras = {i: Camera(F, PP, T) for i, T in enumerate(extrinsics)}
    l2c = {i: [i % len(extrinsics)] for i in range(len(lms))}
    ts = make_tracklets(poses_gt, lms, cameras, noise_lms, l2c, with_depth)
    b = BundleAdjusterKeyframes(backend=backend)
    b.set_solver_time(20.0)
    max_ind = len(poses_gt) - 1 if with_depth else len(poses_gt)  # :928 (depth variant pushes keyframes 0..3)
    if motion_only:
        for i in range(max_ind):
            noisy[i] = poses_gt[i]
    fix = [Keyframe.FIX_POSE, Keyframe.FIX_SCALE] + [Keyframe.FIX_NONE] * 8
    for i in range(max_ind):
        if len(extrinsics) == 1:
            b.push(Keyframe(i, ts, Camera(F, PP, extrinsics[0]) if i == 0 else b.keyframes_[0].cameras_[0], noisy[i], fix[i]))
        else:
            b.push(Keyframe(i, ts, cameras, noisy[i], fix[i], landmark_to_cameras=l2c))
    return b, poses_gt, noisy, lms, ts, cameras, l2c
