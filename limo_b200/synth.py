"""Synthetic optimisation windows with the shapes of BASELINE.json's configs (distribution choices: SURVEY.md 8(d)).

The window that comes out is what BundleAdjusterKeyframes::solve() would hand to solveTrimmed after push() of every
keyframe: landmarks initialised by the reference's push() rule (depth back-projection if the creating keyframe has a
depth measurement, else two-view triangulation; bundle_adjuster_keyframes.cpp:289-382), cheirality-selected
(landmark_selection_scheme_cheirality.cpp:22-60), ground-plane residuals attached to the nearest keyframe
(cpp:517-562) and the scale regulariser weighted as in cpp:703-716.

Deviations from SURVEY 8(d), all forced: the camera extrinsic keeps the test rig's lever arm but looks along the
direction of travel (with the rig's literal axes the camera looks sideways and no track survives 30 keyframes); the RNG is numpy's MT19937 (std::mt19937_64 streams are not reproducible
from numpy), and config 1 ("5 KF / 200 LM / 1.5k obs") is infeasible for a mono rig (200 x 5 = 1000 observations
at most), so its observation count is capped at what the tracks can hold.
"""
import numpy as np

from . import geometry as g
from .capi_types import Window

IMG_W, IMG_H = 1242.0, 375.0
F, CX, CY = 718.856, 607.1928, 185.2157

CONFIGS = {
    1: dict(n_kf=5, n_lm=200, n_obs=1500, depth_frac=0.0, gp_frac=0.0, outlier_frac=0.0),
    2: dict(n_kf=30, n_lm=3000, n_obs=40000, depth_frac=0.4, gp_frac=0.0, outlier_frac=0.05),
    3: dict(n_kf=30, n_lm=3000, n_obs=40000, depth_frac=0.4, gp_frac=0.1, outlier_frac=0.05),
    5: dict(n_kf=100, n_lm=20000, n_obs=300000, depth_frac=0.4, gp_frac=0.0, outlier_frac=0.05),
}


def cam_extrinsics():
    """camera <- vehicle, KITTI-like: optical axis = vehicle x (forward), image x = -vehicle y, image y = -vehicle z;
    camera centre 1.5 m ahead, 0.2 m left and 1.35 m above the vehicle origin (the offsets of the reference test rig,
    test/keyframe_bundle_adjustment.cpp:808-814, whose axes however do not look along the direction of travel)."""
    R = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    c_veh = np.array([1.5, 0.2, 1.35])
    return g.iso(R, -R @ c_veh)


def _rpy(roll, pitch, yaw):
    return g.angle_axis(yaw, [0, 0, 1.0]) @ g.angle_axis(pitch, [0, 1.0, 0]) @ g.angle_axis(roll, [1.0, 0, 0])


def _trajectory(rng, n_kf):
    """origin <- vehicle poses W_k: 1 m forward per keyframe, small attitude noise."""
    W = [np.eye(4)]
    for _ in range(n_kf - 1):
        step = g.iso(_rpy(rng.normal(0, 0.002), rng.normal(0, 0.002), rng.normal(0, 0.01)), [1.0, 0.0, 0.0])
        W.append(W[-1] @ step)
    return W


def _track_lengths(rng, n_lm, n_obs, n_kf, start):
    room = n_kf - start
    mean = max(n_obs / n_lm, 2.0)
    ell = 2 + rng.geometric(1.0 / max(mean - 1.0, 1.0 + 1e-9), size=n_lm) - 1 if mean > 2.0 else np.full(n_lm, 2)
    ell = np.minimum(np.maximum(ell, 2), room)
    target = min(n_obs, int(room.sum()))
    total = int(ell.sum())
    while total != target:  # nudge random tracks until the total is exact
        if total < target:
            cand = np.nonzero(ell < room)[0]
            pick = rng.choice(cand, size=min(target - total, len(cand)), replace=False)
            ell[pick] += 1
        else:
            cand = np.nonzero(ell > 2)[0]
            pick = rng.choice(cand, size=min(total - target, len(cand)), replace=False)
            ell[pick] -= 1
        total = int(ell.sum())
    return ell


def _project(Rc, tc, Rk, tk, P):
    """camera coordinates of points P[n,3] seen from keyframes (Rk[n,3,3], tk[n,3])"""
    px = np.einsum("nij,nj->ni", Rk, P) + tk
    return px @ Rc.T + tc


def make_window(config=2, seed=None, n_kf=None, n_lm=None, n_obs=None, depth_frac=None, gp_frac=None,
                outlier_frac=None, pose_noise=(0.3 * np.pi / 180.0, 0.05), pixel_noise=0.5, depth_noise=0.05,
                return_truth=False):
    cfg = dict(CONFIGS[config])
    for k, v in dict(n_kf=n_kf, n_lm=n_lm, n_obs=n_obs, depth_frac=depth_frac, gp_frac=gp_frac,
                     outlier_frac=outlier_frac).items():
        if v is not None:
            cfg[k] = v
    n_kf, n_lm, n_obs = cfg["n_kf"], cfg["n_lm"], cfg["n_obs"]
    rng = np.random.Generator(np.random.MT19937(0xBA5E0000 + config if seed is None else seed))

    T_cv = cam_extrinsics()
    Rc, tc = T_cv[:3, :3], T_cv[:3, 3]
    W = _trajectory(rng, n_kf)
    T_true = [g.iso_inv(w) for w in W]  # keyframe <- origin
    T_init = []
    for k in range(n_kf):
        d = g.iso(_rpy(*rng.normal(0, pose_noise[0], 3)), rng.normal(0, pose_noise[1], 3))
        T_init.append(T_true[k] @ d)
    Rt = np.stack([T[:3, :3] for T in T_true]); tt = np.stack([T[:3, 3] for T in T_true])
    Ri = np.stack([T[:3, :3] for T in T_init]); ti = np.stack([T[:3, 3] for T in T_init])

    start = rng.integers(0, n_kf - 1, size=n_lm)
    ell = _track_lengths(rng, n_lm, n_obs, n_kf, start)
    n_gp_target = int(round(cfg["gp_frac"] * n_lm))
    is_gp = np.zeros(n_lm, dtype=bool)
    is_gp[rng.choice(n_lm, size=n_gp_target, replace=False)] = True
    is_out = np.zeros(n_lm, dtype=bool)
    is_out[rng.choice(n_lm, size=int(round(cfg["outlier_frac"] * n_lm)), replace=False)] = True
    is_out &= ~is_gp

    # per-landmark storage (ragged via max length)
    L = int(ell.max())
    P_true = np.zeros((n_lm, 3)); P_init = np.zeros((n_lm, 3))
    U = np.zeros((n_lm, L), dtype=np.float32); V = np.zeros((n_lm, L), dtype=np.float32)
    D = np.full((n_lm, L), -1.0, dtype=np.float32)
    todo = np.arange(n_lm)
    Kinv = np.array([[1 / F, 0, -CX / F], [0, 1 / F, -CY / F], [0, 0, 1.0]])
    for _attempt in range(5000):
        if len(todo) == 0:
            break
        if _attempt >= 400 and _attempt % 10 == 0:
            # tracks that cannot be placed in the field of view over their whole length (long tracks on the 100-keyframe
            # trajectory of config 5) are shortened step by step; configs 1-3 never get here (< 100 attempts)
            ell[todo] = np.maximum(2, ell[todo] - 1)
        m = len(todo)
        s = start[todo]
        # candidate position: pixel uniform in the image, depth U(4, 60) in the first observing camera
        z = rng.uniform(np.minimum(np.maximum(4.0, ell[todo] + 1.0), 59.0), 60.0, m)  # U(4,60) | still ahead at track end
        pc = np.stack([(rng.uniform(0, IMG_W, m) - CX) / F * z, (rng.uniform(0, IMG_H, m) - CY) / F * z, z], axis=1)
        gp_here = is_gp[todo]
        px = (pc - tc) @ Rc  # vehicle frame of keyframe s: Rc^T (pc - tc)
        if gp_here.any():  # ground points: keep the bearing, intersect with the plane z_veh = -0.31
            o_veh = -Rc.T @ tc  # camera centre in the vehicle frame
            dirv = px - o_veh
            lam = (-0.31 - o_veh[2]) / np.where(np.abs(dirv[:, 2]) > 1e-9, dirv[:, 2], 1e-9)
            px_g = o_veh + lam[:, None] * dirv
            px = np.where(gp_here[:, None], px_g, px)
        p_o = np.einsum("nji,nj->ni", Rt[s], px - tt[s])
        ok = np.ones(m, dtype=bool)
        u_ = np.zeros((m, L), dtype=np.float32); v_ = np.zeros((m, L), dtype=np.float32)
        d_ = np.full((m, L), -1.0, dtype=np.float32)
        off = np.where(is_out[todo], 1.0, 0.0)[:, None] * rng.uniform(5.0, 30.0, (m, 2)) * rng.choice([-1.0, 1.0], (m, 2))
        for i in range(L):
            act = ell[todo] > i
            k = np.minimum(s + i, n_kf - 1)
            c = _project(Rc, tc, Rt[k], tt[k], p_o)
            uu = F * c[:, 0] / c[:, 2] + CX
            vv = F * c[:, 1] / c[:, 2] + CY
            vis = (c[:, 2] > 0.5) & (uu >= 0) & (uu < IMG_W) & (vv >= 0) & (vv < IMG_H)
            ok &= vis | ~act
            u_[:, i] = (uu + rng.normal(0, pixel_noise, m) + off[:, 0]).astype(np.float32)
            v_[:, i] = (vv + rng.normal(0, pixel_noise, m) + off[:, 1]).astype(np.float32)
            has_d = rng.uniform(size=m) < cfg["depth_frac"]
            if i == 0 and _attempt > 300 and cfg["depth_frac"] > 0:
                has_d[:] = True  # hard-to-place (long, triangulated) tracks: give the creating keyframe a lidar depth
            d_[:, i] = np.where(has_d & act, c[:, 2] + rng.normal(0, depth_noise, m), -1.0).astype(np.float32)
        # landmark initialisation by the push() rule, evaluated with the PERTURBED poses
        k0, k1 = s, np.minimum(s + 1, n_kf - 1)
        def backproject(k, uu, vv, dd):
            pc_ = np.stack([(uu - CX) * dd / F, (vv - CY) * dd / F, dd], axis=1)
            px_ = (pc_ - tc) @ Rc
            return np.einsum("nji,nj->ni", Ri[k], px_ - ti[k])
        u0, v0, d0 = u_[:, 0].astype(float), v_[:, 0].astype(float), d_[:, 0].astype(float)
        u1, v1, d1 = u_[:, 1].astype(float), v_[:, 1].astype(float), d_[:, 1].astype(float)
        p_bp0 = backproject(k0, u0, v0, np.where(d0 >= 0, d0, 1.0))
        p_bp1 = backproject(k1, u1, v1, np.where(d1 >= 0, d1, 1.0))
        # two-ray triangulation (internal/triangulator.hpp:51-75)
        def ray(k, uu, vv):
            r = np.stack([uu, vv, np.ones_like(uu)], axis=1) @ Kinv.T
            r /= np.linalg.norm(r, axis=1, keepdims=True)
            Roc = np.einsum("nji,jk->nik", Ri[k], Rc.T)        # origin <- camera rotation
            toc = np.einsum("nji,nj->ni", Ri[k], (-Rc.T @ tc)[None, :] - ti[k])
            return np.einsum("nij,nj->ni", Roc, r), toc
        r0, c0 = ray(k0, u0, v0); r1, c1 = ray(k1, u1, v1)
        I3 = np.eye(3)[None]
        M0 = I3 - r0[:, :, None] * r0[:, None, :]; M1 = I3 - r1[:, :, None] * r1[:, None, :]
        A = M0 + M1
        bvec = np.einsum("nij,nj->ni", M0, c0) + np.einsum("nij,nj->ni", M1, c1)
        det = np.linalg.det(A)
        good = np.abs(det) > 1e-12
        p_tri = np.zeros((m, 3))
        p_tri[good] = np.linalg.solve(A[good], bvec[good][:, :, None])[:, :, 0]
        p_i = np.where((d0 >= 0)[:, None], p_bp0, np.where((d1 >= 0)[:, None], p_bp1, p_tri))
        ok &= good | (d0 >= 0) | (d1 >= 0)
        # cheirality selection at the initial state: z_cam >= 0 in every observing keyframe
        for i in range(L):
            act = ell[todo] > i
            k = np.minimum(s + i, n_kf - 1)
            c = _project(Rc, tc, Ri[k], ti[k], p_i)
            ok &= (c[:, 2] >= 0.05) | ~act  # keep clear of the |z| < 0.01 evaluation-failure band
        acc = todo[ok]
        P_true[acc] = p_o[ok]; P_init[acc] = p_i[ok]
        U[acc] = u_[ok]; V[acc] = v_[ok]; D[acc] = d_[ok]
        todo = todo[~ok]
    if len(todo):
        raise RuntimeError("synthetic window generation did not converge (%d landmarks left)" % len(todo))

    # flatten landmark-major
    ptr = np.zeros(n_lm + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(ell)
    idx_l = np.repeat(np.arange(n_lm), ell)
    idx_i = np.arange(ptr[-1]) - np.repeat(ptr[:-1], ell)
    obs_kf = (start[idx_l] + idx_i).astype(np.int32)
    obs_u, obs_v, obs_d = U[idx_l, idx_i], V[idx_l, idx_i], D[idx_l, idx_i]
    n_depth = int((obs_d > 0).sum())

    kf_pose = np.stack([g.iso_to_pose(T) for T in T_init])
    kf_fixed = np.zeros(n_kf, dtype=np.uint8); kf_fixed[0] = 1
    args = dict(kf_pose=kf_pose, kf_fixed=kf_fixed, cam_intr=[[F, CX, CY]], cam_pose=[g.iso_to_pose(T_cv)],
                lm_pos=P_init, lm_weight=np.ones(n_lm), lm_obs_ptr=ptr, obs_kf=obs_kf, obs_u=obs_u, obs_v=obs_v,
                obs_d=obs_d)
    n_gp = 0
    if is_gp.any():  # addGroundPlaneResiduals (cpp:517-562) at the initial poses
        gp_idx = np.nonzero(is_gp)[0]
        pk = np.einsum("kij,nj->nki", Ri, P_init[gp_idx]) + ti[None]
        dist = np.linalg.norm(pk, axis=2)
        best = np.argmin(dist, axis=1)
        md = dist[np.arange(len(gp_idx)), best]
        keep = md < 25.0
        args.update(gp_lm=gp_idx[keep].astype(np.int32), gp_kf=best[keep].astype(np.int32),
                    gp_weight=10.0 * (1.0 - md[keep] / 25.0))
        n_gp = int(keep.sum())
        plane = np.tile(np.array([0.0, 0.0, 1.0, 0.31]), (n_kf, 1))  # keyframe_ba_monolid.launch:56
        args.update(kf_plane=plane)
        if n_gp > 0:
            args.update(plane_reg_weight=10.0)
    scale_weight = 0.0
    if n_depth > 10 or n_gp > 10:
        if n_gp < 30:
            scale_weight = 1000.0 / (float(n_depth) + float(n_gp))
    else:
        scale_weight = 1000.0
    if scale_weight > 0:
        T10 = T_init[1] @ g.iso_inv(T_init[0])
        args.update(scale_kf0=0, scale_kf1=1, scale_weight=scale_weight, scale_value=float(np.linalg.norm(T10[:3, 3])))
    args.update(plane_dist_fixed=(n_depth < 10))
    win = Window(**args)
    if return_truth:
        truth = dict(kf_pose=np.stack([g.iso_to_pose(T) for T in T_true]), lm_pos=P_true, is_outlier=is_out, is_gp=is_gp)
        return win, truth
    return win


# ---- lidar scene of BASELINE config 4 --------------------------------------------------------------------------
def velo_to_cam():
    """camera <- velodyne, KITTI nominal: camera z = velodyne x, camera x = -velodyne y, camera y = -velodyne z;
    the camera sits 0.27 m ahead of and 0.08 m below the lidar."""
    R = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    return g.iso(R, [0.0, -0.08, -0.27])


def make_lidar_scene(seed=0xBA5E0004, n_rings=64, n_azimuth=1875, n_features=2000, n_facades=20):
    """64 rings x 1875 azimuth steps = 120k returns of a synthetic scene: ground plane z = -1.73 m (velodyne frame) plus
    vertical planar facades 5..60 m away; xyz + intensity float32 (KITTI .bin layout); features at uniform random pixels."""
    rng = np.random.Generator(np.random.MT19937(seed))
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_rings))
    az = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)
    E, A = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    rng_ = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        lam = -1.73 / d[:, 2]
    rng_ = np.where((d[:, 2] < 0) & (lam < 120.0), lam, rng_)
    for _ in range(n_facades):  # vertical facade: plane n . p = dist, bounded patch
        ang = rng.uniform(-np.pi, np.pi)
        n = np.array([np.cos(ang), np.sin(ang), 0.0])
        dist = rng.uniform(5.0, 60.0)
        centre = n * dist
        half_w, top = rng.uniform(2.0, 8.0), rng.uniform(1.0, 6.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            lam = dist / (d @ n)
        hit = d * lam[:, None]
        tang = np.array([-n[1], n[0], 0.0])
        ok = (lam > 0) & (np.abs((hit - centre) @ tang) < half_w) & (hit[:, 2] < top) & (hit[:, 2] > -1.73)
        rng_ = np.where(ok & (lam < rng_), lam, rng_)
    keep = np.isfinite(rng_)
    pts = d[keep] * (rng_[keep] + rng.normal(0, 0.01, keep.sum()))[:, None]
    cloud = np.concatenate([pts, rng.uniform(0, 1, (len(pts), 1))], axis=1).astype(np.float32)
    feats = np.stack([rng.uniform(0, IMG_W, n_features), rng.uniform(0, IMG_H, n_features)], axis=1).astype(np.float32)
    return cloud, g.iso_to_pose(velo_to_cam()), np.array([F, CX, CY]), feats
