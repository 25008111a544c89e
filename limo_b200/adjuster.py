"""Host-side mirror of limo's `BundleAdjusterKeyframes` window management (Python edition).

Same names, argument meaning and error behaviour as the reference class
(keyframe_bundle_adjustment/include/keyframe_bundle_adjustment/bundle_adjuster_keyframes.hpp:40-335); the part of
solve()/adjustPoseOnly() that the reference hands to Ceres is packed into a `kba_window` and handed to the C-ABI
(`kba_solve_window`, include/kba_b200.h).  The C++ facade under include/keyframe_bundle_adjustment/ is the drop-in
for C++ callers; this module drives the same C-ABI from Python for tests, benchmarks and scripting.

There is no CPU path here: the default backend is the CUDA library and raises if it cannot be loaded.
"""
import numpy as np

from . import geometry as g
from .capi_types import Window


class NotEnoughKeyframesException(Exception):
    """bundle_adjuster_keyframes.hpp:59-70"""

    def __init__(self, num_is, num_should_be):
        super().__init__("Not enough keyframes available in bundle_adjuster_keyframes. Should be %d is %d"
                         % (num_should_be, num_is))
        self.num_is, self.num_should_be = num_is, num_should_be


class KeyframeNotFoundException(Exception):
    """bundle_adjuster_keyframes.hpp:71-77"""


class FeaturePoint:
    """matches_msg_types/feature_point.hpp: float u, v, d (d = -1: no depth)."""
    __slots__ = ("u", "v", "d")

    def __init__(self, u, v, d=-1.0):
        self.u, self.v, self.d = np.float32(u), np.float32(v), np.float32(d)


class Tracklet:
    """matches_msg_types/tracklet.hpp"""

    def __init__(self, id, feature_points=None, is_outlier=False, label=-2, age=0):
        self.id, self.feature_points = id, list(feature_points or [])
        self.is_outlier, self.label, self.age = is_outlier, label, age


class Tracklets:
    """matches_msg_types/tracklets.hpp"""

    def __init__(self, stamps=None, tracks=None):
        self.stamps, self.tracks = list(stamps or []), list(tracks or [])


class Plane:
    """definitions.hpp:27-34: default distance -DBL_MAX means 'no ground plane'."""

    def __init__(self, direction=(0.0, 0.0, 1.0), distance=-np.finfo(float).max):
        self.direction = np.array(direction, dtype=float)
        self.distance = float(distance)


class Landmark:
    """definitions.hpp:42-68"""

    def __init__(self, pos, has_measured_depth=False):
        self.pos = np.array(pos, dtype=float)
        self.has_measured_depth = has_measured_depth
        self.is_ground_plane = False
        self.weight = 1.0


class Camera:
    """definitions.hpp:93-124, definitions.cpp:30-56"""

    def __init__(self, f, pp, pose_cam_veh):
        self.focal_length = float(f)
        self.principal_point = np.array(pp, dtype=float)
        self.pose_camera_vehicle = g.iso_to_pose(pose_cam_veh)
        self.intrin_inv = np.linalg.inv(self.getIntrinsicMatrix())

    def getIntrinsicMatrix(self):
        f, c = self.focal_length, self.principal_point
        return np.array([[f, 0, c[0]], [0, f, c[1]], [0, 0, 1.0]])

    def getEigenPose(self):
        return g.pose_to_iso(self.pose_camera_vehicle)


class Keyframe:
    """keyframe.hpp:27-196, keyframe.cpp"""
    FIX_POSE, FIX_SCALE, FIX_NONE = "Pose", "Scale", "None"

    def __init__(self, timestamp, tracklets, cameras, pose, fix_stat="None", ground_plane=None,
                 landmark_to_cameras=None):
        self.timestamp_ = int(timestamp)
        self.fixation_status_ = fix_stat
        self.local_ground_plane_ = ground_plane if ground_plane is not None else Plane()
        self.is_active_ = True
        self.measurements_ = {}  # lm id -> {cam id -> FeaturePoint}
        if isinstance(cameras, dict):
            self.cameras_ = dict(cameras)
            per_cam = {}
            for track in tracklets.tracks:  # keyframe.cpp:44-58
                for cam_id in landmark_to_cameras[track.id]:
                    per_cam.setdefault(cam_id, Tracklets(tracklets.stamps, [])).tracks.append(track)
            for cam_id in sorted(per_cam):
                self.assignMeasurements(per_cam[cam_id], cam_id)
        else:
            self.cameras_ = {0: cameras}
            self.assignMeasurements(tracklets, 0)
        self.pose_ = g.iso_to_pose(pose)

    def assignMeasurements(self, tracklets, cam_id):
        """keyframe.cpp:61-75"""
        try:
            index = tracklets.stamps.index(self.timestamp_)
        except ValueError:
            index = len(tracklets.stamps)
        for track in tracklets.tracks:
            if index < len(track.feature_points):
                self.measurements_.setdefault(track.id, {})[cam_id] = track.feature_points[index]

    def getEigenPose(self):
        return g.pose_to_iso(self.pose_)

    def hasMeasurement(self, lm_id, cam_id=None):
        m = self.measurements_.get(lm_id)
        if m is None:
            return False
        return True if cam_id is None else cam_id in m

    def getProjectedLandmarkPosition(self, lm_id, lm):
        """keyframe.cpp:81-104"""
        m = self.measurements_.get(lm_id)
        if m is None:
            return {}
        p_vehicle = g.apply(self.getEigenPose(), lm.pos)
        return {cam_id: g.apply(self.cameras_[cam_id].getEigenPose(), p_vehicle) for cam_id in m}


class LandmarkSelector:
    """landmark_selector.hpp:40-345 with the library default scheme only: cheirality rejection
    (src/landmark_selection_scheme_cheirality.cpp:22-60)."""

    def __init__(self):
        self.outlier_ids_ = set()
        self.unselected_lms_ = {}
        self.last_time_seen_ = {}
        self.last_selected_lms_ = set()

    def select(self, landmarks, kfs):
        non_rejected = {i: lm for i, lm in landmarks.items() if i not in self.outlier_ids_}
        selection = set()
        for lm_id, lm in non_rejected.items():
            ok = True
            for kf in kfs.values():
                if kf.is_active_ and any(p[2] < 0.0 for p in kf.getProjectedLandmarkPosition(lm_id, lm).values()):
                    ok = False
                    break
            if ok:
                selection.add(lm_id)
        cur_ts = max(kf.timestamp_ for kf in kfs.values())
        for lm_id in set(landmarks) - selection:  # landmark_selector.hpp:238-241
            self.unselected_lms_[lm_id] = self.unselected_lms_.get(lm_id, 0) + 1
            self.last_time_seen_[lm_id] = cur_ts
        oldest = cur_ts - int(10.0 * 1e9)
        for lm_id in [i for i, t in self.last_time_seen_.items() if t < oldest]:
            self.unselected_lms_.pop(lm_id, None)
            self.last_time_seen_.pop(lm_id, None)
        self.last_selected_lms_ = set(selection)
        return selection

    def getLastSelection(self):
        return set(self.last_selected_lms_)

    def getOutliers(self):
        return self.outlier_ids_

    def clearOutliers(self):
        self.outlier_ids_ = set()

    def setOutlier(self, ids):
        self.outlier_ids_ |= set(ids) if not isinstance(ids, int) else {ids}


class OutlierRejectionOptions:
    """bundle_adjuster_keyframes.hpp:79-89"""

    def __init__(self):
        self.depth_thres = 0.16
        self.reprojection_thres = 1.6
        self.depth_quantile = 0.95
        self.reprojection_quantile = 0.95
        self.num_iterations = 1


def triangulate_rays(poses_rays):
    """Triangulator::triangulate_rays (internal/triangulator.hpp:51-75); poses are origin <- camera."""
    A = np.zeros((3, 3)); b = np.zeros(3)
    for T, ray in poses_rays:
        r = T[:3, :3] @ ray
        M = np.eye(3) - np.outer(r, r)
        A += M
        b += M @ T[:3, 3]
    U, s, Vt = np.linalg.svd(A)  # jacobiSvd(...).solve(rhs)
    tol = np.finfo(float).eps * 3 * s[0]
    sinv = np.where(s > tol, 1.0 / np.where(s > tol, s, 1.0), 0.0)
    return Vt.T @ (sinv * (U.T @ b))


class BundleAdjusterKeyframes:
    """Window state + problem assembly of the reference class; the numerical solve goes through `backend`.

    backend: object with solve_window(Window, options) -> Result (limo_b200.capi.Handle by default).
    """

    def __init__(self, backend=None):
        self.keyframes_ = {}
        self.landmarks_ = {}
        self.active_keyframe_ids_ = set()
        self.active_landmark_ids_ = set()
        self.selected_landmark_ids_ = set()
        self.outlier_rejection_options_ = OutlierRejectionOptions()
        self.landmark_selector_ = LandmarkSelector()
        self.labels_ = {"outliers": {23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33}, "shrubbery": {21},
                        "ground": {6, 7, 8, 9, 10}}
        self.solver_time_sec = 0.2  # cpp:115
        self._backend = backend
        self.last_window = None
        self.last_result = None

    # ---- backend ----------------------------------------------------------------------------------
    def _get_backend(self):
        if self._backend is None:
            from . import capi  # raises if the CUDA library is missing: there is no CPU fallback
            self._backend = capi.Handle()
        return self._backend

    def set_solver_time(self, solver_time_sec):
        self.solver_time_sec = solver_time_sec

    # ---- push / landmark creation (cpp:282-382) -----------------------------------------------------
    def push(self, kf):
        if isinstance(kf, (list, tuple)):
            for k in kf:
                self.push(k)
            return
        import copy
        self.keyframes_[kf.timestamp_] = copy.copy(kf)
        self.active_keyframe_ids_.add(kf.timestamp_)
        for lm_id in sorted(kf.measurements_):
            if lm_id not in self.landmarks_:
                has_depth = any(m.d >= 0 for m in kf.measurements_[lm_id].values())  # containsDepth, cpp:37-48
                p = self.calculateLandmark(kf, lm_id) if has_depth else self.calculateLandmark(None, lm_id)
                if p is None:
                    continue
                self.landmarks_[lm_id] = Landmark(p, has_depth)
            self.active_landmark_ids_.add(lm_id)

    def calculateLandmark(self, kf, lm_id):
        if kf is not None:  # cpp:332-355, depth back-projection
            for cam_id in sorted(kf.measurements_[lm_id]):
                m = kf.measurements_[lm_id][cam_id]
                if m.d < 0:
                    continue
                cam = kf.cameras_[cam_id]
                z = float(m.d)
                x = (float(m.u) - cam.principal_point[0]) * z / cam.focal_length
                y = (float(m.v) - cam.principal_point[1]) * z / cam.focal_length
                return g.apply(g.iso_inv(cam.getEigenPose() @ kf.getEigenPose()), [x, y, z])
            return None
        poses_rays = []  # cpp:125-159, 358-382
        for kf_id in sorted(self.active_keyframe_ids_):
            k = self.keyframes_[kf_id]
            for cam_id in sorted(k.cameras_):
                if k.hasMeasurement(lm_id, cam_id):
                    cam = k.cameras_[cam_id]
                    m = k.measurements_[lm_id][cam_id]
                    ray = cam.intrin_inv @ np.array([float(m.u), float(m.v), 1.0])
                    ray /= np.linalg.norm(ray)
                    poses_rays.append((g.iso_inv(cam.getEigenPose() @ k.getEigenPose()), ray))
        if len(poses_rays) < 2:
            return None
        return triangulate_rays(poses_rays)

    # ---- labels (cpp:388-431) -------------------------------------------------------------------------
    def updateLabels(self, t, shrubbery_weight=1.0):
        outlier_ids = {i for i in self.landmark_selector_.getOutliers() if i in self.active_landmark_ids_}
        for track in t.tracks:
            if track.is_outlier or track.label in self.labels_["outliers"]:
                outlier_ids.add(track.id)
        self.landmark_selector_.clearOutliers()
        self.landmark_selector_.setOutlier(outlier_ids)
        for track in t.tracks:
            if track.id in self.active_landmark_ids_:
                if track.label in self.labels_["shrubbery"]:
                    self.landmarks_[track.id].weight = shrubbery_weight
                self.landmarks_[track.id].is_ground_plane = track.label in self.labels_["ground"]

    # ---- accessors ---------------------------------------------------------------------------------------
    def getActiveKeyframePtrs(self):
        return {i: self.keyframes_[i] for i in sorted(self.active_keyframe_ids_)}

    getActiveKeyframeConstPtrs = getActiveKeyframePtrs

    def getSortedActiveKeyframePtrs(self):
        return sorted(self.getActiveKeyframePtrs().values(), key=lambda k: k.timestamp_)

    def getActiveLandmarkConstPtrs(self):
        return {i: self.landmarks_[i] for i in sorted(self.active_landmark_ids_) if i in self.landmarks_}

    def getSelectedLandmarkConstPtrs(self):
        return {i: self.landmarks_[i] for i in sorted(self.selected_landmark_ids_) if i in self.landmarks_}

    def getKeyframe(self, timestamp=-1.0):
        if not self.keyframes_:
            raise NotEnoughKeyframesException(0, 1)
        if timestamp < 0:
            return self.keyframes_[max(self.active_keyframe_ids_, key=lambda i: self.keyframes_[i].timestamp_)]
        ts = int(timestamp * 1e9)
        for k in self.keyframes_.values():
            if k.timestamp_ == ts:
                return k
        raise KeyframeNotFoundException(ts)

    # ---- window management (cpp:907-987) -----------------------------------------------------------------
    def deactivateKeyframes(self, min_num_connecting_landmarks=3, min_size_optimization_window=4,
                            max_size_optimization_window=20):
        sorted_kfs = self.getSortedActiveKeyframePtrs()
        newest = sorted_kfs[-1]
        for n, kf in enumerate(reversed(sorted_kfs)):
            if n > max_size_optimization_window - 1:
                kf.is_active_ = False
            elif n < min_size_optimization_window - 1:
                kf.is_active_ = True
            else:
                common = set(kf.measurements_) & set(newest.measurements_)
                kf.is_active_ = len(common) > min_num_connecting_landmarks
            if not kf.is_active_:
                self.active_keyframe_ids_.discard(kf.timestamp_)
        new_active = set()
        for kf_id in self.active_keyframe_ids_:
            for lm_id in self.keyframes_[kf_id].measurements_:
                if lm_id in self.active_landmark_ids_:
                    new_active.add(lm_id)
        self.active_landmark_ids_ = new_active
        ordered = sorted(self.active_keyframe_ids_, key=lambda i: self.keyframes_[i].timestamp_)
        self.keyframes_[ordered[0]].fixation_status_ = Keyframe.FIX_POSE
        self.keyframes_[ordered[1]].fixation_status_ = Keyframe.FIX_SCALE

    # ---- problem assembly -----------------------------------------------------------------------------------
    def _options(self, min_landmarks_for_trimming):
        backend = self._get_backend()
        opt = backend.default_options()
        o = self.outlier_rejection_options_
        opt.depth_thres, opt.reprojection_thres = o.depth_thres, o.reprojection_thres
        opt.depth_quantile, opt.reprojection_quantile = o.depth_quantile, o.reprojection_quantile
        opt.num_rounds_option = int(o.num_iterations)
        opt.num_trim_rounds = -1
        opt.min_landmarks_for_trimming = min_landmarks_for_trimming
        opt.solver_time_sec = self.solver_time_sec
        return opt

    def _pack(self, kfs, lm_ids, landmarks_fixed=False):
        """addKeyframeToProblem (cpp:564-627) as a landmark-major CSR.  kfs: keyframes in ascending id order."""
        kf_index = {kf.timestamp_: i for i, kf in enumerate(kfs)}
        # cameras are de-duplicated by VALUE: the production node creates a new Camera object per frame
        # (mono_lidar.cpp:112), object identity would give one camera per keyframe
        def cam_key(cam):
            return (cam.focal_length, *np.asarray(cam.principal_point, dtype=float), *np.asarray(cam.pose_camera_vehicle, dtype=float))
        cam_list, cam_index = [], {}
        for kf in kfs:
            for cam_id in sorted(kf.cameras_):
                cam = kf.cameras_[cam_id]
                if cam_key(cam) not in cam_index:
                    cam_index[cam_key(cam)] = len(cam_list)
                    cam_list.append(cam)
        lm_ids = sorted(lm_ids)
        ptr, okf, ocam, ou, ov, od = [0], [], [], [], [], []
        for lm_id in lm_ids:
            for kf in kfs:
                m = kf.measurements_.get(lm_id)
                if m is None:
                    continue
                for cam_id in sorted(m):
                    fp = m[cam_id]
                    okf.append(kf_index[kf.timestamp_]); ocam.append(cam_index[cam_key(kf.cameras_[cam_id])])
                    ou.append(fp.u); ov.append(fp.v); od.append(fp.d)
            ptr.append(len(okf))
        return dict(
            kf_pose=[kf.pose_ for kf in kfs],
            kf_fixed=[1 if kf.fixation_status_ == Keyframe.FIX_POSE else 0 for kf in kfs],
            kf_plane=[list(kf.local_ground_plane_.direction) + [kf.local_ground_plane_.distance] for kf in kfs],
            cam_intr=[[c.focal_length, c.principal_point[0], c.principal_point[1]] for c in cam_list],
            cam_pose=[c.pose_camera_vehicle for c in cam_list],
            lm_pos=[self.landmarks_[i].pos for i in lm_ids] if lm_ids else np.zeros((0, 3)),
            lm_weight=[self.landmarks_[i].weight for i in lm_ids], lm_obs_ptr=ptr,
            obs_kf=okf, obs_cam=ocam, obs_u=ou, obs_v=ov, obs_d=od, landmarks_fixed=landmarks_fixed), lm_ids, kf_index

    def solve(self):
        """cpp:629-767"""
        if len(self.keyframes_) < 3:
            raise NotEnoughKeyframesException(len(self.keyframes_), 3)
        active_landmarks = self.getActiveLandmarkConstPtrs()
        active_keyframes = self.getActiveKeyframeConstPtrs()
        self.selected_landmark_ids_ = self.landmark_selector_.select(active_landmarks, active_keyframes)
        kfs = [self.keyframes_[i] for i in sorted(self.active_keyframe_ids_)]
        args, lm_ids, kf_index = self._pack(kfs, self.selected_landmark_ids_)
        lm_index = {i: n for n, i in enumerate(lm_ids)}
        n_depth = int(np.sum(np.asarray(args["obs_d"], dtype=np.float32) > 0))

        # addGroundPlaneResiduals(10.) (cpp:517-562)
        gp_lm, gp_kf, gp_w = [], [], []
        for lm_id in lm_ids:
            lm = self.landmarks_[lm_id]
            if not lm.is_ground_plane:
                continue
            min_dist, best = np.finfo(float).max, None
            for kf in kfs:
                if kf.local_ground_plane_.distance < -10.0:
                    continue
                dist = np.linalg.norm(g.apply(kf.getEigenPose(), lm.pos))
                if dist < min_dist:
                    min_dist, best = dist, kf
            if best is None:
                continue
            if min_dist < 25.0:
                gp_lm.append(lm_index[lm_id]); gp_kf.append(kf_index[best.timestamp_])
                gp_w.append(10.0 * (1.0 - min_dist / 25.0))
        n_gp = len(gp_lm)
        if n_gp:
            args.update(gp_lm=gp_lm, gp_kf=gp_kf, gp_weight=gp_w)

        # scale handling (cpp:703-716), addScaleRegularization (cpp:890-904)
        scale_weight = 0.0
        if n_depth > 10 or n_gp > 10:
            if n_gp < 30:
                scale_weight = 1000.0 / (float(n_depth) + float(n_gp))
        else:
            scale_weight = 1000.0
        if scale_weight > 0 and len(kfs) > 1:
            T = kfs[1].getEigenPose() @ g.iso_inv(kfs[0].getEigenPose())
            args.update(scale_kf0=0, scale_kf1=1, scale_weight=scale_weight,
                        scale_value=float(np.linalg.norm(T[:3, 3])))
        if n_gp > 0:
            args.update(plane_reg_weight=10.0)  # cpp:717-719
        args.update(plane_dist_fixed=(n_depth < 10))  # cpp:722-728

        win = Window(**args)
        res = self._get_backend().solve_window(win, self._options(100))
        self._scatter(win, res, kfs, lm_ids)
        return self._report(res)

    def adjustPoseOnly(self, kf):
        """cpp:820-888: motion-only refinement of one frame against the last landmark selection."""
        self.selected_landmark_ids_ = self.landmark_selector_.getLastSelection()
        lm_ids = [i for i in self.selected_landmark_ids_ if i in kf.measurements_ and i in self.landmarks_]
        args, lm_ids, _ = self._pack([kf], lm_ids, landmarks_fixed=True)
        args["kf_fixed"] = [0]  # deactivatePoseParameters only visits active_keyframe_ids_ (cpp:198-219); kf is not pushed
        if len(self.active_keyframe_ids_) > 2:
            s = self.getSortedActiveKeyframePtrs()
            b, b2 = s[-1], s[-2]
            rot_diff = g.quaternion_angle(b.pose_, b2.pose_)
            if rot_diff < 0.03:
                ts_cur, ts_b, ts_b2 = kf.timestamp_ * 1e-9, b.timestamp_ * 1e-9, b2.timestamp_ * 1e-9
                dt_cur, dt_before = ts_cur - ts_b, ts_b - ts_b2
                if dt_cur <= 0.0 or dt_before <= 0.0:
                    raise RuntimeError("In PoseRegularizationSpeed: invalid timestamps")
                Tb, Tb2 = b.getEigenPose(), b2.getEigenPose()
                v_before = (Tb @ g.iso_inv(Tb2))[:3, 3] / dt_before
                args.update(speed_kf=0, speed_weight=1.0 * (1 - rot_diff / 0.03), speed_dt=dt_cur,
                            speed_v_before=v_before, speed_T_origin_before=g.iso_to_pose(g.iso_inv(Tb)))
        win = Window(**args)
        opt = self._options(30)
        if len(self.selected_landmark_ids_) <= 30:
            opt.num_trim_rounds = 0
        else:
            opt.num_trim_rounds = int(self.outlier_rejection_options_.num_iterations)
        opt.gp_quantile = 1.0
        res = self._get_backend().solve_window(win, opt)
        kf.pose_ = np.array(res.kf_pose[0])
        self.last_window, self.last_result = win, res
        return self._report(res)

    def _scatter(self, win, res, kfs, lm_ids):
        for i, kf in enumerate(kfs):
            kf.pose_ = np.array(res.kf_pose[i])
            kf.local_ground_plane_.direction = np.array(res.kf_plane[i, :3]) if win.kf_plane is not None else kf.local_ground_plane_.direction
            if win.kf_plane is not None:
                kf.local_ground_plane_.distance = float(res.kf_plane[i, 3])
        for n, lm_id in enumerate(lm_ids):
            self.landmarks_[lm_id].pos = np.array(res.lm_pos[n])
        self.last_window, self.last_result = win, res

    @staticmethod
    def _report(res):
        """robust_optimization::Summary::FullReport (robust_solving.hpp:54-59), abbreviated."""
        term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}
        lines = ["Merged summaries:"]
        for i, s in enumerate(res.solves):
            lines.append("--------------------------------------------------\nIteration No.%d" % i)
            lines.append("Residual blocks %d, landmarks %d; initial cost %.6e, final cost %.6e, iterations %d "
                         "(successful %d), termination %s" % (s.num_residual_blocks, s.num_landmarks, s.initial_cost,
                                                               s.final_cost, s.num_iterations, s.num_successful_steps,
                                                               term.get(s.termination, "?")))
        lines.append("Duration solveTrimmed=%g sec" % res.c.time_sec)
        return "\n".join(lines) + "\n"
