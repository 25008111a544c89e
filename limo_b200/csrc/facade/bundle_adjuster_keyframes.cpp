// bundle_adjuster_keyframes.cpp -- the replacement translation unit for limo's
// keyframe_bundle_adjustment/src/bundle_adjuster_keyframes.cpp (the other sources of the facade: definitions.cpp,
// keyframe.cpp, landmark_selection.cpp): window bookkeeping on the host exactly as the reference does it, the solve
// through the C ABI (kba_b200.h).
#include "keyframe_bundle_adjustment/bundle_adjuster_keyframes.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <iterator>
#include <sstream>
#include <stdexcept>

#include "kba_b200.h"

namespace keyframe_bundle_adjustment {

// ---- triangulation ---------------------------------------------------------------------------------------------------------
Eigen::Vector3d triangulate_rays(const std::vector<std::pair<EigenPose, Eigen::Vector3d>>& poses_rays) {
    Eigen::Matrix3d sum = Eigen::Matrix3d::Zero();
    Eigen::Vector3d rhs = Eigen::Vector3d::Zero();
    for (const auto& p_r : poses_rays) {
        const Eigen::Vector3d r = p_r.first.rotation() * p_r.second;
        const Eigen::Matrix3d cur = Eigen::Matrix3d::Identity() - r * r.transpose();
        sum += cur;
        rhs += cur * p_r.first.translation();
    }
    return sum.inverse() * rhs;  // the reference solves with a Jacobi SVD; identical for >= 2 non-parallel rays
}

// ---- exceptions --------------------------------------------------------------------------------------------------------------
BundleAdjusterKeyframes::NotEnoughKeyframesException::NotEnoughKeyframesException(size_t is, size_t should)
        : num_is(is), num_should_be(should) {
    std::stringstream ss;
    ss << "Not enough keyframes available in bundle_adjuster_keyframes. Should be " << num_should_be << " is " << num_is;
    msg = ss.str();
}
BundleAdjusterKeyframes::KeyframeNotFoundException::KeyframeNotFoundException(TimestampNSec timestamp) : ts_(timestamp) {
    std::stringstream ss;
    ss << "keyframe corresponding to timestamp " << ts_ << " nano seconds not found";
    msg = ss.str();
}

// ---- the adjuster ------------------------------------------------------------------------------------------------------------
BundleAdjusterKeyframes::BundleAdjusterKeyframes() : solver_time_sec(0.2) {
    landmark_selector_ = std::make_unique<LandmarkSelector>();
    landmark_selector_->addScheme(LandmarkRejectionSchemeCheirality::create());  // cpp:116-118
}
BundleAdjusterKeyframes::~BundleAdjusterKeyframes() {
    if (track_) kba_track_destroy(track_);
    if (handle_) kba_destroy(handle_);
}

void BundleAdjusterKeyframes::push(const std::vector<Keyframe>& kfs) { for (const auto& kf : kfs) push(kf); }

void BundleAdjusterKeyframes::push(const Keyframe& kf) {  // cpp:289-329
    keyframes_[kf.timestamp_] = std::make_shared<Keyframe>(kf);
    active_keyframe_ids_.insert(kf.timestamp_);
    for (const auto& m : kf.measurements_) {
        if (landmarks_.find(m.first) == landmarks_.cend()) {
            bool has_depth = false;
            for (const auto& cam_meas : m.second) if (cam_meas.second.d >= 0) has_depth = true;  // containsDepth, cpp:37-48
            v3 p;
            const bool success = has_depth ? calculateLandmark(kf, m.first, p) : calculateLandmark(m.first, p);
            if (!success) continue;
            landmarks_.insert(std::make_pair(m.first, std::make_shared<Landmark>(p, has_depth)));
            new_landmarks_.insert(m.first);
        }
        active_landmark_ids_.insert(m.first);
    }
    // the device-resident store gets the keyframe lazily, at the next solve() (trackPush): callers may still edit the stored
    // copy before (mono_lidar.cpp sets the pose prior after push)
}

bool BundleAdjusterKeyframes::calculateLandmark(const Keyframe& kf, const LandmarkId& lId, v3& posAbs) {  // cpp:332-355
    for (const auto& m : kf.measurements_.at(lId)) {
        if (m.second.d < 0) continue;
        const auto cam = kf.cameras_.at(m.first);
        const double z = static_cast<double>(m.second.d);
        const double x = (static_cast<double>(m.second.u) - cam->principal_point[0]) * z / cam->focal_length;
        const double y = (static_cast<double>(m.second.v) - cam->principal_point[1]) * z / cam->focal_length;
        posAbs = (cam->getEigenPose() * kf.getEigenPose()).inverse() * v3(x, y, z);
        return true;
    }
    return false;
}

bool BundleAdjusterKeyframes::calculateLandmark(const LandmarkId& lId, v3& posAbs) {  // cpp:125-159, 358-382
    std::vector<std::pair<EigenPose, v3>> poses_rays;
    for (const auto& id : active_keyframe_ids_) {
        const Keyframe& kf = *keyframes_.at(id);
        for (const auto& id_cam : kf.cameras_) {
            if (!kf.hasMeasurement(lId, id_cam.first)) continue;
            const Measurement& m = kf.getMeasurement(lId, id_cam.first);
            const v3 ray = (id_cam.second->intrin_inv * v3(static_cast<double>(m.u), static_cast<double>(m.v), 1.)).normalized();
            poses_rays.emplace_back((id_cam.second->getEigenPose() * kf.getEigenPose()).inverse(), ray);
        }
    }
    if (poses_rays.size() < 2) return false;
    posAbs = triangulate_rays(poses_rays);
    return true;
}

void BundleAdjusterKeyframes::updateLabels(const Tracklets& t, double shrubbery_weight) {  // cpp:388-431
    std::set<LandmarkId> outlier_ids;
    for (const auto& id : landmark_selector_->getOutliers())
        if (active_landmark_ids_.count(id)) outlier_ids.insert(id);
    for (const auto& track : t.tracks)
        if (track.is_outlier || labels_["outliers"].count(track.label)) outlier_ids.insert(track.id);
    landmark_selector_->clearOutliers();
    landmark_selector_->setOutlier(outlier_ids);
    for (const auto& track : t.tracks) {
        if (!active_landmark_ids_.count(track.id)) continue;
        if (labels_["shrubbery"].count(track.label)) { landmarks_.at(track.id)->weight = shrubbery_weight; dirty_weights_.insert(track.id); }
        landmarks_.at(track.id)->is_ground_plane = labels_["ground"].count(track.label) > 0;
    }
}

std::map<LandmarkId, Landmark::ConstPtr> BundleAdjusterKeyframes::filterLandmarksById(const std::set<LandmarkId>& ids) const {
    std::map<LandmarkId, Landmark::ConstPtr> out;
    for (const auto& id : ids) { auto it = landmarks_.find(id); if (it != landmarks_.cend()) out[id] = it->second; }
    return out;
}
std::map<LandmarkId, Landmark::ConstPtr> BundleAdjusterKeyframes::getActiveLandmarkConstPtrs() const { return filterLandmarksById(active_landmark_ids_); }
std::map<LandmarkId, Landmark::ConstPtr> BundleAdjusterKeyframes::getSelectedLandmarkConstPtrs() const { return filterLandmarksById(selected_landmark_ids_); }
std::map<KeyframeId, Keyframe::Ptr> BundleAdjusterKeyframes::getActiveKeyframePtrs() const {
    std::map<KeyframeId, Keyframe::Ptr> out;
    for (const auto& id : active_keyframe_ids_) out[id] = keyframes_.at(id);
    return out;
}
std::map<KeyframeId, Keyframe::ConstPtr> BundleAdjusterKeyframes::getActiveKeyframeConstPtrs() const {
    std::map<KeyframeId, Keyframe::ConstPtr> out;
    for (const auto& id : active_keyframe_ids_) out[id] = keyframes_.at(id);
    return out;
}
std::vector<std::pair<KeyframeId, Keyframe::Ptr>> BundleAdjusterKeyframes::getSortedIdsWithActiveKeyframePtrs() const {
    std::vector<std::pair<KeyframeId, Keyframe::Ptr>> v;
    for (const auto& kf : getActiveKeyframePtrs()) v.push_back(kf);
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return *(a.second) < *(b.second); });
    return v;
}
std::vector<Keyframe::Ptr> BundleAdjusterKeyframes::getSortedActiveKeyframePtrs() const {
    std::vector<Keyframe::Ptr> out;
    for (const auto& el : getSortedIdsWithActiveKeyframePtrs()) out.push_back(el.second);
    return out;
}

const Keyframe& BundleAdjusterKeyframes::getKeyframe(TimestampSec timestamp) const {  // cpp:989-1021
    if (keyframes_.size() == 0) throw NotEnoughKeyframesException(keyframes_.size(), 1);
    if (timestamp < 0.) {
        auto it = std::max_element(active_keyframe_ids_.cbegin(), active_keyframe_ids_.cend(), [&](const auto& a, const auto& b) {
            return keyframes_.at(a)->timestamp_ < keyframes_.at(b)->timestamp_;
        });
        return *keyframes_.at(*it);
    }
    const TimestampNSec ts_nsec = convert(timestamp);
    for (const auto& kf : keyframes_)
        if (kf.second->timestamp_ == ts_nsec) return *kf.second;
    throw KeyframeNotFoundException(ts_nsec);
}

void BundleAdjusterKeyframes::deactivateKeyframes(int min_num_connecting_landmarks, int min_size_optimization_window,
                                                  int max_size_optimization_window) {  // cpp:907-987
    auto sorted = getSortedIdsWithActiveKeyframePtrs();
    auto newest = sorted.back().second;
    int n = 0;
    for (auto it = sorted.crbegin(); it != sorted.crend(); ++it, ++n) {
        auto& cur = *it->second;
        if (n > max_size_optimization_window - 1) cur.is_active_ = false;
        else if (n < min_size_optimization_window - 1) cur.is_active_ = true;
        else {
            int common = 0;
            for (const auto& m : cur.measurements_) common += newest->measurements_.count(m.first) ? 1 : 0;
            cur.is_active_ = common > min_num_connecting_landmarks;
        }
        if (!cur.is_active_) active_keyframe_ids_.erase(it->first);
    }
    std::set<LandmarkId> new_active;
    for (const auto& id : active_keyframe_ids_)
        for (const auto& m : keyframes_.at(id)->measurements_)
            if (active_landmark_ids_.count(m.first)) new_active.insert(m.first);
    active_landmark_ids_ = new_active;
    auto rest = getSortedIdsWithActiveKeyframePtrs();
    rest[0].second->fixation_status_ = Keyframe::FixationStatus::Pose;
    rest[1].second->fixation_status_ = Keyframe::FixationStatus::Scale;
}

// Pack -> kba_solve_window -> scatter.  Replaces addActiveKeyframesToProblem / addKeyframeToProblem (cpp:498-627),
// addGroundPlaneResiduals (:517-562), the scale / plane regulariser set-up (:703-728, 769-818, 890-904) and
// robust_optimization::solveTrimmed (:765, :886).
std::string BundleAdjusterKeyframes::runWindow(const std::vector<Keyframe*>& kfs, const std::vector<LandmarkId>& lm_ids,
                                               bool motion_only, Keyframe* speed_kf) {
    ensureHandle();
    std::vector<double> kf_pose, kf_plane, cam_intr, cam_pose, lm_pos, lm_weight, gp_weight;
    std::vector<uint8_t> kf_fixed;
    std::vector<int32_t> lm_obs_ptr{0}, obs_kf, obs_cam, gp_lm, gp_kf;
    std::vector<float> obs_u, obs_v, obs_d;
    // Cameras are de-duplicated BY VALUE: the production node makes a new Camera object for every frame
    // (mono_lidar.cpp:112, mono_standalone.cpp:101), so pointer identity would give one "camera" per keyframe.
    using CamKey = std::array<double, 10>;  // focal length, principal point, pose_camera_vehicle
    auto cam_key = [](const Camera& c) {
        CamKey k{{c.focal_length, c.principal_point[0], c.principal_point[1]}};
        std::copy(c.pose_camera_vehicle.begin(), c.pose_camera_vehicle.end(), k.begin() + 3);
        return k;
    };
    std::map<CamKey, int> cam_index;
    for (const Keyframe* kf : kfs) {
        kf_pose.insert(kf_pose.end(), kf->pose_.begin(), kf->pose_.end());
        kf_fixed.push_back(!motion_only && kf->fixation_status_ == Keyframe::FixationStatus::Pose);  // cpp:198-219
        kf_plane.insert(kf_plane.end(), kf->local_ground_plane_.direction.begin(), kf->local_ground_plane_.direction.end());
        kf_plane.push_back(kf->local_ground_plane_.distance);
        for (const auto& c : kf->cameras_)
            if (cam_index.emplace(cam_key(*c.second), int(cam_index.size())).second) {
                cam_intr.insert(cam_intr.end(), {c.second->focal_length, c.second->principal_point[0], c.second->principal_point[1]});
                cam_pose.insert(cam_pose.end(), c.second->pose_camera_vehicle.begin(), c.second->pose_camera_vehicle.end());
            }
    }
    int n_depth = 0;
    for (const auto lm_id : lm_ids) {
        const Landmark& lm = *landmarks_.at(lm_id);
        lm_pos.insert(lm_pos.end(), lm.pos.begin(), lm.pos.end());
        lm_weight.push_back(lm.weight);
        for (size_t k = 0; k < kfs.size(); ++k) {
            auto it = kfs[k]->measurements_.find(lm_id);
            if (it == kfs[k]->measurements_.end()) continue;
            for (const auto& cm : it->second) {
                obs_kf.push_back(int32_t(k));
                obs_cam.push_back(cam_index.at(cam_key(*kfs[k]->cameras_.at(cm.first))));
                obs_u.push_back(cm.second.u); obs_v.push_back(cm.second.v); obs_d.push_back(cm.second.d);
                n_depth += cm.second.d > 0.0f;
            }
        }
        lm_obs_ptr.push_back(int32_t(obs_kf.size()));
    }
    kba_window w{};
    if (!motion_only) {
        for (size_t j = 0; j < lm_ids.size(); ++j) {  // addGroundPlaneResiduals(10.), cpp:517-562
            const Landmark& lm = *landmarks_.at(lm_ids[j]);
            if (!lm.is_ground_plane) continue;
            double min_dist = std::numeric_limits<double>::max();
            int best = -1;
            for (size_t k = 0; k < kfs.size(); ++k) {
                if (kfs[k]->local_ground_plane_.distance < -10.) continue;
                const double dist = (kfs[k]->getEigenPose() * v3(lm.pos.data())).norm();
                if (dist < min_dist) { min_dist = dist; best = int(k); }
            }
            if (best < 0 || !(min_dist < 25.)) continue;
            gp_lm.push_back(int32_t(j)); gp_kf.push_back(best); gp_weight.push_back(10. * (1. - min_dist / 25.));
        }
        const int n_gp = int(gp_lm.size());
        double scale_weight = 0.;  // cpp:703-716
        if (n_depth > 10 || n_gp > 10) { if (n_gp < 30) scale_weight = 1000. / (double(n_depth) + double(n_gp)); }
        else scale_weight = 1000.;
        if (scale_weight > 0 && kfs.size() > 1) {
            w.scale_kf0 = 0; w.scale_kf1 = 1; w.scale_weight = scale_weight;
            w.scale_value = (kfs[1]->getEigenPose() * kfs[0]->getEigenPose().inverse()).translation().norm();
        }
        if (n_gp > 0) w.plane_reg_weight = 10.;  // cpp:717-719
        w.plane_dist_fixed = n_depth < 10;       // cpp:722-728
    } else if (speed_kf && active_keyframe_ids_.size() > 2) {  // cpp:835-853
        auto sorted = getSortedActiveKeyframePtrs();
        const Keyframe& b0 = *sorted[sorted.size() - 1];
        const Keyframe& b1 = *sorted[sorted.size() - 2];
        const double rot_diff = calcQuaternionDiff(b0.pose_, b1.pose_);
        if (rot_diff < 0.03) {
            const double dt_cur = convert(speed_kf->timestamp_) - convert(b0.timestamp_);
            const double dt_before = convert(b0.timestamp_) - convert(b1.timestamp_);
            if (dt_cur <= 0. || dt_before <= 0.) throw std::runtime_error("In PoseRegularizationSpeed: invalid timestamps");
            const v3 v_before = (b0.getEigenPose() * b1.getEigenPose().inverse()).translation() / dt_before;
            const Pose T_ob = convert(b0.getEigenPose().inverse());
            w.speed_kf = 0; w.speed_weight = 1. * (1 - rot_diff / 0.03); w.speed_dt = dt_cur;
            for (int i = 0; i < 3; ++i) w.speed_v_before[i] = v_before[i];
            for (int i = 0; i < 7; ++i) w.speed_T_origin_before[i] = T_ob[i];
        }
    }
    w.landmarks_fixed = motion_only;
    w.n_kf = int(kfs.size()); w.n_cam = int(cam_index.size()); w.n_lm = int(lm_ids.size()); w.n_obs = int(obs_kf.size());
    w.n_gp = int(gp_lm.size());
    w.kf_pose = kf_pose.data(); w.kf_fixed = kf_fixed.data(); w.kf_plane = kf_plane.data();
    w.cam_intr = cam_intr.data(); w.cam_pose = cam_pose.data();
    w.lm_pos = lm_pos.data(); w.lm_weight = lm_weight.data(); w.lm_obs_ptr = lm_obs_ptr.data();
    w.obs_kf = obs_kf.data(); w.obs_cam = obs_cam.data(); w.obs_u = obs_u.data(); w.obs_v = obs_v.data(); w.obs_d = obs_d.data();
    w.gp_lm = gp_lm.data(); w.gp_kf = gp_kf.data(); w.gp_weight = gp_weight.data();

    kba_options opt;
    kba_default_options(&opt);
    opt.depth_thres = outlier_rejection_options_.depth_thres;
    opt.reprojection_thres = outlier_rejection_options_.reprojection_thres;
    opt.depth_quantile = outlier_rejection_options_.depth_quantile;
    opt.reprojection_quantile = outlier_rejection_options_.reprojection_quantile;
    opt.num_rounds_option = outlier_rejection_options_.num_iterations;
    opt.solver_time_sec = solver_time_sec;
    if (motion_only) {  // cpp:864-869
        opt.min_landmarks_for_trimming = 30;
        opt.num_trim_rounds = selected_landmark_ids_.size() > 30 ? outlier_rejection_options_.num_iterations : 0;
    }
    std::vector<double> out_pose(kf_pose.size()), out_plane(kf_plane.size()), out_lm(lm_pos.size() + 3);
    kba_result r{};
    r.kf_pose = out_pose.data(); r.kf_plane = out_plane.data(); r.lm_pos = out_lm.data();
    if (kba_solve_window(handle_, &w, &opt, &r) != KBA_OK) throw std::runtime_error(std::string("kba_b200: ") + kba_last_error());
    // what the rebuild path uploads: the window's arrays as passed (kba_batch_transfer_bytes reports the same for a batch)
    last_solve_h2d_ = (long long)(kf_pose.size() + kf_plane.size() + lm_pos.size() + lm_weight.size() + cam_intr.size() + cam_pose.size()) * 8 +
                      (long long)(lm_obs_ptr.size() + 2 * obs_kf.size()) * 4 + (long long)obs_u.size() * 12 + (long long)kf_fixed.size();

    for (size_t k = 0; k < kfs.size(); ++k) {  // the reference optimises in place (cpp:554-557, 592-593)
        std::copy_n(out_pose.begin() + 7 * k, 7, kfs[k]->pose_.begin());
        if (!motion_only && w.n_gp > 0) {
            std::copy_n(out_plane.begin() + 4 * k, 3, kfs[k]->local_ground_plane_.direction.begin());
            kfs[k]->local_ground_plane_.distance = out_plane[4 * k + 3];
        }
    }
    if (!motion_only)
        for (size_t j = 0; j < lm_ids.size(); ++j) std::copy_n(out_lm.begin() + 3 * j, 3, landmarks_.at(lm_ids[j])->pos.begin());

    static const char* term[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    std::stringstream ss;  // stands in for robust_optimization::Summary::FullReport (robust_solving.hpp:54-59)
    ss << "Merged summaries:\n";
    for (int i = 0; i < r.num_solves; ++i) {
        const kba_solve_summary& s = r.solves[i];
        ss << "--------------------------------------------------\nIteration No." << i << "\n"
           << "Residual blocks " << s.num_residual_blocks << ", landmarks " << s.num_landmarks << "; initial cost "
           << s.initial_cost << ", final cost " << s.final_cost << ", iterations " << s.num_iterations << " (successful "
           << s.num_successful_steps << "), termination " << term[s.termination < 3 ? s.termination : 2] << "\n";
    }
    if (r.status != KBA_OK)  // like a Ceres failure, not surfaced as an exception (reference: only text in the report)
        ss << "\nsolver did not finish (kba status " << r.status << "): the last accepted iterate was written back\n";
    ss << "\nDuration solveTrimmed=" << r.time_sec << " sec\n";
    return ss.str();
}

std::string BundleAdjusterKeyframes::solve() {  // cpp:629-767
    if (keyframes_.size() < 3) throw NotEnoughKeyframesException(keyframes_.size(), 3);
    selected_landmark_ids_ = landmark_selector_->select(getActiveLandmarkConstPtrs(), getActiveKeyframeConstPtrs());
    std::vector<Keyframe*> kfs;
    for (const auto& id : active_keyframe_ids_) kfs.push_back(keyframes_.at(id).get());
    std::vector<LandmarkId> lm_ids(selected_landmark_ids_.begin(), selected_landmark_ids_.end());
    std::string report;
    if (persistent_window_ && !track_failed_ && solveTracked(kfs, lm_ids, report)) return report;
    return runWindow(kfs, lm_ids, false, nullptr);
}

// ---- persistent device-resident window ---------------------------------------------------------------------------------------
bool BundleAdjusterKeyframes::ensureHandle() {
    if (!handle_ && kba_create(&handle_, 0) != KBA_OK) throw std::runtime_error(std::string("kba_b200: ") + kba_last_error());
    return true;
}

namespace {
std::array<double, 10> camera_value(const Camera& c) {
    std::array<double, 10> k{{c.focal_length, c.principal_point[0], c.principal_point[1]}};
    std::copy(c.pose_camera_vehicle.begin(), c.pose_camera_vehicle.end(), k.begin() + 3);
    return k;
}
constexpr int kTrackKeyframes = 256, kTrackLandmarks = 1 << 17, kTrackMeasurements = 1 << 21;
constexpr int kTrackWinKeyframes = 30, kTrackWinLandmarks = 16384, kTrackWinObservations = 1 << 18;
}  // namespace

// the keyframe's measurements go to the device ONCE (landmark slot, camera, u, v, d in measurements_ order: landmark id, then
// camera id -- the order addKeyframeToProblem enumerates, cpp:564-627); false: this keyframe cannot live in the store
bool BundleAdjusterKeyframes::trackPush(const Keyframe& kf) {
    ensureHandle();
    if (!track_) {  // created from the cameras of the first keyframe that reaches it
        std::vector<double> intr, pose;
        for (const auto& c : kf.cameras_) {
            const auto v = camera_value(*c.second);
            if (std::find(track_cams_.begin(), track_cams_.end(), v) != track_cams_.end()) continue;
            track_cams_.push_back(v);
            intr.insert(intr.end(), v.begin(), v.begin() + 3);
            pose.insert(pose.end(), v.begin() + 3, v.end());
        }
        kba_track_caps caps{kTrackKeyframes, kTrackLandmarks, kTrackMeasurements, kTrackWinKeyframes, kTrackWinLandmarks, kTrackWinObservations, 0};
        if (kba_track_create(handle_, &caps, int(track_cams_.size()), intr.data(), pose.data(), &track_) != KBA_OK) return false;
        for (int i = kTrackKeyframes - 1; i >= 0; --i) free_kf_slots_.push_back(i);
    }
    std::map<CameraId, int> cam_of;
    for (const auto& c : kf.cameras_) {
        const auto it = std::find(track_cams_.begin(), track_cams_.end(), camera_value(*c.second));
        if (it == track_cams_.end()) return false;  // a camera the store does not know
        cam_of[c.first] = int(it - track_cams_.begin());
    }
    if (free_kf_slots_.empty()) {  // reclaim the slots of the oldest keyframes that are no longer active
        for (auto it = kf_slot_.begin(); it != kf_slot_.end() && free_kf_slots_.size() < 32;) {
            if (active_keyframe_ids_.count(it->first)) { ++it; continue; }
            kba_track_drop_keyframe(track_, it->second);
            free_kf_slots_.push_back(it->second);
            it = kf_slot_.erase(it);
        }
        if (free_kf_slots_.empty()) return false;
    }
    std::vector<int32_t> lm, cam;
    std::vector<float> u, v, d;
    for (const auto& m : kf.measurements_) {
        auto it = lm_slot_.find(m.first);
        if (it == lm_slot_.end()) {
            if (int(lm_slot_.size()) >= kTrackLandmarks) return false;
            it = lm_slot_.emplace(m.first, int(lm_slot_.size())).first;
        }
        for (const auto& cm : m.second) {
            lm.push_back(it->second); cam.push_back(cam_of.at(cm.first));
            u.push_back(cm.second.u); v.push_back(cm.second.v); d.push_back(cm.second.d);
        }
    }
    const int slot = free_kf_slots_.back();
    double plane[4] = {kf.local_ground_plane_.direction[0], kf.local_ground_plane_.direction[1], kf.local_ground_plane_.direction[2], kf.local_ground_plane_.distance};
    if (kba_track_push_keyframe(track_, slot, kf.pose_.data(), plane, int(lm.size()), lm.data(), cam.data(), u.data(), v.data(), d.data()) != KBA_OK) return false;
    free_kf_slots_.pop_back();
    kf_slot_[kf.timestamp_] = slot;
    return true;
}

// solve() on the device-resident window: only the selection goes up.  false: not possible for this window (caller rebuilds).
bool BundleAdjusterKeyframes::solveTracked(const std::vector<Keyframe*>& kfs, const std::vector<LandmarkId>& lm_ids, std::string& report) {
    if (int(kfs.size()) > kTrackWinKeyframes || int(lm_ids.size()) > kTrackWinLandmarks) return false;
    for (const auto lm_id : lm_ids)
        if (landmarks_.at(lm_id)->is_ground_plane) return false;  // ground-plane residuals: plane blocks, rebuild path
    for (const Keyframe* kf : kfs)
        if (!kf_slot_.count(kf->timestamp_) && !trackPush(*kf)) { track_failed_ = true; return false; }
    // state the host may have changed since the last solve: poses / planes of the active keyframes, new landmarks, weights
    const int n_kf = int(kfs.size()), n_lm = int(lm_ids.size());
    std::vector<int32_t> kf_slots, lm_slots;
    std::vector<uint8_t> fixed;
    std::vector<double> poses, planes;
    for (const Keyframe* kf : kfs) {
        kf_slots.push_back(kf_slot_.at(kf->timestamp_));
        fixed.push_back(kf->fixation_status_ == Keyframe::FixationStatus::Pose);
        poses.insert(poses.end(), kf->pose_.begin(), kf->pose_.end());
        planes.insert(planes.end(), kf->local_ground_plane_.direction.begin(), kf->local_ground_plane_.direction.end());
        planes.push_back(kf->local_ground_plane_.distance);
    }
    if (kba_track_set_keyframe_poses(track_, n_kf, kf_slots.data(), poses.data(), planes.data()) != KBA_OK) { track_failed_ = true; return false; }
    {
        std::vector<int32_t> slots; std::vector<double> pos, wgt;
        for (const auto id : new_landmarks_) {
            auto it = lm_slot_.find(id);
            if (it == lm_slot_.end()) continue;  // created, but its keyframe has not reached the store yet
            const Landmark& lm = *landmarks_.at(id);
            slots.push_back(it->second); pos.insert(pos.end(), lm.pos.begin(), lm.pos.end()); wgt.push_back(lm.weight);
        }
        if (!slots.empty() && kba_track_set_landmarks(track_, int(slots.size()), slots.data(), pos.data(), wgt.data()) != KBA_OK) { track_failed_ = true; return false; }
        for (const auto id : std::set<LandmarkId>(new_landmarks_)) if (lm_slot_.count(id)) new_landmarks_.erase(id);
        slots.clear(); wgt.clear();
        for (const auto id : dirty_weights_) { auto it = lm_slot_.find(id); if (it != lm_slot_.end()) { slots.push_back(it->second); wgt.push_back(landmarks_.at(id)->weight); } }
        if (!slots.empty() && kba_track_set_landmarks(track_, int(slots.size()), slots.data(), nullptr, wgt.data()) != KBA_OK) { track_failed_ = true; return false; }
        dirty_weights_.clear();
    }
    for (const auto id : lm_ids) {
        auto it = lm_slot_.find(id);
        if (it == lm_slot_.end()) return false;  // selected but never measured by a stored keyframe: let the rebuild path decide
        lm_slots.push_back(it->second);
    }
    kba_window sel{};
    sel.n_kf = n_kf; sel.n_lm = n_lm;
    sel.scale_kf0 = 0; sel.scale_kf1 = 1;
    sel.scale_weight = -1.;  // the reference's rule (cpp:703-716), evaluated on the device from the gathered window
    sel.scale_value = n_kf > 1 ? (kfs[1]->getEigenPose() * kfs[0]->getEigenPose().inverse()).translation().norm() : 0.;
    kba_options opt;
    kba_default_options(&opt);
    opt.depth_thres = outlier_rejection_options_.depth_thres;
    opt.reprojection_thres = outlier_rejection_options_.reprojection_thres;
    opt.depth_quantile = outlier_rejection_options_.depth_quantile;
    opt.reprojection_quantile = outlier_rejection_options_.reprojection_quantile;
    opt.num_rounds_option = outlier_rejection_options_.num_iterations;
    opt.solver_time_sec = solver_time_sec;
    std::vector<double> out_pose(7 * size_t(n_kf)), out_plane(4 * size_t(n_kf)), out_lm(3 * size_t(n_lm) + 3);
    kba_result r{};
    r.kf_pose = out_pose.data(); r.kf_plane = out_plane.data(); r.lm_pos = out_lm.data();
    const int rc = kba_track_solve(track_, n_kf, kf_slots.data(), fixed.data(), n_lm, lm_slots.data(), &sel, &opt, &r);
    if (rc == KBA_ERR_CAPACITY) return false;  // e.g. more observations than the store's window capacity: rebuild
    if (rc != KBA_OK) throw std::runtime_error(std::string("kba_b200: ") + kba_last_error());
    int64_t h2d = 0, d2h = 0, pushes = 0;
    kba_track_transfer_bytes(track_, &h2d, &d2h, &pushes);
    push_h2d_ = (long long)pushes;
    last_solve_h2d_ = (long long)h2d + (long long)n_kf * (7 + 4) * 8;  // the selection lists + the active keyframes' poses
    for (int k = 0; k < n_kf; ++k) std::copy_n(out_pose.begin() + 7 * k, 7, kfs[k]->pose_.begin());  // in place, as the reference (cpp:554-557)
    for (int j = 0; j < n_lm; ++j) std::copy_n(out_lm.begin() + 3 * j, 3, landmarks_.at(lm_ids[j])->pos.begin());
    static const char* term[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    std::stringstream ss;
    ss << "Merged summaries:\n";
    for (int i = 0; i < r.num_solves; ++i) {
        const kba_solve_summary& s = r.solves[i];
        ss << "--------------------------------------------------\nIteration No." << i << "\n"
           << "Residual blocks " << s.num_residual_blocks << ", landmarks " << s.num_landmarks << "; initial cost "
           << s.initial_cost << ", final cost " << s.final_cost << ", iterations " << s.num_iterations << " (successful "
           << s.num_successful_steps << "), termination " << term[s.termination < 3 ? s.termination : 2] << "\n";
    }
    if (r.status != KBA_OK) ss << "\nsolver did not finish (kba status " << r.status << "): the last accepted iterate was written back\n";
    ss << "\nDuration solveTrimmed=" << r.time_sec << " sec (device-resident window)\n";
    report = ss.str();
    return true;
}

std::string BundleAdjusterKeyframes::adjustPoseOnly(Keyframe& kf) {  // cpp:820-888
    selected_landmark_ids_ = landmark_selector_->getLastSelection();
    std::vector<LandmarkId> lm_ids;
    for (const auto& m : kf.measurements_)
        if (selected_landmark_ids_.count(m.first)) lm_ids.push_back(m.first);  // landmarks_.at() would throw like the reference
    return runWindow({&kf}, lm_ids, true, &kf);
}

}  // namespace keyframe_bundle_adjustment
