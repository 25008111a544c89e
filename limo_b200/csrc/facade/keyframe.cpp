// keyframe.cpp -- Keyframe of the facade: Tracklets -> measurements and the per-camera landmark projection (reference:
// keyframe_bundle_adjustment/src/keyframe.cpp:15-104).
#include "keyframe_bundle_adjustment/keyframe.hpp"

#include <algorithm>
#include <iterator>

namespace keyframe_bundle_adjustment {

Keyframe::Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, std::map<CameraId, Camera::Ptr> cameras,
                   std::map<LandmarkId, CameraIds> landmark_to_cameras, EigenPose p, FixationStatus fix_stat,
                   Plane ground_plane)
        : timestamp_(timestamp), cameras_(cameras), fixation_status_(fix_stat), local_ground_plane_(ground_plane),
          is_active_(true) {
    assignMeasurements(tracklets, landmark_to_cameras);
    assignPose(p);
}
Keyframe::Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, Camera::Ptr camera, EigenPose p,
                   FixationStatus fix_stat, Plane ground_plane)
        : timestamp_(timestamp), fixation_status_(fix_stat), local_ground_plane_(ground_plane), is_active_(true) {
    cameras_[0] = camera;
    assignMeasurements(tracklets, CameraId(0));
    assignPose(p);
}
void Keyframe::assignMeasurements(const Tracklets& tracklets, const std::map<LandmarkId, CameraIds>& lookup) {
    std::map<CameraId, Tracklets> out;
    for (const auto& track : tracklets.tracks)
        for (const auto& cam_id : lookup.at(track.id)) {
            out[cam_id].stamps = tracklets.stamps;
            out[cam_id].tracks.push_back(track);
        }
    for (const auto& el : out) assignMeasurements(el.second, el.first);
}
void Keyframe::assignMeasurements(const Tracklets& tracklets, const CameraId& cam_id) {
    auto iter = std::find(tracklets.stamps.begin(), tracklets.stamps.end(), this->timestamp_);
    const int index = int(std::distance(tracklets.stamps.begin(), iter));
    for (const auto& track : tracklets.tracks)
        if (index < int(track.feature_points.size())) measurements_[track.id][cam_id] = track.feature_points[index];
}
std::map<CameraId, Measurement> Keyframe::getMeasurements(LandmarkId lm_id) const {
    std::map<CameraId, Measurement> out;
    for (const auto& cam : cameras_)
        if (hasMeasurement(lm_id, cam.first)) out[cam.first] = getMeasurement(lm_id, cam.first);
    return out;
}
bool Keyframe::hasMeasurement(const LandmarkId& lm_id, const CameraId& cam_id) const {
    auto it = measurements_.find(lm_id);
    return it != measurements_.cend() && it->second.find(cam_id) != it->second.cend();
}
bool Keyframe::hasMeasurement(LandmarkId lm_id) const {
    for (const auto& cam : cameras_)
        if (hasMeasurement(lm_id, cam.first)) return true;
    return false;
}
std::map<CameraId, Eigen::Vector3d> Keyframe::getProjectedLandmarkPosition(
    const std::pair<LandmarkId, Landmark::ConstPtr>& id_lm) const {
    std::map<CameraId, Eigen::Vector3d> out;
    auto it = measurements_.find(id_lm.first);
    if (it == measurements_.cend()) return out;
    const Eigen::Vector3d p_vehicle = getEigenPose() * Eigen::Vector3d(id_lm.second->pos.data());
    for (const auto& cam_meas : it->second) out[cam_meas.first] = cameras_.at(cam_meas.first)->getEigenPose() * p_vehicle;
    return out;
}

}  // namespace keyframe_bundle_adjustment
