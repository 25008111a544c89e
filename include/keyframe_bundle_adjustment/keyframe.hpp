// keyframe.hpp -- one keyframe of the window (reference: keyframe_bundle_adjustment/include/keyframe_bundle_adjustment/
// keyframe.hpp:27-196, src/keyframe.cpp): time stamp (= id), pose keyframe <- origin, cameras of the rig, fixation
// status, local ground plane and the measurements landmark -> camera -> (u, v, d).  Public data members as in the
// reference: callers read poses and planes back from them after solve().
#pragma once
#include <map>
#include <memory>
#include <utility>

#include "internal/definitions.hpp"

namespace keyframe_bundle_adjustment {

class Keyframe {
public:
    enum class FixationStatus { Pose, Scale, None };
    using Ptr = std::shared_ptr<Keyframe>;
    using ConstPtr = std::shared_ptr<const Keyframe>;
    Keyframe() {}
    Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, std::map<CameraId, Camera::Ptr> cameras,
             std::map<LandmarkId, CameraIds> landmark_to_cameras, EigenPose p,
             FixationStatus fix_stat = FixationStatus::None, Plane ground_plane = Plane());
    Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, Camera::Ptr camera, EigenPose p,
             FixationStatus fix_stat = FixationStatus::None, Plane ground_plane = Plane());
    bool operator<(const Keyframe& kf) const { return timestamp_ < kf.timestamp_; }
    void assignMeasurements(const Tracklets&, const CameraId&);
    void assignMeasurements(const Tracklets& tracklets, const std::map<LandmarkId, CameraIds>& landmark_lookup);
    void assignPose(const EigenPose& p) { pose_ = convert(p); }
    Measurement& getMeasurement(LandmarkId lm_id, CameraId cam_id) { return measurements_.at(lm_id).at(cam_id); }
    const Measurement& getMeasurement(LandmarkId lm_id, CameraId cam_id) const { return measurements_.at(lm_id).at(cam_id); }
    std::map<CameraId, Measurement> getMeasurements(LandmarkId lm_id) const;
    bool hasMeasurement(const LandmarkId& lm_id, const CameraId& cam_id) const;
    bool hasMeasurement(LandmarkId lm_id) const;
    // landmark position in the frame of every camera that observes it (keyframe.cpp:81-104)
    std::map<CameraId, Eigen::Vector3d> getProjectedLandmarkPosition(const std::pair<LandmarkId, Landmark::ConstPtr>& landmark_origin) const;
    EigenPose getEigenPose() const { return convert(pose_); }
    std::shared_ptr<Pose> getPosePtr() const { return std::make_shared<Pose>(pose_); }

    TimestampNSec timestamp_;
    std::map<CameraId, Camera::Ptr> cameras_;
    FixationStatus fixation_status_;
    Pose pose_;
    Plane local_ground_plane_;
    std::map<LandmarkId, std::map<CameraId, Measurement>> measurements_;
    bool is_active_;
};

}  // namespace keyframe_bundle_adjustment
