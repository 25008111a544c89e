// definitions.cpp -- pose-array <-> Eigen conversions and the Camera type of the facade (reference:
// keyframe_bundle_adjustment/src/definitions.cpp, internal/definitions.hpp:75-124).
#include "keyframe_bundle_adjustment/internal/definitions.hpp"

namespace keyframe_bundle_adjustment {

Pose convert(EigenPose p) {
    Eigen::Quaterniond q(p.rotation());
    return Pose{{q.w(), q.x(), q.y(), q.z(), p.translation()[0], p.translation()[1], p.translation()[2]}};
}
EigenPose convert(const Pose& pose) {
    EigenPose p = EigenPose::Identity();
    p.translate(Eigen::Vector3d(pose[4], pose[5], pose[6]));
    p.rotate(Eigen::Quaterniond(pose[0], pose[1], pose[2], pose[3]));
    return p;
}
TimestampSec convert(const TimestampNSec& ts) { return static_cast<TimestampSec>(ts * 1e-09); }
TimestampNSec convert(const TimestampSec& ts) { return static_cast<TimestampNSec>(ts * 1e09); }
double calcQuaternionDiff(const Pose& p0, const Pose& p1) {
    Eigen::Quaterniond q0(p0[0], p0[1], p0[2], p0[3]), q1(p1[0], p1[1], p1[2], p1[3]);
    return Eigen::AngleAxisd(q1.inverse() * q0).angle();
}

Camera::Camera(double f, const Eigen::Vector2d& pp, const EigenPose& pose_cam_veh) : focal_length(f), principal_point(pp) {
    pose_camera_vehicle = convert(pose_cam_veh);
    intrin_inv = getIntrinsicMatrix().inverse();
}
Eigen::Matrix3d Camera::getIntrinsicMatrix() const {
    Eigen::Matrix3d K = Eigen::Matrix3d::Zero();
    K(0, 0) = focal_length; K(0, 2) = principal_point[0]; K(1, 1) = focal_length; K(1, 2) = principal_point[1]; K(2, 2) = 1.;
    return K;
}
EigenPose Camera::getEigenPose() const { return convert(pose_camera_vehicle); }

}  // namespace keyframe_bundle_adjustment
