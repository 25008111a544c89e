// internal/definitions.hpp -- the data model of the window (reference: keyframe_bundle_adjustment/include/
// keyframe_bundle_adjustment/internal/definitions.hpp:13-124, src/definitions.cpp): ids, the 7-vector pose convention,
// Plane, Landmark, Camera and the conversions between pose arrays and Eigen transforms.  OpenCV and Ceres headers of the
// reference are not needed by any caller of this interface and are dropped; ResidualId becomes an opaque integer.
#pragma once
#include <array>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <vector>

#include "matches_msg_types/tracklets.hpp"
#include "mini_eigen.hpp"

namespace keyframe_bundle_adjustment {

using CameraId = unsigned long;
using TimestampNSec = matches_msg_types::TimestampNSec;
using TimestampSec = double;
using LandmarkId = unsigned long;
using KeyframeId = unsigned long;
using CameraIds = std::vector<CameraId>;
using PoseId = KeyframeId;
using EigenPose = Eigen::Isometry3d;
using Pose = std::array<double, 7>;  // quaternion (w, x, y, z), translation: p_keyframe = R(q) p_origin + t
using ResidualId = long;             // opaque (ceres::ResidualBlockId in the reference)
using Direction = std::array<double, 3>;
using FeaturePoint = matches_msg_types::FeaturePoint;
using Tracklet = matches_msg_types::Tracklet;
using Tracklets = matches_msg_types::Tracklets;
using Measurement = FeaturePoint;

struct Plane {  // definitions.hpp:27-34; the default distance marks "no plane estimate"
    Plane() : direction{{0., 0., 1.}}, distance(-std::numeric_limits<double>::max()) {}
    Direction direction;
    double distance;
};

struct Landmark {  // definitions.hpp:42-68
    using Ptr = std::shared_ptr<Landmark>;
    using ConstPtr = std::shared_ptr<const Landmark>;
    Landmark() {}
    Landmark(const Eigen::Vector3d& p, bool has_depth = false) : pos{{p[0], p[1], p[2]}}, has_measured_depth(has_depth) {}
    std::array<double, 3> pos;
    bool has_measured_depth{false};
    bool is_ground_plane{false};
    double weight{1.};
};

Pose convert(EigenPose p);                                   // definitions.cpp:14-28
EigenPose convert(const Pose& pose);                         // definitions.hpp:75-88
TimestampSec convert(const TimestampNSec& ts);               // definitions.cpp:68-70
TimestampNSec convert(const TimestampSec& ts);
double calcQuaternionDiff(const Pose& p0, const Pose& p1);   // definitions.cpp:104-111

struct Camera {  // definitions.hpp:93-124: pinhole, one focal length, camera <- vehicle extrinsics
    using Ptr = std::shared_ptr<Camera>;
    Camera(double f, const Eigen::Vector2d& pp, const EigenPose& pose_cam_veh);
    Eigen::Matrix3d getIntrinsicMatrix() const;
    EigenPose getEigenPose() const;
    double focal_length;
    Eigen::Vector2d principal_point;
    Pose pose_camera_vehicle;
    Eigen::Matrix3d intrin_inv;
};

}  // namespace keyframe_bundle_adjustment
