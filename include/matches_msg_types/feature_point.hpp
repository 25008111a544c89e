// matches_msg_types/feature_point.hpp -- one tracked image measurement (reference: matches_msg_types/include/
// matches_msg_types/feature_point.hpp:4-36): pixel (u, v) and, when a lidar depth was attached by the front end, the depth
// d along the optical axis; d = -1 means "no depth".  Single precision like the reference's message type.
#pragma once
#include "keyframe_bundle_adjustment/internal/mini_eigen.hpp"

namespace matches_msg_types {

struct FeaturePoint {
    FeaturePoint() {}
    FeaturePoint(Eigen::Vector2d p) : u(float(p[0])), v(float(p[1])), d(-1) {}
    FeaturePoint(Eigen::Vector3d p) : u(float(p[0])), v(float(p[1])), d(float(p[2])) {}
    FeaturePoint(float u, float v) : u(u), v(v), d(-1) {}
    FeaturePoint(float u, float v, float d) : u(u), v(v), d(d) {}
    Eigen::Vector2d toEigen2d() const { return Eigen::Vector2d(u, v); }
    Eigen::Vector3d toEigen3d() const { return Eigen::Vector3d(u, v, d); }
    float u, v, d;
};

}  // namespace matches_msg_types
