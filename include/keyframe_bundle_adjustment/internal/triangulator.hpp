// internal/triangulator.hpp -- landmark from viewing rays (reference: internal/triangulator.hpp:51-75): the point closest,
// in the least-squares sense, to all rays (origin <- camera pose, unit direction in the camera frame).
#pragma once
#include <utility>
#include <vector>

#include "definitions.hpp"

namespace keyframe_bundle_adjustment {

Eigen::Vector3d triangulate_rays(const std::vector<std::pair<EigenPose, Eigen::Vector3d>>& poses_rays);

template <typename T = double>
class Triangulator {  // the reference wraps the same routine in a class template (bundle_adjuster_keyframes.hpp:332)
public:
    using PoseAndRay = std::pair<EigenPose, Eigen::Vector3d>;
    Eigen::Vector3d triangulate_rays(const std::vector<PoseAndRay>& poses_rays) const {
        return keyframe_bundle_adjustment::triangulate_rays(poses_rays);
    }
};

}  // namespace keyframe_bundle_adjustment
