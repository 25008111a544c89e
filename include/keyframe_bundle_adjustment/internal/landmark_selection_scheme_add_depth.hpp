// internal/landmark_selection_scheme_add_depth.hpp -- "always take" scheme that guarantees a number of landmarks of a
// given kind on chosen keyframes (reference: internal/landmark_selection_scheme_add_depth.hpp:24-85,
// src/landmark_selection_scheme_add_depth.cpp:16-75).  The production node uses it to keep the 50 nearest ground-plane
// landmarks of every window keyframe (mono_lidar.cpp:413-429).
#pragma once
#include <functional>
#include <tuple>
#include <vector>

#include "landmark_selection_scheme_base.hpp"

namespace keyframe_bundle_adjustment {

class LandmarkSelectionSchemeAddDepth : public LandmarkSelectionSchemeBase {
public:
    using FrameIndex = int;       // 0 = oldest active keyframe
    using NumberLandmarks = int;  // how many landmarks to guarantee on that keyframe
    using Comparator = std::function<bool(const Landmark::ConstPtr&)>;                    // eligible?
    using Sorter = std::function<float(const Measurement&, const Eigen::Vector3d&)>;      // smaller = taken first
    struct Parameters {
        std::vector<std::tuple<FrameIndex, NumberLandmarks, Comparator, Sorter>> params_per_keyframe;
    };
    LandmarkSelectionSchemeAddDepth(Parameters p) : params_(p) { identifier = "add depth"; }
    std::set<LandmarkId> getSelection(const LandmarkMap& landmarks, const KeyframeMap& keyframes) const override;
    static ConstPtr createConst(Parameters p) { return ConstPtr(new LandmarkSelectionSchemeAddDepth(p)); }
    static Ptr create(Parameters p) { return Ptr(new LandmarkSelectionSchemeAddDepth(p)); }
    Parameters params_;
};

}  // namespace keyframe_bundle_adjustment
