// matches_msg_types/tracklet.hpp -- one feature track over the frames of a Tracklets message (reference:
// matches_msg_types/include/matches_msg_types/tracklet.hpp): newest measurement first, semantic label and the front
// end's outlier flag.
#pragma once
#include <vector>

#include "feature_point.hpp"

namespace matches_msg_types {

struct Tracklet {
    std::vector<FeaturePoint> feature_points;
    unsigned long id;
    unsigned long age;
    bool is_outlier{false};
    int label{-2};
};

}  // namespace matches_msg_types
