// internal/landmark_categorization_interface.hpp -- schemes that also say which depth band a landmark fell into
// (reference: internal/landmark_categorization_interface.hpp:17-29).  LandmarkSelector keeps the categories of the last
// categorising scheme it ran (landmark_selector.hpp:79-99).
#pragma once
#include <map>

#include "../keyframe.hpp"
#include "definitions.hpp"

namespace keyframe_bundle_adjustment {

struct LandmarkCategorizatonInterface {  // (sic) the reference's spelling is part of the API
    enum class Category { NearField, MiddleField, FarField };
    virtual ~LandmarkCategorizatonInterface() = default;
    virtual std::map<LandmarkId, Category> getCategorizedSelection(
        const std::map<LandmarkId, Landmark::ConstPtr>& landmarks,
        const std::map<KeyframeId, Keyframe::ConstPtr>& keyframes) const = 0;
};

}  // namespace keyframe_bundle_adjustment
