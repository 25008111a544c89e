// internal/landmark_selection_scheme_base.hpp -- the three kinds of landmark schemes the selector chains (reference:
// internal/landmark_selection_scheme_base.hpp): rejection (never take), selection (always take), sparsification (thin
// out what is left).  A scheme maps (landmarks, keyframes) to the set of landmark ids it lets through.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>

#include "../keyframe.hpp"
#include "definitions.hpp"

namespace keyframe_bundle_adjustment {

class LandmarkSchemeBase {
public:
    using LandmarkMap = std::map<LandmarkId, Landmark::ConstPtr>;
    using KeyframeMap = std::map<KeyframeId, Keyframe::ConstPtr>;
    virtual ~LandmarkSchemeBase() = default;
    virtual std::set<LandmarkId> getSelection(const LandmarkMap& landmarks, const KeyframeMap& keyframes) const = 0;
    std::string identifier = "";
};
class LandmarkRejectionSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkRejectionSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkRejectionSchemeBase>;
};
class LandmarkSelectionSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkSelectionSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkSelectionSchemeBase>;
};
class LandmarkSparsificationSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkSparsificationSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkSparsificationSchemeBase>;
};

}  // namespace keyframe_bundle_adjustment
