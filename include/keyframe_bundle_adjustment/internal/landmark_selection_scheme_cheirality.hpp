// internal/landmark_selection_scheme_cheirality.hpp -- reject landmarks that lie behind a camera observing them
// (reference: internal/landmark_selection_scheme_cheirality.hpp, src/landmark_selection_scheme_cheirality.cpp:22-60).
// The library default of BundleAdjusterKeyframes (cpp:116-118).  The batch form of the test runs on the device:
// kba_init_landmarks (include/kba_b200.h).
#pragma once
#include "landmark_selection_scheme_base.hpp"

namespace keyframe_bundle_adjustment {

class LandmarkRejectionSchemeCheirality : public LandmarkRejectionSchemeBase {
public:
    LandmarkRejectionSchemeCheirality() { identifier = "cheirality"; }
    std::set<LandmarkId> getSelection(const LandmarkMap& landmarks, const KeyframeMap& keyframes) const override;
    static ConstPtr createConst() { return ConstPtr(new LandmarkRejectionSchemeCheirality()); }
    static Ptr create() { return Ptr(new LandmarkRejectionSchemeCheirality()); }
};

}  // namespace keyframe_bundle_adjustment
