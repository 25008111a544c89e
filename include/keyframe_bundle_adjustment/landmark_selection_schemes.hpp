// forwarding header: the facade keeps the reference include paths (keyframe_bundle_adjustment/landmark_selection_schemes.hpp)
#pragma once
#include "bundle_adjuster_keyframes.hpp"
