// forwarding header: the facade keeps the reference include paths (keyframe_bundle_adjustment/keyframe.hpp)
#pragma once
#include "bundle_adjuster_keyframes.hpp"
