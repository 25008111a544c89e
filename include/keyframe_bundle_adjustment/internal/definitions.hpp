// forwarding header (reference: keyframe_bundle_adjustment/internal/definitions.hpp)
#pragma once
#include "../bundle_adjuster_keyframes.hpp"
