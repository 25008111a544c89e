// forwarding header (reference: matches_msg_types/tracklets.hpp)
#pragma once
#include "keyframe_bundle_adjustment/bundle_adjuster_keyframes.hpp"
