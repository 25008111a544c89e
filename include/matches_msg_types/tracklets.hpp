// matches_msg_types/tracklets.hpp -- all tracks of one message plus the time stamps their feature_points refer to
// (reference: matches_msg_types/include/matches_msg_types/tracklets.hpp).
#pragma once
#include <cstdint>
#include <vector>

#include "tracklet.hpp"

namespace matches_msg_types {

using TimestampNSec = uint64_t;  // unix time in nanoseconds; a keyframe's id IS its time stamp

struct Tracklets {
    std::vector<TimestampNSec> stamps;
    std::vector<Tracklet> tracks;
};

}  // namespace matches_msg_types
