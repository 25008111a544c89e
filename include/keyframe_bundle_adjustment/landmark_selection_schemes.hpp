// landmark_selection_schemes.hpp -- every landmark scheme of this interface (reference: landmark_selection_schemes.hpp).
// Not provided: the random, observability and dimension-plausibility schemes (unused by the production node,
// mono_lidar.cpp:383-429; SURVEY.md section 8 marks them out of scope).
#pragma once
#include "internal/landmark_selection_scheme_add_depth.hpp"
#include "internal/landmark_selection_scheme_base.hpp"
#include "internal/landmark_selection_scheme_cheirality.hpp"
#include "internal/landmark_selection_scheme_voxel.hpp"
