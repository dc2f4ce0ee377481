// oracle/shims — minimal stand-ins so that two reference files compile where they lie
// (COLMAP itself is not installable offline).  Only the typedefs graph.h needs.
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
namespace colmap {
typedef uint32_t camera_t;
typedef uint32_t image_t;
typedef uint64_t image_pair_t;
typedef uint32_t point2D_t;
typedef uint64_t point3D_t;
}  // namespace colmap
