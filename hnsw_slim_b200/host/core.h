// Global knobs of the host layer: same names, defaults and meaning as the reference's
// include/core.h:30-38 (set from the command line in main.cc, copied by SolveStrategy's ctor).
#pragma once
#include <cstddef>
#include <string>

inline size_t K = 10;                          // top-k
inline size_t M = 32;                          // neighbours per node
inline size_t M0 = 32;                         // parsed, unused: maxM0 = 2 M (slim.h:115)
inline size_t EF_CONSTRUCTION = 1024;
inline size_t EF_SEARCH = 64;
inline std::string BRANCHING_FACTOR = "4";
inline size_t THRESHOLD_LEVEL = 0;
