// Stand-in for include/configurations/RealVectorSpaceConfig.h (static mutable globals in the reference).
#pragma once
#include <cstddef>
class RealVectorSpaceConfig {
public:
    static inline float EQUALITY_THRESHOLD = 1e-6f;
    static inline size_t NUM_INTERPOLATION_VALIDITY_CHECKS = 15;
    static inline float MAX_DISTANCE = 1000.f;
};
