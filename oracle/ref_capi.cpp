// ref_capi.cpp — C entry points around the REFERENCE's own base::CollisionAndDistance, compiled from
// /root/reference/src/state_spaces/real_vector_space/CollisionAndDistance.cpp (unmodified, where it lies) against
// oracle/eigen_shim.  TEST INFRASTRUCTURE ONLY: used to pin the oracle restatement and to generate tests/golden/.
// Output goes to oracle/_ref/ (git-ignored). Row layout of `out` matches oracle/rvs_oracle_capi.cpp:
// (d, has_points, P1xyz, P2xyz).
#include <cstdint>

#include "CollisionAndDistance.h"

namespace {
using base::CollisionAndDistance;
Eigen::Vector3f v3(const float *p) { return Eigen::Vector3f(p[0], p[1], p[2]); }
Eigen::VectorXf obs6(const float *p) { Eigen::VectorXf o(6); for (int i = 0; i < 6; i++) o(i) = p[i]; return o; }
void put(const std::tuple<float, std::shared_ptr<Eigen::MatrixXf>> &r, float *out) {
    const auto &pts = std::get<1>(r);
    out[0] = std::get<0>(r);
    out[1] = pts ? 1.f : 0.f;
    for (int k = 0; k < 6; k++) out[2 + k] = pts ? (*pts)(k % 3, k / 3) : 0.f;
}
}  // namespace

extern "C" {
void ref_set_config(float eq_thr, float max_dist) {
    RealVectorSpaceConfig::EQUALITY_THRESHOLD = eq_thr;
    RealVectorSpaceConfig::MAX_DISTANCE = max_dist;
}
int ref_seg_seg_f32(int64_t K, const float *ABCD, float *out) {
    for (int64_t i = 0; i < K; i++) {
        const float *p = ABCD + 12 * i;
        put(CollisionAndDistance::distanceLineSegToLineSeg(v3(p), v3(p + 3), v3(p + 6), v3(p + 9)), out + 8 * i);
    }
    return 0;
}
int ref_seg_point_f32(int64_t K, const float *ABC, float *out) {
    for (int64_t i = 0; i < K; i++) {
        const float *p = ABC + 9 * i;
        put(CollisionAndDistance::distanceLineSegToPoint(v3(p), v3(p + 3), v3(p + 6)), out + 8 * i);
    }
    return 0;
}
int ref_capsule_box_distance_f32(int64_t K, const float *AB, const float *radius, const float *obs, float *out) {
    for (int64_t i = 0; i < K; i++) {
        Eigen::VectorXf o = obs6(obs + 6 * i);
        put(CollisionAndDistance::distanceCapsuleToBox(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], o), out + 8 * i);
    }
    return 0;
}
int ref_capsule_box_collision_f32(int64_t K, const float *AB, const float *radius, const float *obs, uint8_t *out) {
    for (int64_t i = 0; i < K; i++) {
        Eigen::VectorXf o = obs6(obs + 6 * i);
        out[i] = CollisionAndDistance::collisionCapsuleToBox(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], o) ? 1 : 0;
    }
    return 0;
}
int ref_capsule_rect_collision_f32(int64_t K, const float *AB, const float *radius, const float *obs, const int32_t *face, uint8_t *out) {
    for (int64_t i = 0; i < K; i++) {
        Eigen::VectorXf o = obs6(obs + 6 * i);
        out[i] = CollisionAndDistance::collisionCapsuleToRectangle(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], o, size_t(face[i])) ? 1 : 0;
    }
    return 0;
}
}
