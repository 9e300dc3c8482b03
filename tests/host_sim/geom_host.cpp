// geom_host.cpp — TEST-ONLY: compiles the DEVICE headers (rvs_geom.cuh, rvs_fk.cuh) for the host with g++ so that the
// branch logic of the sm_100a narrow phase can be compared with the oracle on a machine without a GPU.
// This is not a product path and nothing in rpmplv2_b200/ loads it; the GPU parity tests (-m gpu) are the real gate.
#include <cmath>
#include <cstdint>
static inline int __ffs(int x) { return __builtin_ffs(x); }
#define __noinline__
#include "../../rpmplv2_b200/csrc/rvs_fk.cuh"
#include "../../rpmplv2_b200/csrc/rvs_geom.cuh"
#include "../../rpmplv2_b200/csrc/rvs_robot_host.h"

using namespace rvs;

template <class R>
static void dist_batch(int64_t K, const R *AB, const R *radius, const R *obs, R thr, R maxd, int pts, R *out) {
    for (int64_t i = 0; i < K; i++) {
        Capsule<R> cp;
        cp.set({AB[6 * i], AB[6 * i + 1], AB[6 * i + 2]}, {AB[6 * i + 3], AB[6 * i + 4], AB[6 * i + 5]}, radius[i], R(0));
        Box<R> bx = {{obs[6 * i], obs[6 * i + 1], obs[6 * i + 2]}, {obs[6 * i + 3], obs[6 * i + 4], obs[6 * i + 5]}};
        V3<R> P1 = {0, 0, 0}, P2 = {0, 0, 0};
        R d = pts ? capsule_box_distance<R, true>(cp, bx, thr, maxd, P1, P2) : capsule_box_distance<R, false>(cp, bx, thr, maxd, P1, P2);
        R *o = out + 8 * i;
        o[0] = d; o[1] = 0; o[2] = P1.x; o[3] = P1.y; o[4] = P1.z; o[5] = P2.x; o[6] = P2.y; o[7] = P2.z;
    }
}

// The pair routine of nearest_points_kernel (rvs_distance_full.cuh): classify, park the undecided pair as (plan word, candidate distance), and
// evaluate its edges LATER from the parked words alone — the candidate points are rebuilt from PairPlan::cand.
template <class R>
static void points_deferred_batch(int64_t K, const R *AB, const R *radius, const R *obs, R thr, R maxd, R *out) {
    for (int64_t i = 0; i < K; i++) {
        Capsule<R> cp;
        cp.set({AB[6 * i], AB[6 * i + 1], AB[6 * i + 2]}, {AB[6 * i + 3], AB[6 * i + 4], AB[6 * i + 5]}, radius[i], R(0));
        Box<R> bx = {{obs[6 * i], obs[6 * i + 1], obs[6 * i + 2]}, {obs[6 * i + 3], obs[6 * i + 4], obs[6 * i + 5]}};
        V3<R> P1 = {0, 0, 0}, P2 = {0, 0, 0};
        const PairPlan<R> pl = classify_pair<R, true>(cp, bx, thr, maxd, P1, P2);
        R d = pl.d;
        if (!pl.done) {
            const uint32_t plan = (pl.twelve ? 0u : pl.list) | (uint32_t(pl.count) << 24) | (pl.cand << 28);      // what the ring holds
            const R parked = pl.d;
            // ---- later ----
            const int edges = int((plan >> 24) & 15u);
            R best = parked;
            bool hit = false;
            plan_candidate_points(cp, bx, plan >> 28, P1, P2);
            PairPlan<R> q;
            q.list = plan & 0xffffffu; q.count = edges; q.twelve = edges == 12;
            for (int k = 0; k < edges; k++) {
                V3<R> q1, q2;
                const R de = plan_edge_distance<R, true>(cp, bx, plan_edge(q, k), thr, maxd, q1, q2);
                if (de <= R(0)) hit = true;
                if (de > R(0) && de < best) { best = de; P1 = q1; P2 = q2; }
            }
            d = hit ? R(0) - cp.radius : best - cp.radius;
        }
        R *o = out + 8 * i;
        o[0] = d; o[1] = pl.done ? 1 : 0; o[2] = P1.x; o[3] = P1.y; o[4] = P1.z; o[5] = P2.x; o[6] = P2.y; o[7] = P2.z;
    }
}

extern "C" {
int hs_capsule_box_points_deferred_f32(int64_t K, const float *AB, const float *radius, const float *obs, float thr, float maxd, float *out) {
    points_deferred_batch<float>(K, AB, radius, obs, thr, maxd, out);
    return 0;
}
int hs_capsule_box_points_deferred_f64(int64_t K, const double *AB, const double *radius, const double *obs, double thr, double maxd, double *out) {
    points_deferred_batch<double>(K, AB, radius, obs, thr, maxd, out);
    return 0;
}
int hs_capsule_box_distance_f32(int64_t K, const float *AB, const float *radius, const float *obs, float thr, float maxd, int pts, float *out) {
    dist_batch<float>(K, AB, radius, obs, thr, maxd, pts, out);
    return 0;
}
int hs_capsule_box_distance_f64(int64_t K, const double *AB, const double *radius, const double *obs, double thr, double maxd, int pts, double *out) {
    dist_batch<double>(K, AB, radius, obs, thr, maxd, pts, out);
    return 0;
}
int hs_capsule_box_collision_f32(int64_t K, const float *AB, const float *radius, const float *obs, float thr, float maxd, uint8_t *out) {
    for (int64_t i = 0; i < K; i++) {
        Capsule<float> cp;
        cp.set({AB[6 * i], AB[6 * i + 1], AB[6 * i + 2]}, {AB[6 * i + 3], AB[6 * i + 4], AB[6 * i + 5]}, radius[i],
               float(double(radius[i]) * std::sqrt(3.0) / 3));
        Box<float> bx = {{obs[6 * i], obs[6 * i + 1], obs[6 * i + 2]}, {obs[6 * i + 3], obs[6 * i + 4], obs[6 * i + 5]}};
        out[i] = capsule_hits_box<float>(cp, bx, thr, maxd) ? 1 : 0;
    }
    return 0;
}
int hs_skeleton(const rvs_robot_desc *rd, const float *q, int64_t N, int fk_f64, float *sk32, double *sk64) {
    RobotDev rb;
    if (fill_robot_dev(rd, &rb)) return 1;
    const int n = rb.n;
    for (int64_t i = 0; i < N; i++) {
        if (sk32) { if (fk_f64) fk_skeleton<double, float>(&rb, q + i * n, sk32 + i * 3 * (n + 1)); else fk_skeleton<float, float>(&rb, q + i * n, sk32 + i * 3 * (n + 1)); }
        if (sk64) fk_skeleton<double, double>(&rb, q + i * n, sk64 + i * 3 * (n + 1));
    }
    return 0;
}
}
