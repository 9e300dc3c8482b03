// rvs_kernels.cuh — sm_100a kernels of the RealVectorSpaceCUDA engine (included by rvs_variant.cu, which is
// compiled once per arithmetic variant: -fmad=false for REF32/F64, -fmad=true for FAST32).
//
// Execution model of the batch kernels: ONE THREAD PER CONFIGURATION.
//   1. forward kinematics of the serial chain in registers (FP64 for REF32/F64, FP32 for FAST32), streamed joint
//      by joint into the capsule skeleton (per-thread local array, touched once per link per obstacle tile);
//   2. the obstacle table is staged tile by tile (512 records of 32 B) into shared memory with coalesced 16-byte
//      loads; inside a tile the loop is link-major / obstacle-minor so the per-link capsule constants stay in
//      registers and every obstacle read is a warp-wide broadcast LDS.128;
//   3. the narrow phase is rvs_geom.cuh.
// Results are written straight to HBM (4..28 B per configuration in MIN/PROFILE mode).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "rvs_fk.cuh"
#include "rvs_geom.cuh"
#include "rvs_types.h"

namespace rvs {

constexpr int kBlock = 128;

template <class R> __device__ __forceinline__ Box<R> load_box(const ObstacleRec &o) {
    return {{R(o.lo[0]), R(o.lo[1]), R(o.lo[2])}, {R(o.hi[0]), R(o.hi[1]), R(o.hi[2])}};
}

// cooperative, coalesced 16-byte staging of one obstacle tile
// The trip count is the same for every thread (the body is predicated): a thread-dependent trip count lets a warp reach
// the following __syncthreads() in two diverged fragments, and the late fragment then waits at the barrier until the
// other warps of the block exit (seen in profiles/r1b: 21 % of all stall samples).
__device__ __forceinline__ void stage_tile(ObstacleRec *tile, const ObstacleRec *__restrict__ src, int count) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(tile);
    const int total = 2 * count;
    for (int k0 = 0; k0 < total; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        if (k < total) d4[k] = __ldg(s4 + k);
    }
    __syncwarp();
}

template <class R> struct RobotScalars;
template <> struct RobotScalars<float> {
    static __device__ __forceinline__ float radius(const RobotDev *rb, int i) { return rb->radius[i]; }
    static __device__ __forceinline__ float radius_in(const RobotDev *rb, int i) { return rb->radius_in[i]; }
};
template <> struct RobotScalars<double> {
    static __device__ __forceinline__ double radius(const RobotDev *rb, int i) { return rb->radius_d[i]; }
    static __device__ __forceinline__ double radius_in(const RobotDev *rb, int i) { return rb->radius_in_d[i]; }
};

// ---- computeDistance (RealVectorSpace.cpp:318-392) ----------------------------------------------------------------
// MODE: 0 MIN, 1 PROFILE, 2 FULL.  F = FK scalar, R = geometry scalar.
#ifndef RVS_SEQ_MINBLOCKS
#define RVS_SEQ_MINBLOCKS 1
#endif
template <class F, class R, int MODE>
__global__ void __launch_bounds__(kBlock, sizeof(R) == 4 ? RVS_SEQ_MINBLOCKS : 1) distance_kernel(const KArgs a) {
    __shared__ ObstacleRec tile[kTileObstacles];
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n;
    const int64_t slot_k = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    const bool live = slot_k < a.N;
    // a.perm (rvs_sort.cuh): the lanes of a warp walk the same (link, obstacle) pairs in lockstep, so arms that are alike take the same
    // branches of the reference's case tree
    const int64_t cfg = live && a.perm ? int64_t(__ldg(a.perm + slot_k)) : slot_k;
    const R eq_thr = R(a.eq_thr), max_dist = R(a.max_dist);

    R sk[3 * (kMaxDofs + 1)];
    R prof[kMaxDofs];
    if (live) fk_skeleton<F, R>(rb, a.q + cfg * L, sk);
    for (int i = 0; i < L; i++) prof[i] = max_dist;
    int limit = live ? L : 0;      // links >= limit are no longer evaluated (first collision semantics, :371-378)
    R hit_d = 0;
    int hit_j = 0;
    R *nearest = (MODE == RVS_DIST_FULL && live) ? static_cast<R *>(a.nearest) + size_t(cfg) * size_t(a.M) * size_t(L) * 6 : nullptr;
    const int ground_included = rb->ground_included;

    for (int base = 0; base < a.M; base += kTileObstacles) {
        const int count = min(kTileObstacles, a.M - base);
        __syncthreads();
        stage_tile(tile, a.obstacles + base, count);
        __syncthreads();
#pragma unroll 1
        for (int i = 0; i < limit; i++) {
            Capsule<R> cp;
            cp.set({sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]},
                   RobotScalars<R>::radius(rb, i), R(0));
            R best = prof[i];
#pragma unroll 1
            for (int j = 0; j < count; j++) {
                const ObstacleRec &o = tile[j];
                R d;
                V3<R> P1 = {0, 0, 0}, P2 = {0, 0, 0};
                if ((o.flags & 1u) && i < ground_included) {           // ground rule (:336-341)
                    d = max_dist;
                    P2.z = -max_dist;
                } else {
                    const Box<R> bx = load_box<R>(o);
                    d = capsule_box_distance<R, MODE == RVS_DIST_FULL>(cp, bx, eq_thr, max_dist, P1, P2);
                }
                if (d > R(o.tol)) d = max_dist;                        // min_dist_tol (:367-368)
                if (d <= R(0)) { limit = i; hit_d = d; hit_j = base + j; break; }        // first collision of the lowest link so far
                if (d < best) best = d;
                if (MODE == RVS_DIST_FULL) {
                    R *dst = nearest + (size_t(base + j) * L + i) * 6;
                    dst[0] = P1.x; dst[1] = P1.y; dst[2] = P1.z; dst[3] = P2.x; dst[4] = P2.y; dst[5] = P2.z;
                }
            }
            prof[i] = best;
        }
    }
    // Work accounting in the REFERENCE's loop order (link-major, stop at the first collision): the number of
    // distanceCapsuleToBox calls RealVectorSpace::computeDistance would have made for this configuration.
    if (a.pair_count) {
        unsigned visited = 0;
        if (live) {
            for (int i = 0; i < limit; i++) visited += unsigned(a.M - (i < ground_included ? a.num_ground : 0));
            if (limit < L) {
                unsigned upto = unsigned(hit_j + 1);
                if (limit < ground_included && a.num_ground > 0)
                    for (int j = 0; j <= hit_j; j++) upto -= (a.obstacles[j].flags & 1u);
                visited += upto;
            }
        }
        for (int off = 16; off > 0; off >>= 1) visited += __shfl_down_sync(0xffffffffu, visited, off);
        if ((threadIdx.x & 31) == 0 && visited) atomicAdd(a.pair_count, (unsigned long long)visited);
    }
    if (!live) return;
    const bool collided = limit < L;
    R d_c = max_dist;
    for (int i = 0; i < L; i++) if (prof[i] < d_c) d_c = prof[i];
    static_cast<R *>(a.d_c)[cfg] = collided ? R(0) : d_c;
    if (MODE >= RVS_DIST_PROFILE) {
        R *out = static_cast<R *>(a.profile) + cfg * L;
        for (int i = 0; i < L; i++) out[i] = !collided ? prof[i] : (i < limit ? prof[i] : (i == limit ? hit_d : R(0)));
    }
}

// ---- isValid(q) (RealVectorSpace.cpp:277-309) ------------------------------------------------------------------------
template <class F, class R>
__device__ __forceinline__ bool config_is_valid(const KArgs &a, const RobotDev *__restrict__ rb, ObstacleRec *tile, const R *sk, bool live) {
    const int L = rb->n;
    const R eq_thr = R(a.eq_thr), max_dist = R(a.max_dist);
    const int ground_included = rb->ground_included;
    bool valid = true;
    for (int base = 0; base < a.M; base += kTileObstacles) {
        const int count = min(kTileObstacles, a.M - base);
        __syncthreads();
        stage_tile(tile, a.obstacles + base, count);
        __syncthreads();
        if (!live || !valid) continue;
#pragma unroll 1
        for (int i = 0; i < L && valid; i++) {
            Capsule<R> cp;
            cp.set({sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]},
                   RobotScalars<R>::radius(rb, i), RobotScalars<R>::radius_in(rb, i));
#pragma unroll 1
            for (int j = 0; j < count; j++) {
                const ObstacleRec &o = tile[j];
                if ((o.flags & 1u) && i < ground_included) continue;
                if (capsule_hits_box<R>(cp, load_box<R>(o), eq_thr, max_dist)) { valid = false; break; }
            }
        }
    }
    return valid;
}

template <class F, class R>
__global__ void __launch_bounds__(kBlock) is_valid_kernel(const KArgs a) {
    __shared__ ObstacleRec tile[kTileObstacles];
    const RobotDev *__restrict__ rb = a.robot;
    const int64_t cfg = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    const bool live = cfg < a.N;
    R sk[3 * (kMaxDofs + 1)];
    if (live) fk_skeleton<F, R>(rb, a.q + cfg * rb->n, sk);
    const bool valid = config_is_valid<F, R>(a, rb, tile, sk, live);
    if (live) a.verdict[cfg] = valid ? 1 : 0;
}

// (q1 - q2).norm() of a dynamic VectorXf: Eigen's SSE2 linear-vectorised sum order (see oracle/rvs_oracle.hpp coord_norm)
__device__ __forceinline__ float coord_norm(const float *__restrict__ a, const float *__restrict__ b, int n) {
    const int aligned = (n / 4) * 4, aligned2 = (n / 8) * 8;
    float res;
    auto sq = [&](int i) { const float d = a[i] - b[i]; return d * d; };
    if (aligned) {
        float p0[4], p1[4];
#pragma unroll
        for (int l = 0; l < 4; l++) p0[l] = sq(l);
        if (aligned > 4) {
#pragma unroll
            for (int l = 0; l < 4; l++) p1[l] = sq(4 + l);
            for (int idx = 8; idx < aligned2; idx += 8) {
#pragma unroll
                for (int l = 0; l < 4; l++) { p0[l] = p0[l] + sq(idx + l); p1[l] = p1[l] + sq(idx + 4 + l); }
            }
#pragma unroll
            for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
            if (aligned > aligned2) {
#pragma unroll
                for (int l = 0; l < 4; l++) p0[l] = p0[l] + sq(aligned2 + l);
            }
        }
        res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
        for (int i = aligned; i < n; i++) res = res + sq(i);
    } else {
        res = sq(0);
        for (int i = 1; i < n; i++) res = res + sq(i);
    }
    return sqrtf(res);
}

// ---- isValid(q1, q2) (RealVectorSpace.cpp:261-275): one thread per (edge, check); verdict = AND over the checks -------
// verdict[] must be pre-set to 1; threads that find a collision store 0 (idempotent).
template <class F, class R>
__global__ void __launch_bounds__(kBlock) edge_valid_kernel(const KArgs a) {
    __shared__ ObstacleRec tile[kTileObstacles];
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t tid = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    const int64_t edge = tid / a.num_checks;
    const int kidx = int(tid - edge * a.num_checks);
    const bool live = edge < a.N;
    R sk[3 * (kMaxDofs + 1)];
    if (live) {
        const float *q1 = a.q + edge * n, *q2 = a.q2 + edge * n;
        const float dist = coord_norm(q1, q2, n);
        const float k = float(a.num_checks - kidx);                 // k = N .. 1
        const float step = k / float(a.num_checks) * dist;
        float qn[kMaxDofs];
        if (dist > 0.f) {
            const float f = step / dist;
            for (int i = 0; i < n; i++) qn[i] = q1[i] + (q2[i] - q1[i]) * f;
        } else
            for (int i = 0; i < n; i++) qn[i] = q2[i];
        fk_skeleton<F, R>(rb, qn, sk);
    }
    const bool valid = config_is_valid<F, R>(a, rb, tile, sk, live);
    if (live && !valid) a.verdict[edge] = 0;
}

// ---- computeDistanceUnderestimation (RealVectorSpace.cpp:402-446) ----------------------------------------------------------
// planes: [M][L][6] = (R, O) per (obstacle, link); prof[L] receives the per-link distance-to-planes (unspecified when
// rejected). Returns d_c; rejected = a plane distance < 0 (the reference returns 0 and leaves the state untouched, :428-429).
template <class R>
__device__ __forceinline__ R underestimation_of(const RobotDev *__restrict__ rb, int L, int M, const R *sk, const float *__restrict__ planes,
                                                R max_dist, R *prof, bool &rejected) {
    R d_c = max_dist;
    rejected = false;
#pragma unroll 1
    for (int i = 0; i < L && !rejected; i++) {
        const V3<R> A = {sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, B = {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]};
        const R radius = RobotScalars<R>::radius(rb, i);
        R best = max_dist;
#pragma unroll 1
        for (int j = 0; j < M; j++) {
            const float2 *src = reinterpret_cast<const float2 *>(planes + (size_t(j) * L + i) * 6);
            const float2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
            const V3<R> Rp = {R(w0.x), R(w0.y), R(w1.x)}, Op = {R(w1.y), R(w2.x), R(w2.y)};
            if (rsqrt_of(dotF(Op, Op)) == max_dist) continue;         // ground placeholder (:420-422)
            V3<R> nrm = sub(Rp, Op);
            const R z = dotF(nrm, nrm);
            if (z > R(0)) { const R len = rsqrt_of(z); nrm = {nrm.x / len, nrm.y / len, nrm.z / len}; }
            const V3<R> ao = sub(A, Op), bo = sub(B, Op);
            const R da = (ao.x * nrm.x + ao.y * nrm.y) + ao.z * nrm.z;   // dynamic-order dot (skeleton column lhs)
            const R db = (bo.x * nrm.x + bo.y * nrm.y) + bo.z * nrm.z;
            const R d = (db < da ? db : da) - radius;
            if (d < R(0)) { rejected = true; break; }
            if (d < best) best = d;
        }
        prof[i] = best;
        if (best < d_c) d_c = best;
    }
    return rejected ? R(0) : d_c;
}

template <class F, class R>
__global__ void __launch_bounds__(kBlock) underestimation_kernel(const KArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n;
    const int64_t cfg = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (cfg >= a.N) return;
    R sk[3 * (kMaxDofs + 1)];
    R prof[kMaxDofs];
    fk_skeleton<F, R>(rb, a.q + cfg * L, sk);
    const int root = a.plane_idx ? a.plane_idx[cfg] : 0;
    bool rejected;
    const R d_c = underestimation_of<R>(rb, L, a.M, sk, a.planes + size_t(root) * size_t(a.M) * size_t(L) * 6, R(a.max_dist), prof, rejected);
    static_cast<R *>(a.d_c)[cfg] = d_c;
    if (a.status) a.status[cfg] = rejected ? 1 : 0;
    if (a.profile) {
        R *out = static_cast<R *>(a.profile) + cfg * L;
        for (int i = 0; i < L; i++) out[i] = rejected ? R(0) : prof[i];
    }
}

// ---- skeleton / enclosing radii / xArm6 capsule self-collision ---------------------------------------------------------
template <class F, class R>
__global__ void __launch_bounds__(kBlock) skeleton_kernel(const KArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t cfg = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (cfg >= a.N) return;
    R sk[3 * (kMaxDofs + 1)];
    fk_skeleton<F, R>(rb, a.q + cfg * n, sk);
    R *out = static_cast<R *>(a.d_c) + cfg * 3 * (n + 1);
    for (int k = 0; k < 3 * (n + 1); k++) out[k] = sk[k];
}

// computeEnclosingRadii: xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30. out: n x (n+1) row-major.
// row i of AbstractRobot::computeEnclosingRadii (xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30): out[j], j = 0..n
template <class R>
__device__ __forceinline__ void enclosing_radii_row(const RobotDev *__restrict__ rb, const R *sk, int i, R *out) {
    const int n = rb->n;
    for (int j = 0; j <= n; j++) out[j] = R(0);
    auto P = [&](int k) { return V3<R>{sk[3 * k], sk[3 * k + 1], sk[3 * k + 2]}; };
    auto norm_xy = [&](int k) { return rsqrt_of(sk[3 * k] * sk[3 * k] + sk[3 * k + 1] * sk[3 * k + 1]); };
    auto dist = [&](const V3<R> &u, const V3<R> &v) { return rsqrt_of(sqD(u.x - v.x, u.y - v.y, u.z - v.z)); };
    if (rb->kind == RVS_ROBOT_XARM6) {
        if (i >= n - 1) return;
        for (int j = i + 1; j <= n; j++) {
            const R rn = R(rb->radius_new[j - 1]);
            R len;
            if (i == 0) { if (j < 2) continue; len = norm_xy(j); }
            else if (i == 3) {
                if (j < 5) continue;
                const V3<R> AB = sub(P(4), P(3)), v = sub(P(j), P(3));
                const R t = ((v.x * AB.x + v.y * AB.y) + v.z * AB.z) / dotF(AB, AB);
                len = dist(P(j), lerp_from(P(3), t, AB));
            } else len = dist(P(j), P(i));
            const R cand = len + rn, prev = out[j - 1];
            out[j] = prev < cand ? cand : prev;
        }
    } else {
        for (int j = i + 1; j <= n; j++) {
            const R len = (rb->kind == RVS_ROBOT_SPATIAL && i == 0) ? norm_xy(j) : dist(P(j), P(i));
            const R cand = len + R(rb->radius[j - 1]), prev = out[j - 1];
            out[j] = prev < cand ? cand : prev;
        }
    }
}

template <class R>
__device__ __forceinline__ void enclosing_radii_of(const RobotDev *__restrict__ rb, const R *sk, R *out) {
    const int n = rb->n, W = n + 1;
    for (int i = 0; i < n; i++) enclosing_radii_row<R>(rb, sk, i, out + i * W);
}

template <class F, class R>
__global__ void __launch_bounds__(kBlock) enclosing_radii_kernel(const KArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t cfg = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (cfg >= a.N) return;
    R sk[3 * (kMaxDofs + 1)];
    fk_skeleton<F, R>(rb, a.q + cfg * n, sk);
    enclosing_radii_of<R>(rb, sk, static_cast<R *>(a.d_c) + cfg * n * (n + 1));
}

// generic segment-segment distance (distanceLineSegToLineSeg, CollisionAndDistance.cpp:217-320), distance only;
// used by the xArm6 capsule self-collision pre-filter (xArm6.cpp:417-424) where neither segment is axis aligned.
template <class R>
__device__ __forceinline__ R seg_seg_distance(const V3<R> &A, const V3<R> &B, const V3<R> &C, const V3<R> &D, R eq_thr, R max_dist) {
    const V3<R> AB = sub(B, A), CD = sub(D, C), DC = sub(C, D), CA = sub(A, C), BA = sub(A, B);
    const R alpha1 = dotF(AB, AB), alpha2 = dotF(AB, CD), beta1 = dotF(DC, AB), beta2 = dotF(DC, CD);
    const R gamma1 = dotF(CA, BA), gamma2 = dotF(CA, DC);
    const R s = (alpha1 * gamma2 - alpha2 * gamma1) / (alpha1 * beta2 - alpha2 * beta1);
    const R t = (gamma1 - beta1 * s) / alpha1;
    if (t > R(0) && t < R(1) && s > R(0) && s < R(1)) {
        const V3<R> p1 = lerp_from(A, t, AB), p2 = lerp_from(C, s, CD);
        const R d = rsqrt_of(sqD(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z));
        return d < eq_thr ? R(0) : d;
    }
    const R alpha3 = dotF(DC, DC);
    const R opt[4] = {dotF(CA, BA) / alpha1, dotF(sub(A, D), BA) / alpha1, dotF(CA, CD) / alpha3, dotF(sub(B, C), CD) / alpha3};
    R best = max_dist;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        V3<R> p1, p2;
        if (i < 2) {
            p1 = opt[i] < R(0) ? A : (opt[i] > R(1) ? B : lerp_from(A, opt[i], AB));
            p2 = i == 0 ? C : D;
        } else {
            p1 = i == 2 ? A : B;
            p2 = opt[i] < R(0) ? C : (opt[i] > R(1) ? D : lerp_from(C, opt[i], CD));
        }
        const R d = rsqrt_of(sqD(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z));
        if (d < best) best = d;
    }
    return best;
}

template <class F, class R>
__global__ void __launch_bounds__(kBlock) self_collision_kernel(const KArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t cfg = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (cfg >= a.N) return;
    uint8_t mask = 0;
    if (rb->kind == RVS_ROBOT_XARM6) {
        R sk[3 * (kMaxDofs + 1)];
        fk_skeleton<F, R>(rb, a.q + cfg * n, sk);
        auto P = [&](int k) { return V3<R>{sk[3 * k], sk[3 * k + 1], sk[3 * k + 2]}; };
        const int pa[4] = {0, 1, 0, 1}, pb[4] = {5, 5, 4, 4};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const R d = seg_seg_distance<R>(P(pa[k]), P(pa[k] + 1), P(pb[k]), P(pb[k] + 1), R(a.eq_thr), R(a.max_dist)) -
                        RobotScalars<R>::radius(rb, pa[k]) - RobotScalars<R>::radius(rb, pb[k]);
            if (d < R(0)) mask |= uint8_t(1u << k);
        }
    }
    if (a.mode == 1) { if (mask) a.verdict[cfg] = 0; }          // rvs_check_states: clear the validity verdict of a self-colliding configuration
    else a.verdict[cfg] = mask;
}


// ---- batched primitives (CollisionAndDistance::distanceCapsuleToBox / collisionCapsuleToBox as public statics) ------------
// a.q = AB [K x 6], a.q2 = radius [K], a.planes = obs [K x 6]; a.d_c = out [K x 8] (distance) or a.verdict (collision).
template <class F, class R>
__global__ void __launch_bounds__(kBlock) primitive_kernel(const KArgs a) {
    const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (i >= a.N) return;
    const float *ab = a.q + 6 * i, *ob = a.planes + 6 * i;
    const R radius = R(a.q2[i]);
    Capsule<R> cp;
    cp.set({R(ab[0]), R(ab[1]), R(ab[2])}, {R(ab[3]), R(ab[4]), R(ab[5])}, radius, R(double(radius) * sqrt(3.0) / 3));
    const Box<R> bx = {{R(ob[0]), R(ob[1]), R(ob[2])}, {R(ob[3]), R(ob[4]), R(ob[5])}};
    if (a.verdict) {
        a.verdict[i] = capsule_hits_box<R>(cp, bx, R(a.eq_thr), R(a.max_dist)) ? 1 : 0;
    } else {
        V3<R> P1 = {0, 0, 0}, P2 = {0, 0, 0};
        const R d = capsule_box_distance<R, true>(cp, bx, R(a.eq_thr), R(a.max_dist), P1, P2);
        // the reference returns a null point pair exactly when the result is 0 (endpoint inside) or the snap paths
        // produced d_c = 0 before the radius was subtracted; report has_points = (d > -radius-ish) conservatively as d > 0
        float *o = static_cast<float *>(a.d_c) + 8 * i;
        o[0] = float(d); o[1] = d > R(0) ? 1.f : 0.f;
        o[2] = float(P1.x); o[3] = float(P1.y); o[4] = float(P1.z); o[5] = float(P2.x); o[6] = float(P2.y); o[7] = float(P2.z);
    }
}


// ---- generalized-bur spines (RBTConnect::extendSpine, RBTConnect.cpp:85-150; RGBTConnect::extendGenSpine, RGBTConnect.cpp:73-92) ----
// One thread per target configuration q_e; all spines start at the same root (its profile, d_c and nearest-point planes come
// from one FULL distance query). Self-collision truncation (RBTConnect.cpp:129-131) is applied through self_collision_edge_of()
// when rvs_set_self_collision enabled it: the capsule stage of xArm6::checkSelfCollision(q_temp, q_new); `false` for every other robot.

// a.dot(b) of dynamic Eigen vectors: SSE2 linear-vectorised reduction order (packets of 4, two accumulators, predux =
// (l0+l2)+(l1+l3), scalar tail) — the same assumption as coord_norm above. a is read with a stride (a matrix column).
__device__ __forceinline__ float dot_dyn(const float *a, int stride, const float *b, int n) {
    const int aligned = (n / 4) * 4, aligned2 = (n / 8) * 8;
    auto pr = [&](int i) { return a[i * stride] * b[i]; };
    float res;
    if (aligned) {
        float p0[4], p1[4];
#pragma unroll
        for (int l = 0; l < 4; l++) p0[l] = pr(l);
        if (aligned > 4) {
#pragma unroll
            for (int l = 0; l < 4; l++) p1[l] = pr(4 + l);
            for (int idx = 8; idx < aligned2; idx += 8) {
#pragma unroll
                for (int l = 0; l < 4; l++) { p0[l] = p0[l] + pr(idx + l); p1[l] = p1[l] + pr(idx + 4 + l); }
            }
#pragma unroll
            for (int l = 0; l < 4; l++) p0[l] = p0[l] + p1[l];
            if (aligned > aligned2) {
#pragma unroll
                for (int l = 0; l < 4; l++) p0[l] = p0[l] + pr(aligned2 + l);
            }
        }
        res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
        for (int i = aligned; i < n; i++) res = res + pr(i);
    } else {
        res = pr(0);
        for (int i = 1; i < n; i++) res = res + pr(i);
    }
    return res;
}

// xArm6::computeCapsulesDistance (xArm6.cpp:417-424) on a packed skeleton
__device__ __forceinline__ float capsules_distance(const RobotDev *__restrict__ rb, const float *sk, int a, int b, float eq_thr, float max_dist) {
    auto P = [&](int k) { return V3<float>{sk[3 * k], sk[3 * k + 1], sk[3 * k + 2]}; };
    return seg_seg_distance<float>(P(a), P(a + 1), P(b), P(b + 1), eq_thr, max_dist) - rb->radius[a] - rb->radius[b];
}

// The four bubble steps of one iteration of xArm6::checkSelfCollision(q1, q2) (xArm6.cpp:303-314) from the skeleton and the
// enclosing radii of q1_temp and delta = |q2 - q1_temp|; returns step = min(min(step04, step05), min(step14, step15)) with the
// reference's std::min / std::max ((b < a) ? b : a, (a < b) ? b : a: inf / NaN of a zero-length motion propagate as there).
__device__ __forceinline__ float self_collision_steps(const RobotDev *__restrict__ rb, const float *sk, const float *Rm, const float *delta, float eq_thr,
                                                      float max_dist, float &step04, float &step05, float &step14, float &step15) {
    constexpr int n = 6, W = 7;
    const float phi2 = dot_dyn(Rm + 2, W, delta, n), phi4 = dot_dyn(Rm + 4, W, delta, n), phi5 = dot_dyn(Rm + 5, W, delta, n), phi6 = dot_dyn(Rm + 6, W, delta, n);
    const float m45 = phi4 < phi5 ? phi5 : phi4, m56 = phi5 < phi6 ? phi6 : phi5;
    step04 = capsules_distance(rb, sk, 0, 4, eq_thr, max_dist) / m45;
    step05 = capsules_distance(rb, sk, 0, 5, eq_thr, max_dist) / m56;
    step14 = capsules_distance(rb, sk, 1, 4, eq_thr, max_dist) / (phi2 + m45);
    step15 = capsules_distance(rb, sk, 1, 5, eq_thr, max_dist) / (phi2 + m56);
    const float a = step05 < step04 ? step05 : step04, b = step15 < step14 ? step15 : step14;
    return b < a ? b : a;
}

// xArm6::checkSelfCollision(q1, q2&) (xArm6.cpp:284-368), capsule stage, one thread.  The FCL mesh confirmation of
// checkRealSelfCollision (xArm6.cpp:426-464) is outside this engine: a negative capsule distance counts as a self-collision
// (conservative).  On a hit q2 is overwritten with the last collision-free configuration.  Every other robot class
// answers false (Planar2DOF.cpp:151-161, Spatial10DOF.cpp:32-43).  Out of line: it is called from the spine kernels' rare path.
template <class F>
__device__ __noinline__ bool self_collision_edge_of(const RobotDev *__restrict__ rb, const float *q1, float *q2, float eq_thr, float max_dist) {
    if (rb->kind != RVS_ROBOT_XARM6 || !rb->self_collision || rb->n != 6) return false;
    constexpr int n = 6, W = 7;
    float sk[3 * W], Rm[n * W], delta[n], q1t[n], qn[n];
    float step04 = 0.f, step05 = 0.f, step14 = 0.f, step15 = 0.f;
    for (int i = 0; i < n; i++) q1t[i] = q1[i];
    for (int num_iter = 0;;) {
        fk_skeleton<F, float>(rb, q1t, sk);
        enclosing_radii_of<float>(rb, sk, Rm);
        for (int i = 0; i < n; i++) delta[i] = fabsf(q2[i] - q1t[i]);
        const float step = self_collision_steps(rb, sk, Rm, delta, eq_thr, max_dist, step04, step05, step14, step15);
        if (step > 1.f) return false;
        if (++num_iter == 5 || step <= 0.f) break;
        if (step > 0.f)
            for (int i = 0; i < n; i++) q1t[i] = q1t[i] + step * (q2[i] - q1t[i]);
    }
    const float D = coord_norm(q2, q1t, n), eps = rb->eps_step;
    float dist = D;
    for (bool advanced = true; advanced;) {
        if (dist > eps) { for (int i = 0; i < n; i++) qn[i] = q1t[i] + (q2[i] - q1t[i]) / dist * eps; }
        else { for (int i = 0; i < n; i++) qn[i] = q2[i]; advanced = false; }
        dist = coord_norm(q2, qn, n);
        const float step_crit = 1.f - dist / D;
        fk_skeleton<F, float>(rb, qn, sk);
        // pairs in the reference's order {0-5, 1-5, 0-4, 1-4}; a pair is skipped while its own bubble step covers the marched fraction
        const bool hit = (!(step05 > step_crit) && capsules_distance(rb, sk, 0, 5, eq_thr, max_dist) < 0.f) ||
                         (!(step15 > step_crit) && capsules_distance(rb, sk, 1, 5, eq_thr, max_dist) < 0.f) ||
                         (!(step04 > step_crit) && capsules_distance(rb, sk, 0, 4, eq_thr, max_dist) < 0.f) ||
                         (!(step14 > step_crit) && capsules_distance(rb, sk, 1, 4, eq_thr, max_dist) < 0.f);
        if (hit) { for (int i = 0; i < n; i++) q2[i] = q1t[i]; return true; }
        for (int i = 0; i < n; i++) q1t[i] = qn[i];
    }
    return false;
}

// rvs_self_collision_edges: a.q = q1 [E x n], a.q2 = q2 [E x n]; a.verdict[E] = 1 on a self-collision, a.d_c = q2 after the call [E x n]
template <class F, class R>
__global__ void __launch_bounds__(kBlock) self_collision_edge_kernel(const KArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t e = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    if (e >= a.N) return;
    float q1[kMaxDofs], q2[kMaxDofs];
    for (int i = 0; i < n; i++) { q1[i] = a.q[e * n + i]; q2[i] = a.q2[e * n + i]; }
    const bool hit = self_collision_edge_of<F>(rb, q1, q2, a.eq_thr, a.max_dist);
    if (!a.d_c) { if (hit) a.verdict[e] = 0; return; }          // rvs_check_edges: clear the validity verdict of a self-colliding motion
    a.verdict[e] = hit ? 1 : 0;
    float *out = static_cast<float *>(a.d_c) + e * n;
    for (int i = 0; i < n; i++) out[i] = q2[i];
}

// The same routine for a PLANNER-SIZED batch: one WARP per edge.  A lone thread walks up to five bubble iterations and then the EPS_STEP
// marching, one FP64 forward kinematics each — 30 to 150 us of dependent instructions for a single edge.  Here the bubble iterations use
// the warp-cooperative FK (lane k evaluates sincos(q_k)), rows of the enclosing-radii matrix and the four capsule distances are spread
// over lanes, and the marching points — which do not depend on the verdicts — are laid out by lane 0 and tested 32 at a time, one per
// lane; the first hit in marching order decides.  Every value is computed by the same instruction sequence as in
// self_collision_edge_of(), so verdicts and truncated configurations are bit-identical (tests/test_gpu_parity.py).
constexpr int kScWarps = 4;
// self_collision_steps() by a whole warp: lanes 0-3 take one capsule distance each, lanes 4-7 one phi each; same arithmetic per value,
// so the four steps and their minimum are those of the one-thread version.  sk / Rm / delta may live in shared memory (every lane reads them).
__device__ __forceinline__ float self_collision_steps_warp(const RobotDev *__restrict__ rb, const float *sk, const float *Rm, const float *delta, float eq_thr,
                                                           float max_dist, float &step04, float &step05, float &step14, float &step15) {
    constexpr int n = 6, W = 7;
    const int lane = threadIdx.x & 31;
    float v = 0.f;
    if (lane < 4) v = capsules_distance(rb, sk, lane >> 1, 4 + (lane & 1), eq_thr, max_dist);          // 0-4, 0-5, 1-4, 1-5
    else if (lane < 8) v = dot_dyn(Rm + (lane == 4 ? 2 : lane - 1), W, delta, n);                      // phi2, phi4, phi5, phi6
    const float d04 = __shfl_sync(0xffffffffu, v, 0), d05 = __shfl_sync(0xffffffffu, v, 1), d14 = __shfl_sync(0xffffffffu, v, 2), d15 = __shfl_sync(0xffffffffu, v, 3);
    const float phi2 = __shfl_sync(0xffffffffu, v, 4), phi4 = __shfl_sync(0xffffffffu, v, 5), phi5 = __shfl_sync(0xffffffffu, v, 6), phi6 = __shfl_sync(0xffffffffu, v, 7);
    const float m45 = phi4 < phi5 ? phi5 : phi4, m56 = phi5 < phi6 ? phi6 : phi5;
    step04 = d04 / m45; step05 = d05 / m56; step14 = d14 / (phi2 + m45); step15 = d15 / (phi2 + m56);
    const float sa = step05 < step04 ? step05 : step04, sb = step15 < step14 ? step15 : step14;
    return sb < sa ? sb : sa;
}

// The routine itself, executed by ALL 32 lanes of a warp with identical q1 / q2 in their registers (at most kScWarps warps per CTA: the
// scratch is static shared memory indexed by the warp's number in the CTA).  Returns the verdict (warp-uniform); on a hit q2 holds the
// last collision-free configuration in every lane.
template <class F>
__device__ __noinline__ bool self_collision_edge_warp(const RobotDev *__restrict__ rb, const float *q1, float *q2, float eq_thr, float max_dist) {
    constexpr int n = 6, W = 7;
    __shared__ float s_sk[kScWarps][3 * W], s_Rm[kScWarps][n * W], s_delta[kScWarps][n], s_pts[kScWarps][33][n], s_dist[kScWarps][32];
    if (rb->kind != RVS_ROBOT_XARM6 || !rb->self_collision || rb->n != 6) return false;
    const int warp = (threadIdx.x >> 5) % kScWarps, lane = threadIdx.x & 31;
    float *sk = s_sk[warp], *Rm = s_Rm[warp], *delta = s_delta[warp];
    float q1t[n];
    for (int i = 0; i < n; i++) q1t[i] = q1[i];
    float step04 = 0.f, step05 = 0.f, step14 = 0.f, step15 = 0.f;
    for (int num_iter = 0;;) {                                                  // xArm6.cpp:297-325
        __syncwarp();
        float mine[3 * W];
        fk_skeleton_warp<F, float>(rb, q1t, mine);
        if (lane == 0) for (int k = 0; k < 3 * W; k++) sk[k] = mine[k];
        __syncwarp();
        if (lane < n) { enclosing_radii_row<float>(rb, sk, lane, Rm + lane * W); delta[lane] = fabsf(q2[lane] - q1t[lane]); }
        __syncwarp();
        const float step = self_collision_steps_warp(rb, sk, Rm, delta, eq_thr, max_dist, step04, step05, step14, step15);
        if (step > 1.f) return false;
        if (++num_iter == 5 || step <= 0.f) break;
        if (step > 0.f)
            for (int i = 0; i < n; i++) q1t[i] = q1t[i] + step * (q2[i] - q1t[i]);
    }
    // marching (xArm6.cpp:327-362): q_m = q_{m-1} + (q2 - q_{m-1}) / dist * EPS_STEP until q2 is reached; pts[0] = q1_temp
    const float D = coord_norm(q2, q1t, n), eps = rb->eps_step;
    float (*pts)[n] = s_pts[warp];
    float dist = D;
    bool advanced = true;
    __syncwarp();
    if (lane == 0) for (int i = 0; i < n; i++) pts[0][i] = q1t[i];
    while (advanced) {
        int count = 0;
        __syncwarp();
        if (lane == 0) {
            for (; count < 32 && advanced; count++) {
                const float *prev = pts[count];
                float *qn = pts[count + 1];
                if (dist > eps) { for (int i = 0; i < n; i++) qn[i] = prev[i] + (q2[i] - prev[i]) / dist * eps; }
                else { for (int i = 0; i < n; i++) qn[i] = q2[i]; advanced = false; }
                dist = coord_norm(q2, qn, n);
                s_dist[warp][count] = dist;
            }
        }
        count = __shfl_sync(0xffffffffu, count, 0);
        advanced = __shfl_sync(0xffffffffu, int(advanced), 0) != 0;
        dist = __shfl_sync(0xffffffffu, dist, 0);
        __syncwarp();
        bool mine_hit = false;
        if (lane < count) {
            const float step_crit = 1.f - s_dist[warp][lane] / D;
            float skm[3 * W];
            fk_skeleton<F, float>(rb, pts[lane + 1], skm);
            mine_hit = (!(step05 > step_crit) && capsules_distance(rb, skm, 0, 5, eq_thr, max_dist) < 0.f) ||
                       (!(step15 > step_crit) && capsules_distance(rb, skm, 1, 5, eq_thr, max_dist) < 0.f) ||
                       (!(step04 > step_crit) && capsules_distance(rb, skm, 0, 4, eq_thr, max_dist) < 0.f) ||
                       (!(step14 > step_crit) && capsules_distance(rb, skm, 1, 4, eq_thr, max_dist) < 0.f);
        }
        const unsigned hb = __ballot_sync(0xffffffffu, mine_hit);
        if (hb) {
            const int f = __ffs(int(hb)) - 1;                                   // first colliding point: q2 := the point before it
            for (int i = 0; i < n; i++) q2[i] = pts[f][i];
            __syncwarp();
            return true;
        }
        if (advanced) {
            __syncwarp();
            if (lane == 0) for (int i = 0; i < n; i++) pts[0][i] = pts[count][i];          // next round starts where this one ended
        }
    }
    return false;
}

template <class F, class R>
__global__ void __launch_bounds__(32 * kScWarps) self_collision_edge_warp_kernel(const KArgs a) {
    constexpr int n = 6;
    const RobotDev *__restrict__ rb = a.robot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t e = int64_t(blockIdx.x) * kScWarps + warp;
    if (e >= a.N) return;                                                       // (whole warps leave together)
    float q1[n], q2[n];
    for (int i = 0; i < n; i++) { q1[i] = a.q[e * n + i]; q2[i] = a.q2[e * n + i]; }
    const bool hit = self_collision_edge_warp<F>(rb, q1, q2, a.eq_thr, a.max_dist);
    if (lane != 0) return;
    if (!a.d_c) { if (hit) a.verdict[e] = 0; return; }
    a.verdict[e] = hit ? 1 : 0;
    float *out = static_cast<float *>(a.d_c) + e * n;
    for (int i = 0; i < n; i++) out[i] = q2[i];
}

template <class F>
__global__ void __launch_bounds__(64) spine_kernel(const SpineArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n, L = n, W = n + 1;
    const int s = blockIdx.x * 64 + threadIdx.x;
    if (s >= a.S) return;
    const float *q_e = a.q_targets + size_t(s) * n;
    float *q_out = a.q_reached + size_t(s) * n;
    float q_cur[kMaxDofs], q_new[kMaxDofs], prof_cur[kMaxDofs], prof_new[kMaxDofs], delta[kMaxDofs], rho_profile[kMaxDofs];
    float sk[3 * (kMaxDofs + 1)], sk_new[3 * (kMaxDofs + 1)];
    float Rm[kMaxDofs * (kMaxDofs + 1)];
    for (int i = 0; i < n; i++) { q_cur[i] = a.q_root[i]; q_new[i] = a.q_root[i]; prof_cur[i] = a.root_profile[i]; }
    float dc_cur = a.root_d_c, d_out = a.root_d_c;
    int status = 2;                                                     // Trapped when the root itself has no clearance (:91-92)
    if (a.root_d_c > 0.f) {
        status = 0;
        for (int layer = 0; layer < a.num_layers; layer++) {
            // ---- extendSpine(q_cur, q_e)
            status = 1;
            fk_skeleton<F, float>(rb, q_cur, sk);
            float rho = 0.f;
            for (int i = 0; i < L; i++) rho_profile[i] = 0.f;
            float q_temp[kMaxDofs];
            for (int i = 0; i < n; i++) q_temp[i] = q_cur[i];
            for (int counter = 1;; counter++) {
                if (counter == 1) enclosing_radii_of<float>(rb, sk, Rm);
                else { fk_skeleton<F, float>(rb, q_temp, sk_new); enclosing_radii_of<float>(rb, sk_new, Rm); }
                for (int i = 0; i < n; i++) delta[i] = fabsf(q_e[i] - q_temp[i]);
                float step;
                if (a.use_expanded_bubble) {
                    step = __int_as_float(0x7f800000);
                    for (int i = 0; i < n; i++) {
                        const float v = (prof_cur[i] - rho_profile[i]) / dot_dyn(Rm + (i + 1), W, delta, n);
                        if (v < step) step = v;                       // std::min(step, v)
                    }
                } else
                    step = (dc_cur - rho) / dot_dyn(Rm + n, W, delta, n);
                if (step > 1.f) {
                    for (int i = 0; i < n; i++) q_new[i] = q_e[i];
                    status = 3;
                } else
                    for (int i = 0; i < n; i++) q_new[i] = q_temp[i] + step * (q_e[i] - q_temp[i]);
                // RBTConnect.cpp:129-131: a self-collision on [q_temp, q_new] truncates q_new; Trapped when nothing was gained
                const bool self_collision = self_collision_edge_of<F>(rb, q_temp, q_new, a.eq_thr, a.max_dist);
                if (self_collision) status = !(coord_norm(q_temp, q_new, n) < a.eq_thr) ? 1 : 2;
                if (counter == a.num_iter_spine || status != 1 || self_collision) break;
                fk_skeleton<F, float>(rb, q_new, sk_new);
                float rho_k_prev = 0.f;
                for (int k = 1; k <= L; k++) {
                    const float rho_k = sqrtf(sqD(sk[3 * k] - sk_new[3 * k], sk[3 * k + 1] - sk_new[3 * k + 1], sk[3 * k + 2] - sk_new[3 * k + 2]));
                    rho_profile[k - 1] = rho_k_prev < rho_k ? rho_k : rho_k_prev;      // std::max(rho_k_prev, rho_k)
                    rho_k_prev = rho_k;
                    rho = rho < rho_k ? rho_k : rho;
                }
                for (int i = 0; i < n; i++) q_temp[i] = q_new[i];
            }
            // ---- distance-to-planes of the new state w.r.t. the ROOT's nearest points (RGBTConnect.cpp:86-88)
            fk_skeleton<F, float>(rb, q_new, sk_new);
            bool rejected;
            d_out = underestimation_of<float>(rb, L, a.M, sk_new, a.root_planes, a.max_dist, prof_new, rejected);
            if (d_out < a.d_crit || status != 1) break;
            for (int i = 0; i < n; i++) { q_cur[i] = q_new[i]; prof_cur[i] = prof_new[i]; }
            dc_cur = d_out;
        }
    }
    for (int i = 0; i < n; i++) q_out[i] = q_new[i];
    a.d_c_reached[s] = d_out;
    a.status[s] = uint8_t(status);
}

// ---- roofline denominators: register-only FMA microbenchmark (FFMA / DFMA issue peak of this device) ---------------------------
#ifdef RVS_VARIANT_EXACT
template <class T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *sink, int iters, T seed) {
    T acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = seed + T(k) + T(threadIdx.x);
    const T m = T(1.0000001), c = T(1e-9);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = fma(acc[k], m, c);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += acc[k];
    if (s == T(-1)) sink[0] = s;        // never true: keeps the chains alive
}
#endif

// ---- stateless synthetic configurations ----------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
#ifdef RVS_VARIANT_EXACT
__global__ void __launch_bounds__(256) generate_configs_kernel(const KArgs a, float *out) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n;
    const int64_t idx = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= a.N * n) return;
    const int64_t i = idx / n;
    const int j = int(idx - i * n);
    const float u = float(splitmix64(a.seed ^ (uint64_t(a.first + i) * 32u + uint64_t(j))) >> 40) * (1.0f / 16777216.0f);
    out[idx] = rb->limit_lo[j] + (rb->limit_hi[j] - rb->limit_lo[j]) * u;
}
#endif

}  // namespace rvs
