// rvs_env.cuh — the dynamic environment on the device.
//
// Reference: env::Environment::updateEnvironment and its motion laws (src/environments/Environment.cpp:194-377), the AABB refresh of
// env::Object::setPosition (src/environments/Object.cpp:21-26) and planning::trajectory::MotionValidity::check
// (src/planners/trajectory/MotionValidity.cpp:30-51): K times { move every dynamic obstacle by delta_time; interpolate the robot
// to validity point k; isValid }.  In the reference every validity point costs a host-side environment update and — behind a
// device engine — an obstacle-table upload.  Here the obstacle states live in HBM (one 64-byte record each) and one call does
//   env_advance_kernel     one thread per obstacle steps its own law K times and stores the state after every step (timeline[K][M]);
//                          obstacles never interact in any of the laws, so the K steps of one obstacle are its only dependency
//   motion_valid_kernel    one thread per (motion, validity point k): interpolation, FK, capsules against the AABBs of timeline[k]
//   env_commit_kernel      adopts the state the reference would be left in (after first_invalid or K updates) and rewrites the
//                          resident obstacle table
//   env_rank_kernel        re-sorts the scan-order copy of the table (nearest to the robot base first) by ranking, no host round trip
// Arithmetic: FP32, no FMA contraction, IEEE division / square root (this header is compiled in the EXACT translation unit), same
// operation order as the reference's Eigen expressions -> bit-identical to oracle/rvs_oracle_env.hpp for the straight, two_tunnels
// and light_directions laws; the circular law calls atan2f / sinf / cosf, whose CUDA and glibc implementations differ in the
// last ulp (tests use 1e-5 m there).  random_directions draws from libc rand() in the reference and is not supported.
#pragma once
#include <cuda_runtime.h>

#include "rvs_fk.cuh"
#include "rvs_geom.cuh"
#include "rvs_kernels.cuh"
#include "rvs_types.h"

namespace rvs {

__device__ __forceinline__ float env_norm3(float x, float y, float z) { return sqrtf(x * x + (y * y + z * z)); }      // Eigen fixed-size reduction order

// one Environment::updateEnvironment(dt) for one obstacle
__device__ __forceinline__ void env_step(EnvRec &o, const EnvArgs &a) {
    if (!o.dynamic) return;
    const float dt = a.dt;
    switch (a.motion_type) {
        case RVS_MOTION_STRAIGHT:                                        // Environment.cpp:224-238
            o.pos[0] -= dt * o.max_vel;
            break;
        case RVS_MOTION_CIRCULAR: {                                      // :241-258
            const float phi = atan2f(o.pos[1], o.pos[0]);
            o.vel[0] = -sinf(phi) * o.max_vel; o.vel[1] = cosf(phi) * o.max_vel; o.vel[2] = 0.0f * o.max_vel;
            for (int c = 0; c < 3; c++) o.pos[c] = o.pos[c] + dt * o.vel[c];
            break;
        }
        case RVS_MOTION_TWO_TUNNELS: {                                   // :261-289
            const float path_len_max = 0.2f;
            if (o.path_len > path_len_max) { o.sign = -o.sign; o.path_len = -path_len_max; }
            const float s = float(o.sign);
            if (o.pos[1] > 0) { o.vel[0] = 1.0f * s * o.max_vel; o.vel[1] = 0.0f * s * o.max_vel; o.vel[2] = 0.0f * s * o.max_vel; }
            else { o.vel[0] = 0.0f * s * o.max_vel; o.vel[1] = 1.0f * s * o.max_vel; o.vel[2] = 0.0f * s * o.max_vel; }
            o.path_len += env_norm3(o.vel[0], o.vel[1], o.vel[2]) * dt;
            for (int c = 0; c < 3; c++) o.pos[c] = o.pos[c] + dt * o.vel[c];
            break;
        }
        case RVS_MOTION_LIGHT_DIRECTIONS: {                              // :336-377
            const float tol_radius = fmaxf(env_norm3(o.vel[0], o.vel[1], o.vel[2]) / a.robot_max_vel, a.base_radius);
            const float nx = o.pos[0] + dt * o.vel[0], ny = o.pos[1] + dt * o.vel[1], nz = o.pos[2] + dt * o.vel[2];      // pos_next
            const float rc = env_norm3(nx - a.ws_center[0], ny - a.ws_center[1], nz - a.ws_center[2]);
            float vx, vy, vz;
            bool change = true;
            if (nz < 0) { vx = 0; vy = 0; vz = 1; }
            else if (rc > a.ws_radius) { vx = -nx; vy = -ny; vz = -(nz - a.ws_center[2]); }
            else if (sqrtf(nx * nx + ny * ny) < tol_radius && nz < a.ws_center[2]) { vx = nx; vy = ny; vz = 0; }
            else if (rc < tol_radius) { vx = nx; vy = ny; vz = nz - a.ws_center[2]; }
            else { o.pos[0] = nx; o.pos[1] = ny; o.pos[2] = nz; change = false; vx = vy = vz = 0; }
            if (change) {
                const float dx = nx - o.pos[0], dy = ny - o.pos[1], dz = nz - o.pos[2];
                const float t_param = (dx * vx + (dy * vy + dz * vz)) / (vx * vx + (vy * vy + vz * vz));
                const float two_t = 2 * t_param;
                const float px = 2 * nx - o.pos[0] - two_t * vx, py = 2 * ny - o.pos[1] - two_t * vy, pz = 2 * nz - o.pos[2] - two_t * vz;
                o.pos[0] = px; o.pos[1] = py; o.pos[2] = pz;
                o.vel[0] = (px - nx) / dt; o.vel[1] = (py - ny) / dt; o.vel[2] = (pz - nz) / dt;
            }
            break;
        }
        default: break;
    }
}

// AABB after setPosition (Object.cpp:21-26; FCL computeAABB of a box: centre -/+ half extents)
__device__ __forceinline__ ObstacleRec env_aabb(const EnvRec &o) {
    ObstacleRec r;
    for (int c = 0; c < 3; c++) { r.lo[c] = o.pos[c] + (-o.half[c]); r.hi[c] = o.pos[c] + o.half[c]; }
    r.tol = o.tol; r.flags = o.flags;
    return r;
}

#ifdef RVS_VARIANT_EXACT
__global__ void __launch_bounds__(128) env_advance_kernel(const EnvArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.M) return;
    EnvRec o = a.state[j];
    for (int k = 0; k < a.K; k++) {
        env_step(o, a);
        a.timeline[size_t(k) * a.M + j] = o;
    }
}
#endif

// One thread per (motion e, validity point k = 1..K).  first_invalid[e] = min k whose configuration collides with the environment
// after k updates (atomicMin; initialised to INT_MAX by the host = "valid").
template <class F, class R>
__global__ void __launch_bounds__(kBlock) motion_valid_kernel(const EnvArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = rb->n, L = n;
    const int64_t item = int64_t(blockIdx.x) * kBlock + threadIdx.x;
    const int64_t e = item / a.K;
    if (e >= a.E) return;
    const int k = int(item - e * a.K) + 1;
    float qn[kMaxDofs];
    if (a.points) {
        // MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-74): the k-th trajectory point as it is
        const float *pt = a.points + (e * a.K + (k - 1)) * n;
        for (int i = 0; i < n; i++) qn[i] = pt[i];
    } else {
        // interpolateEdge(q_previous, q_current, dist * k / K, dist)  (MotionValidity.cpp:43, RealVectorSpace.cpp:75-88)
        const float *qp = a.q_prev + e * n, *qc = a.q_cur + e * n;
        const float dist = coord_norm(qp, qc, n);
        const float step = dist * float(k) / float(a.K);
        if (dist > 0.f) {
            const float f = step / dist;
            for (int i = 0; i < n; i++) qn[i] = qp[i] + (qc[i] - qp[i]) * f;
        } else
            for (int i = 0; i < n; i++) qn[i] = qc[i];
    }
    R sk[3 * (kMaxDofs + 1)];
    fk_skeleton<F, R>(rb, qn, sk);
    const EnvRec *__restrict__ tl = a.timeline + size_t(k - 1) * a.M;
    const R eq_thr = R(a.eq_thr), max_dist = R(a.max_dist);
    const int ground_included = rb->ground_included;
    bool valid = true;
#pragma unroll 1
    for (int i = 0; i < L && valid; i++) {
        Capsule<R> cp;
        cp.set({sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]}, RobotScalars<R>::radius(rb, i), RobotScalars<R>::radius_in(rb, i));
#pragma unroll 1
        for (int j = 0; j < a.M; j++) {
            const ObstacleRec o = env_aabb(tl[j]);
            if ((o.flags & 1u) && i < ground_included) continue;
            if (capsule_hits_box<R>(cp, load_box<R>(o), eq_thr, max_dist)) { valid = false; break; }
        }
    }
    if (!valid) atomicMin(a.first_invalid + e, k);
}

#ifdef RVS_VARIANT_EXACT
// distance of a box from the robot base, rounded DOWN to float (the cut-off of the work-queue kernels must never overestimate it);
// same arithmetic as upload_obstacles() on the host
__device__ __forceinline__ float env_base_distance(const ObstacleRec &o, const RobotDev *rb) {
    const double bx = rb->xyz[0][0], by = rb->xyz[0][1], bz = rb->xyz[0][2];
    const double gx = fmax(fmax(double(o.lo[0]) - bx, bx - double(o.hi[0])), 0.0), gy = fmax(fmax(double(o.lo[1]) - by, by - double(o.hi[1])), 0.0),
                 gz = fmax(fmax(double(o.lo[2]) - bz, bz - double(o.hi[2])), 0.0);
    double d2 = gx * gx + gy * gy + gz * gz;
    if (!(d2 == d2)) d2 = 1e300;
    const double d = sqrt(fmin(d2, 1e60));
    float f = float(d);
    if (double(f) > d) f = __int_as_float(__float_as_int(f) - 1);          // f > d >= 0: the next float towards zero
    return f;
}

// Adopt the state after kc updates, kc = first_invalid[commit_edge] (or K when that motion is valid / commit_edge < 0), and rewrite
// the caller-order obstacle table.  scan_key[j] = distance from the base (input of env_rank_kernel).
__global__ void __launch_bounds__(128) env_commit_kernel(const EnvArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.M) return;
    int kc = a.K;
    if (a.commit_edge >= 0) { const int fi = a.first_invalid[a.commit_edge]; if (fi <= a.K) kc = fi; }
    const EnvRec o = a.timeline[size_t(kc - 1) * a.M + j];
    a.state[j] = o;
    const ObstacleRec r = env_aabb(o);
    a.obstacles[j] = r;
    a.scan_key[j] = env_base_distance(r, a.robot);
}

// scan order = ascending base distance, ties by index (what std::stable_sort gives on the host): rank by counting
__global__ void __launch_bounds__(128) env_rank_kernel(const EnvArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.M) return;
    const float key = a.scan_key[j];
    int rank = 0;
    for (int i = 0; i < a.M; i++) {
        const float ki = a.scan_key[i];
        rank += (ki < key || (ki == key && i < j)) ? 1 : 0;
    }
    ObstacleRec r = a.obstacles[j];
    r.flags = (r.flags & 1u) | (uint32_t(j) << 8);
    a.obstacles_scan[rank] = r;
    a.scan_dist[rank] = key;
}
#endif

}  // namespace rvs
