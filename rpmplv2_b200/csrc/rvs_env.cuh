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
// last ulp (tests use 1e-5 m there).
// random_directions (:289-333) draws from Eigen::Vector3f::Random = libc rand(), ONE stream shared by all obstacles in index order: the
// device keeps its own copy of that stream (LibcRand, rvs_types.h: glibc's additive feedback generator, seeded like srand) and
// env_advance_random_kernel (one warp) steps 32 obstacles at a time in parallel — an obstacle only touches the stream when its move is
// rejected — and serialises the rejected ones in index order, so the draws land on the same obstacles as in the reference's loop.
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

// ---- random_directions (:289-333) -----------------------------------------------------------------------------------------------------
// Environment::isValid(pos, vel) (:169-190)
__device__ __forceinline__ bool env_pos_valid(const EnvArgs &a, float x, float y, float z, float vel) {
    const float tol_radius = fmaxf(vel / a.robot_max_vel, a.base_radius);
    const float rc = env_norm3(x - a.ws_center[0], y - a.ws_center[1], z - a.ws_center[2]);
    const float rxy = sqrtf(x * x + y * y);
    if (a.env_ground_included > 0) return !(rc > a.ws_radius || z < 0 || (rxy < tol_radius && z < a.ws_center[2]) || rc < tol_radius);
    return !(rc > a.ws_radius || (rxy < tol_radius && z < a.ws_center[2] && z > -a.base_radius) || rc < tol_radius);
}
// one attempt of the loop body: 0 = moved, 1 = velocity above max_vel (wants a new acceleration), 2 = position invalid (wants a new velocity)
__device__ __forceinline__ int env_random_try(EnvRec &o, const EnvArgs &a, float &vel_intensity) {
    const float dt = a.dt;
    const float vx = o.vel[0] + dt * o.acc[0], vy = o.vel[1] + dt * o.acc[1], vz = o.vel[2] + dt * o.acc[2];
    const float px = o.pos[0] + dt * vx, py = o.pos[1] + dt * vy, pz = o.pos[2] + dt * vz;
    vel_intensity = env_norm3(vx, vy, vz);
    if (vel_intensity > o.max_vel) return 1;
    if (!env_pos_valid(a, px, py, pz, vel_intensity)) return 2;
    o.vel[0] = vx; o.vel[1] = vy; o.vel[2] = vz;
    o.pos[0] = px; o.pos[1] = py; o.pos[2] = pz;
    return 0;
}
// the reference's `i--` retries of ONE obstacle, drawing from the shared stream until the move is accepted (bounded: the reference
// would spin forever on an obstacle that cannot be placed, e.g. zero acceleration above max_vel)
__device__ __forceinline__ void env_random_retry(EnvRec &o, const EnvArgs &a, LibcRand &g, int why, float vel_intensity) {
    for (int tries = 0; tries < 100000 && why != 0; tries++) {
        const float rx = g.uniform(), ry = g.uniform(), rz = g.uniform();
        const float nr = env_norm3(rx, ry, rz);
        const float ux = rx / nr, uy = ry / nr, uz = rz / nr;                              // normalize(): *this /= norm()
        if (why == 1) {
            const float keep = env_norm3(o.acc[0], o.acc[1], o.acc[2]);
            o.acc[0] = keep * ux; o.acc[1] = keep * uy; o.acc[2] = keep * uz;
        } else {
            o.vel[0] = vel_intensity * ux; o.vel[1] = vel_intensity * uy; o.vel[2] = vel_intensity * uz;
        }
        why = env_random_try(o, a, vel_intensity);
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

// random_directions: ONE warp.  Per update and chunk of 32 obstacles: every lane tries its obstacle's move; the rejected ones are
// completed one after the other in index order with the shared libc stream (kept in shared memory); the stream state after every
// update is recorded next to the obstacle timeline.
__global__ void __launch_bounds__(32) env_advance_random_kernel(const EnvArgs a) {
    __shared__ LibcRand g;
    const int lane = threadIdx.x;
    for (int w = lane; w < kLibcRandWords; w += 32) reinterpret_cast<uint32_t *>(&g)[w] = reinterpret_cast<const uint32_t *>(a.rand_state)[w];
    __syncwarp();
    for (int k = 0; k < a.K; k++) {
        const EnvRec *__restrict__ from = k == 0 ? a.state : a.timeline + size_t(k - 1) * a.M;
        for (int c0 = 0; c0 < a.M; c0 += 32) {
            const int j = c0 + lane;
            EnvRec o{};
            int why = 0;
            float vel_intensity = 0.f;
            if (j < a.M) {
                o = from[j];
                if (o.dynamic) why = env_random_try(o, a, vel_intensity);
            }
            unsigned pending = __ballot_sync(0xffffffffu, why != 0);
            while (pending) {
                const int l = __ffs(int(pending)) - 1;
                pending &= pending - 1u;
                if (lane == l) env_random_retry(o, a, g, why, vel_intensity);
                __syncwarp();
            }
            if (j < a.M) a.timeline[size_t(k) * a.M + j] = o;
        }
        __syncwarp();
        for (int w = lane; w < kLibcRandWords; w += 32) reinterpret_cast<uint32_t *>(a.rand_timeline + k)[w] = reinterpret_cast<const uint32_t *>(&g)[w];
        __syncwarp();
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
    if (a.motion_type == RVS_MOTION_RANDOM_DIRECTIONS && j == 0) *a.rand_state = a.rand_timeline[kc - 1];    // the stream as the adopted updates left it
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
