// rvs_oracle_env.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT A PRODUCT PATH) of the dynamic environment:
// the obstacle motion laws of env::Environment::updateEnvironment and the discretised motion check of
// planning::trajectory::MotionValidity::check, restated from
//   /root/reference/src/environments/Environment.cpp:194-377   (moveStraight, moveCircular, moveTwoTunnels, moveLightDirections)
//   /root/reference/src/environments/Object.cpp:21-26 + objects/Box.cpp:3-8   (setPosition -> computeAABB: pos -/+ dim/2)
//   /root/reference/src/planners/trajectory/MotionValidity.cpp:30-51          (check(q_previous, q_current))
// in plain FP32 with Eigen's evaluation order (fixed-size Vector3f reductions as x0 + (x1 + x2), see rvs_oracle.hpp).
//
// PARITY STATUS: the reference ships no golden vectors for the motion laws.  The LOGIC of this file is pinned against the reference's
// own Environment.cpp / Object.cpp / objects/Box.cpp, compiled unmodified from /root/reference against oracle/eigen_shim plus
// stand-ins for the slices of FCL and yaml-cpp they touch (oracle/env_shim, oracle/env_ref_capi.cpp, Makefile target `ref`):
// positions, velocities, the two_tunnels bookkeeping and the AABBs agree bit for bit over all four laws, random restarts included
// (tests/test_env.py; committed vectors: tests/golden/env_golden.npz from tests/golden/make_env_golden.py).  MotionValidity::check is
// pinned the same way: MotionValidity.cpp compiled unmodified on top of the reference's own Environment / RealVectorSpace
// (oracle/rvs_ref_capi.cpp: rvsref_motion_run; tests/golden/motion_golden.npz) — runs of consecutive checks agree bit for bit in
// verdicts, passed validity points and the obstacles' state after every check.  What stays "parity
// unpinned" is third-party arithmetic: Eigen's reduction order (the shim encodes the same assumption) and FCL's
// computeAABB (pos -/+ side/2).  Closed-form known answers are checked as well (straight-line motion, circular motion
// spirals out by the Euler step, the reflection law preserves speed and keeps every obstacle inside the workspace shell).
// moveRandomDirections (:289-333) draws from Eigen::Vector3f::Random, i.e. libc rand(): restated here on top of LibcRand, a restatement of
// glibc's rand() (random_r TYPE_3: the additive feedback generator r[i] = r[i-3] + r[i-31] over a 34-word table seeded by the
// Lehmer recurrence 16807 * x mod (2^31 - 1), first 310 outputs discarded, result = r[i] >> 1), so that a run after srand(seed) in the
// reference and a run after LibcRand::seed(seed) here see the same stream (checked against the C library's own rand(), tests/test_env.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "rvs_oracle.hpp"

namespace orc {

enum MotionType { MOTION_STRAIGHT = 0, MOTION_CIRCULAR = 1, MOTION_TWO_TUNNELS = 2, MOTION_RANDOM_DIRECTIONS = 3, MOTION_LIGHT_DIRECTIONS = 4 };

// glibc rand() after srand(seed) (stdlib/random_r.c, TYPE_3); state = 34 table words + position + number of draws so far
struct LibcRand {
    uint32_t r[34];
    uint32_t pos = 0;          // index (mod 34) of the next word to produce
    uint32_t draws = 0;
    void seed(uint32_t s) {
        if (s == 0) s = 1;
        int32_t word = int32_t(s);
        r[0] = uint32_t(word);
        for (int i = 1; i < 31; i++) {
            const int32_t hi = word / 127773, lo = word % 127773;           // 16807 * word mod (2^31 - 1) without overflow (Schrage)
            word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = uint32_t(word);
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        pos = 0; draws = 0;                                                 // word 34 is the next one: 34 mod 34 = 0
        for (int i = 0; i < 310; i++) next();
        draws = 0;
    }
    uint32_t next() {                                                       // r[i] = r[i - 31] + r[i - 3]; rand() returns r[i] >> 1
        const uint32_t v = r[(pos + 3) % 34] + r[(pos + 31) % 34];
        r[pos] = v;
        pos = (pos + 1) % 34;
        draws++;
        return v >> 1;
    }
    // Eigen::internal::random<float>() as the work-alike (and Eigen 3.4 for floats) evaluates it: 2 * rand() / RAND_MAX - 1
    float uniform() { return 2.0f * float(next()) / float(2147483647) - 1.0f; }
};

struct EnvParams {
    int motion_type;
    V3<float> ws_center;
    float ws_radius, base_radius, robot_max_vel;
    int ground_included = 2;   // Environment::ground_included (only Environment::isValid reads it: random_directions)
    LibcRand *rng = nullptr;   // random_directions
};

struct EnvObject {
    V3<float> half;          // half extents of the AABB FCL computes (0.5 * dim; bounding-sphere cube for rotated boxes)
    V3<float> pos, vel;
    float max_vel;
    int sign;                // Environment::sign(i)      (two_tunnels)
    float path_len;          // Environment::path_len(i)  (two_tunnels)
    bool dynamic;            // label == "dynamic_obstacle"
    V3<float> acc{0, 0, 0};  // env::Object::acceleration (random_directions)
};

// Environment::isValid(pos, vel) (Environment.cpp:169-190)
inline bool env_pos_valid(const EnvParams &ep, const V3<float> &pos, float vel) {
    const float tol_radius = std::max(vel / ep.robot_max_vel, ep.base_radius);
    const float rc = normF(pos - ep.ws_center);
    const float rxy = std::sqrt(pos.x * pos.x + pos.y * pos.y);
    if (ep.ground_included > 0) {
        if (rc > ep.ws_radius || pos.z < 0 || (rxy < tol_radius && pos.z < ep.ws_center.z) || rc < tol_radius) return false;
    } else {
        if (rc > ep.ws_radius || (rxy < tol_radius && pos.z < ep.ws_center.z && pos.z > -ep.base_radius) || rc < tol_radius) return false;
    }
    return true;
}

inline float norm2(float x, float y) { return std::sqrt(x * x + y * y); }

// one updateEnvironment(delta_time) (Environment.cpp:194-222)
inline void env_update(const EnvParams &ep, std::vector<EnvObject> &objs, float dt) {
    for (size_t i = 0; i < objs.size(); i++) {
        EnvObject &o = objs[i];
        if (!o.dynamic) continue;
        switch (ep.motion_type) {
            case MOTION_STRAIGHT:                                       // (:224-238)
                o.pos.x -= dt * o.max_vel;
                break;
            case MOTION_CIRCULAR: {                                     // (:241-258)
                const float phi = std::atan2(o.pos.y, o.pos.x);
                o.vel = {-std::sin(phi) * o.max_vel, std::cos(phi) * o.max_vel, 0.0f * o.max_vel};
                o.pos = o.pos + dt * o.vel;
                break;
            }
            case MOTION_TWO_TUNNELS: {                                  // (:261-289)
                const float path_len_max = 0.2f;
                if (o.path_len > path_len_max) { o.sign *= -1; o.path_len = -path_len_max; }
                const float s = float(o.sign);
                if (o.pos.y > 0) o.vel = {1.0f * s * o.max_vel, 0.0f * s * o.max_vel, 0.0f * s * o.max_vel};
                else o.vel = {0.0f * s * o.max_vel, 1.0f * s * o.max_vel, 0.0f * s * o.max_vel};
                o.path_len += normF(o.vel) * dt;
                o.pos = o.pos + dt * o.vel;
                break;
            }
            case MOTION_LIGHT_DIRECTIONS: {                             // (:336-377)
                const float tol_radius = std::max(normF(o.vel) / ep.robot_max_vel, ep.base_radius);
                const V3<float> pos_next = o.pos + dt * o.vel;
                V3<float> n{0, 0, 0};
                bool change = true;
                if (pos_next.z < 0) n = {0, 0, 1};
                else if (normF(pos_next - ep.ws_center) > ep.ws_radius) n = {-pos_next.x, -pos_next.y, -(pos_next.z - ep.ws_center.z)};
                else if (norm2(pos_next.x, pos_next.y) < tol_radius && pos_next.z < ep.ws_center.z) n = {pos_next.x, pos_next.y, 0};
                else if (normF(pos_next - ep.ws_center) < tol_radius) n = {pos_next.x, pos_next.y, pos_next.z - ep.ws_center.z};
                else { o.pos = pos_next; change = false; }
                if (change) {
                    const float t_param = dotF(pos_next - o.pos, n) / sqnF(n);
                    const float two_t = 2 * t_param;
                    const V3<float> p_new{2 * pos_next.x - o.pos.x - two_t * n.x, 2 * pos_next.y - o.pos.y - two_t * n.y, 2 * pos_next.z - o.pos.z - two_t * n.z};
                    o.pos = p_new;
                    o.vel = {(p_new.x - pos_next.x) / dt, (p_new.y - pos_next.y) / dt, (p_new.z - pos_next.z) / dt};
                }
                break;
            }
            case MOTION_RANDOM_DIRECTIONS: {                            // (:289-333); the reference's `i--` = retry the same object
                for (int tries = 0; tries < 100000; tries++) {
                    const V3<float> vel = o.vel + dt * o.acc;
                    const V3<float> pos = o.pos + dt * vel;
                    const float vel_intensity = normF(vel);
                    if (vel_intensity > o.max_vel) {
                        V3<float> a{0, 0, 0};
                        a.x = ep.rng->uniform(); a.y = ep.rng->uniform(); a.z = ep.rng->uniform();
                        const float na = normF(a);
                        a = {a.x / na, a.y / na, a.z / na};                 // normalize(): *this /= norm()
                        const float keep = normF(o.acc);
                        o.acc = {keep * a.x, keep * a.y, keep * a.z};
                    } else if (!env_pos_valid(ep, pos, vel_intensity)) {
                        V3<float> v{0, 0, 0};
                        v.x = ep.rng->uniform(); v.y = ep.rng->uniform(); v.z = ep.rng->uniform();
                        const float nv = normF(v);
                        v = {v.x / nv, v.y / nv, v.z / nv};
                        o.vel = {vel_intensity * v.x, vel_intensity * v.y, vel_intensity * v.z};
                    } else {
                        o.vel = vel;
                        o.pos = pos;
                        break;
                    }
                }
                break;
            }
            default: break;
        }
    }
}

// AABB of object i as FCL's computeAABB leaves it after setPosition (Object.cpp:21-26)
inline void env_aabb(const EnvObject &o, float *mn, float *mx) {
    mn[0] = o.pos.x + (-o.half.x); mn[1] = o.pos.y + (-o.half.y); mn[2] = o.pos.z + (-o.half.z);
    mx[0] = o.pos.x + o.half.x;    mx[1] = o.pos.y + o.half.y;    mx[2] = o.pos.z + o.half.z;
}

// coordinates of the k-th validity point of MotionValidity::check (MotionValidity.cpp:40-44):
//   interpolateEdge(q_previous, q_current, dist * k / K, dist)   (RealVectorSpace.cpp:75-88)
inline void motion_point(const float *q_prev, const float *q_cur, int n, int k, int K, float *out) {
    const float dist = coord_norm(q_prev, q_cur, n);
    const float step = dist * float(k) / float(K);
    if (dist > 0) {
        const float f = step / dist;
        for (int i = 0; i < n; i++) out[i] = q_prev[i] + (q_cur[i] - q_prev[i]) * f;
    } else
        for (int i = 0; i < n; i++) out[i] = q_cur[i];
}

}  // namespace orc
