// rvs_types.h — POD types shared by the C-ABI host code (rvs_capi.cu) and the sm_100a kernels.
#pragma once
#include <stdint.h>

#include "../../include/rvs_cuda.h"

#if defined(__CUDACC__)
#define RVS_HD __host__ __device__
#else
#define RVS_HD
#endif

namespace rvs {

constexpr int kMaxDofs = RVS_MAX_DOFS;
constexpr int kTileObstacles = 512;     // obstacles staged per shared-memory tile (32 B each -> 16 KB)

// Obstacle record as it sits in HBM and in shared memory: two 16-byte words, so that one obstacle is two
// broadcast LDS.128 for the whole warp.   w0 = (lo.x, lo.y, lo.z, hi.x)   w1 = (hi.y, hi.z, min_dist_tol, flags)
// flags (bit pattern in the float slot): bit0 = label == "ground".
struct ObstacleRec {
    float lo[3];
    float hi[3];
    float tol;
    uint32_t flags;
};
static_assert(sizeof(ObstacleRec) == 32, "ObstacleRec must be 32 bytes");

enum AxisCode : int32_t { AXIS_X = 0, AXIS_Y = 1, AXIS_Z = 2, AXIS_GENERAL = 3 };

// Robot chain resident in HBM (one copy per handle). Doubles first for alignment.
struct RobotDev {
    double xyz[kMaxDofs + 1][3];     // joint origin in parent frame
    double rfix[kMaxDofs + 1][9];    // RPY rotation of the joint origin, row-major
    double axis[kMaxDofs][3];        // unit joint axis
    double radius_d[kMaxDofs];       // capsule radii widened (F64 variant)
    double radius_in_d[kMaxDofs];    // radius*sqrt(3)/3 in double (F64 variant)
    float radius[kMaxDofs];          // AbstractRobot::capsules_radius
    float radius_in[kMaxDofs];       // float(double(radius)*sqrt(3)/3)  (CollisionAndDistance.cpp:7)
    float radius_new[kMaxDofs];      // xArm6::capsules_radius_new (xArm6.cpp:98-105); == radius for other kinds
    float limit_lo[kMaxDofs];
    float limit_hi[kMaxDofs];
    int32_t axis_code[kMaxDofs];
    int32_t rfix_identity[kMaxDofs + 1];
    int32_t kind;
    int32_t n;                       // DOFs == links
    int32_t ground_included;
    float gripper_length;
    int32_t self_collision;          // AbstractRobot::getSelfCollisionChecking (rvs_set_self_collision); only xArm6 has a check at all
    float eps_step;                  // RRTConnectConfig::EPS_STEP: marching step of xArm6::checkSelfCollision(q1, q2) (xArm6.cpp:340-341)
};

// Arguments of the hot-path kernels (passed by value).
struct KArgs {
    const RobotDev *robot;
    const ObstacleRec *obstacles;        // caller's order
    const ObstacleRec *obstacles_scan;   // nearest-to-base first; flags bits 8..31 = index in the caller's order (work-queue kernel only)
    const float *scan_dist;              // [M] distance of scan-order obstacle k from the robot base, rounded down (ascending)
    const void *obstacles_pack;          // scan-order table as fp16 boxes rounded outwards, two obstacles per record (rvs_distance_wq.cuh, pack_scan_kernel)
    int32_t M;
    int64_t N;
    const float *q;          // N x n
    const float *q2;         // edges: far end
    float eq_thr;
    float max_dist;
    int32_t num_checks;
    int32_t mode;            // rvs_distance_mode
    int32_t cull;
    int32_t slices;          // validity work-queue kernel: the (sorted) obstacle table is split over gridDim.y blocks per item batch
    void *d_c;               // float* or double*
    void *profile;
    void *nearest;
    uint8_t *verdict;
    // underestimation
    const float *planes;
    const int32_t *plane_idx;
    int32_t R;
    uint8_t *status;
    // generated configurations
    uint64_t seed;
    int64_t first;
    int32_t robot_n_hint;    // num_dofs as known by the host (grid sizing)
    int32_t num_ground;      // obstacles flagged ground
    unsigned long long *pair_count;   // += pair queries the reference's loop would have visited (early exit respected)
    // coarse distance field of the obstacle table (rvs_distance_wq.cuh: df_build_kernel): upper bounds for the first thresholds; nullptr = off
    const int32_t *perm;     // work-queue distance kernel: thread k works on configuration perm[k] (rvs_sort.cuh); nullptr = identity
    int32_t *list;           // FULL mode of a large batch (rvs_distance_full.cuh): [0] = number of collision-free configurations, their indices from [4]; nullptr = one kernel
    const void *df;          // [df_dim^3] __half, distance from the cell centre to the nearest eligible obstacle, rounded up
    float df_origin[3];      // lower corner of the grid
    float df_inv_cell;       // cells per metre
    float df_slack;          // half a cell diagonal (+ rounding): |p - centre| for any p of the cell
    int32_t df_dim;
};

// Arguments of the batched generalized-bur spine kernel (rvs_extend_spines).
struct SpineArgs {
    const RobotDev *robot;
    int32_t M, S;
    const float *q_root;          // [n]
    const float *root_profile;    // [L]
    float root_d_c;
    const float *root_planes;     // [M][L][6]
    const float *q_targets;       // [S][n]
    int32_t num_layers, num_iter_spine, use_expanded_bubble;
    float d_crit, max_dist, eq_thr;
    float *q_reached;             // [S][n]
    float *d_c_reached;           // [S]
    uint8_t *status;              // [S]  base::State::Status: 1 Advanced, 2 Trapped, 3 Reached
};

// ---- device-resident State cache (rvs_store.cuh) ----------------------------------------------------------------------
#ifdef __CUDACC__
#define RVS_HD __host__ __device__
#else
#define RVS_HD
#endif
RVS_HD inline size_t store_header_floats(int n) { return size_t((2 * n + 1 + 3) & ~3); }
RVS_HD inline size_t store_slot_floats(int n, int M) { return store_header_floats(n) + size_t(M) * size_t(n) * 6; }

struct StoreArgs {
    const RobotDev *robot;
    const ObstacleRec *obstacles;
    int32_t M, N, n;
    float eq_thr, max_dist;
    float *store;              // slot array
    const float *q;            // [N][n]
    const int32_t *slots;      // [N]
    float *d_c;                // [N]
    float *profile;            // [N][L] or nullptr
    // spines
    const int32_t *root_slot;  // [S]
    const float *q_targets;    // [S][n]
    int32_t S, num_layers, num_iter_spine, use_expanded_bubble, gen2;
    float d_crit;
    float *q_layers;           // [S][num_layers][n]
    int32_t *n_layers;         // [S]
    float *d_c_last;           // [S]
    uint8_t *status;           // [S]
    // connect chains (rvs_store_connect): the spine and distance kernels of step k are launched back to back on one stream and gate
    // themselves on device state, so a whole connectGenSpine loop costs one synchronisation
    int32_t *chain_state;      // [S] 1 = the chain of request s is still running; nullptr = plain call
    const int32_t *layers_n;   // distance kernel: [N] layers the spine of this step produced (< 0: step not executed); q = its last layer
};

// ---- nearest-neighbour / radius queries over device-resident planner trees (rvs_nn.cuh) ----------------------------------------------
struct NnArgs {
    const float *const *points;     // [T] device pointer of every tree's states
    const int32_t *counts;          // [T] states per tree
    int32_t n;                      // DOFs
    // nearest: Q queries, query k against tree tree_of[k]
    const int32_t *tree_of;         // [Q]
    const float *queries;           // [Q][n]
    int32_t Q;
    int32_t *idx;                   // [Q]  (-1: empty tree)
    float *dist2;                   // [Q]  squared distance (nanoflann's out_dist_sqr)
    // radius: ONE query against ONE tree
    int32_t tree;
    float radius;
    int32_t exclude;                // state index that is the query itself (-1: none)
    int32_t capacity;               // room in out_idx / out_dist
    int32_t *out_idx;               // hits in TREE ORDER
    float *out_dist;
    int32_t *out_count;             // total number of hits (may exceed capacity: the caller then re-queries with more room)
};

// ---- dynamic environment (rvs_env.cuh) -------------------------------------------------------------------------------
// One obstacle of env::Environment as it sits in HBM (64 bytes): what Environment::updateEnvironment reads and writes.
struct EnvRec {
    float half[3];        // half extents of the AABB that FCL's computeAABB produces (0.5 * dim; bounding-sphere cube if rotated)
    float pos[3];         // env::Object::position
    float vel[3];         // env::Object::velocity
    float max_vel;        // env::Object::max_vel
    float path_len;       // Environment::path_len(i)   (two_tunnels)
    int32_t sign;         // Environment::sign(i)       (two_tunnels)
    uint32_t dynamic;     // label == "dynamic_obstacle"
    float tol;            // min_dist_tol
    uint32_t flags;       // bit0: label == "ground"
    float acc[3];         // env::Object::acceleration (random_directions)
    uint32_t pad[2];
};
static_assert(sizeof(EnvRec) == 80, "EnvRec must be 80 bytes (a multiple of 16)");

// glibc rand() after srand(seed) (random_r TYPE_3: r[i] = r[i-3] + r[i-31] over a 34-word table, output r[i] >> 1), the stream
// Eigen::Vector3f::Random draws from in Environment::moveRandomDirections (Environment.cpp:309, :318).  36 words: table, position, draws.
constexpr int kLibcRandWords = 36;
struct LibcRand {
    uint32_t r[34];
    uint32_t pos, draws;
    RVS_HD uint32_t next() {
        const uint32_t v = r[(pos + 3) % 34] + r[(pos + 31) % 34];
        r[pos] = v;
        pos = (pos + 1) % 34;
        draws++;
        return v >> 1;
    }
    RVS_HD float uniform() { return 2.0f * float(next()) / float(2147483647) - 1.0f; }      // Eigen's random<float>(): [-1, 1]
    void seed(uint32_t s) {
        if (s == 0) s = 1;
        int32_t word = int32_t(s);
        r[0] = uint32_t(word);
        for (int i = 1; i < 31; i++) {
            const int32_t hi = word / 127773, lo = word % 127773;         // 16807 * word mod (2^31 - 1), Schrage's split
            word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = uint32_t(word);
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        pos = 0; draws = 0;
        for (int i = 0; i < 310; i++) next();
        draws = 0;
    }
};
static_assert(sizeof(LibcRand) == kLibcRandWords * 4, "LibcRand layout");

struct EnvArgs {
    const RobotDev *robot;
    EnvRec *state;               // [M] current environment
    EnvRec *timeline;            // [K][M] state after k = 1..K updates
    int32_t M, K;
    float dt;
    int32_t motion_type;         // rvs_motion_type
    float ws_center[3], ws_radius, base_radius, robot_max_vel;
    const float *q_prev, *q_cur; // [E][n]
    const float *points;         // [E][K][n] explicit validity points (MotionValidity::check(traj_points)); nullptr = interpolate q_prev -> q_cur
    int64_t E;
    int32_t *first_invalid;      // [E]
    int32_t commit_edge;
    float eq_thr, max_dist;
    ObstacleRec *obstacles, *obstacles_scan;
    float *scan_dist, *scan_key;
    void *obstacles_pack;        // fp16 copy of the scan-order table (re-packed after every re-ranking)
    // random_directions
    int32_t env_ground_included; // Environment::ground_included (read by Environment::isValid only)
    LibcRand *rand_state;        // the libc stream the law draws from
    LibcRand *rand_timeline;     // [K] its state after k = 1..K updates (commit adopts the one that matches the adopted obstacles)
};

}  // namespace rvs
