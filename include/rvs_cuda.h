/*
 * rvs_cuda.h — C ABI of the B200-native collision-and-distance engine that sits behind
 * RPMPLv2's `base::RealVectorSpace` (librvs_cuda.so).
 *
 * Every entry point below replaces one piece of the reference's single-threaded CPU hot path.
 * Citations are `path:line` inside the reference tree (robotics-ETF/RPMPLv2 @ 481d3ffb):
 *
 *   rvs_create / rvs_destroy          <- scenario::Scenario ctor, the `robot.space` switch
 *                                        (src/scenario/Scenario.cpp:105-112) + robot ctors
 *                                        (src/robots/xArm6.cpp:7-78, src/robots/Planar2DOF.cpp:7-64)
 *   rvs_set_obstacles                 <- env::Environment ctor / addObject
 *                                        (src/environments/Environment.cpp:13-62,116-131),
 *                                        env::Box ctor (src/environments/objects/Box.cpp:3-15)
 *   rvs_update_obstacle_aabbs         <- env::Object::setPosition + Environment::updateEnvironment
 *                                        (src/environments/Object.cpp:21-26, Environment.cpp:194-377)
 *   rvs_is_valid                      <- RealVectorSpace::isValid(q)
 *                                        (src/state_spaces/real_vector_space/RealVectorSpace.cpp:277-309)
 *   rvs_is_valid_edges                <- RealVectorSpace::isValid(q1,q2)        (RealVectorSpace.cpp:261-275)
 *   rvs_distance                      <- RealVectorSpace::computeDistance       (RealVectorSpace.cpp:318-392)
 *   rvs_underestimation               <- RealVectorSpace::computeDistanceUnderestimation (RealVectorSpace.cpp:402-446)
 *   rvs_skeleton                      <- AbstractRobot::computeSkeleton         (src/robots/xArm6.cpp:209-235,
 *                                        src/robots/Planar2DOF.cpp:117-130)
 *   rvs_enclosing_radii               <- AbstractRobot::computeEnclosingRadii   (src/robots/xArm6.cpp:237-275,
 *                                        Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30)
 *   rvs_self_collision_capsules       <- xArm6::checkSelfCollision(q) capsule pre-filter
 *                                        (src/robots/xArm6.cpp:374-424; FCL mesh stage is NOT covered)
 *   rvs_set_self_collision            <- AbstractRobot::setSelfCollisionChecking (src/scenario/Scenario.cpp:101-103)
 *   rvs_self_collision_edges          <- xArm6::checkSelfCollision(q1, q2&) (src/robots/xArm6.cpp:284-368), capsule stage
 *   rvs_capsule_box_distance /        <- the public static primitives CollisionAndDistance::distanceCapsuleToBox /
 *   rvs_capsule_box_collision            collisionCapsuleToBox (CollisionAndDistance.cpp:4-42, 208-214), batched
 *   rvs_store_distance /              <- base::State's distance cache (include/state_spaces/State.h:24-41) filled by
 *   rvs_store_extend_spines              computeDistance and consumed by RGBTConnect::extendGenSpine / extendGenSpine2
 *                                        (src/planners/rbt/RGBTConnect.cpp:73-118), kept resident in HBM
 *   rvs_env_set / rvs_env_step /      <- env::Environment's dynamic obstacles: updateEnvironment and its motion laws
 *   rvs_env_get                          (src/environments/Environment.cpp:194-377), env::Object::setPosition (Object.cpp:21-26)
 *   rvs_motion_validity               <- planning::trajectory::MotionValidity::check(q_previous, q_current)
 *                                        (src/planners/trajectory/MotionValidity.cpp:30-51), obstacles moving on the device
 *   rvs_motion_validity_points        <- MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-82): the validity
 *                                        points come from the caller's spline / Ruckig trajectory instead of an interpolated edge
 *   rvs_extend_spines                 <- RBTConnect::extendSpine + RGBTConnect::extendGenSpine over a set of
 *                                        targets (src/planners/rbt/RBTConnect.cpp:85-150,
 *                                        src/planners/rbt/RGBTConnect.cpp:73-118)
 *   rvs_check_states / rvs_check_edges <- the two composite tests the planners write by hand: isValid && !robot->checkSelfCollision
 *                                        (src/planners/rrt/RRTConnect.cpp:13-15, :100; trajectory/MotionValidity.cpp:42)
 *   rvs_store_connect                 <- RGBTConnect::connectGenSpine's loop / RBTConnect::connectSpine
 *                                        (src/planners/rbt/RGBTConnect.cpp:120-150, RBTConnect.cpp:152-180) as one call
 *   rvs_env_set_random_directions     <- Environment::moveRandomDirections' state: accelerations, ground_included, srand(seed)
 *                                        (src/environments/Environment.cpp:169-190, 289-333)
 *   rvs_tree_create / append /        <- base::Tree's states and getNearestState (src/state_spaces/Tree.cpp:35-50, :90-103),
 *   nearest / radius / release           RRTx::findNeighbors (src/planners/rrtx/RRTx.cpp:424-452), resident in HBM
 *   rvs_set_distance_field /          <- (no reference counterpart) result-neutral switches of the large-batch path
 *   rvs_set_config_sort / rvs_set_culling
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / torch types; all functions return an rvs_status (0 = ok);
 *    no exception crosses the boundary; a collision is data (`d_c == 0`, verdict byte 0), never an error.
 *  - configurations `q` are row-major  N x num_dofs  float32 (the reference's Eigen::VectorXf coord).
 *  - one handle = one device + one CUDA stream; calls on one handle must be serialised by the caller.
 *  - functions without suffix take HOST pointers and are synchronous (H2D, kernel, D2H inside).
 *    `_dev` functions take DEVICE pointers, enqueue on the handle's stream and return immediately;
 *    use rvs_sync() or your own event on rvs_get_stream().
 *  - there is no CPU fallback: without a usable CUDA device rvs_create fails with RVS_ERR_NO_DEVICE.
 */
#ifndef RVS_CUDA_H
#define RVS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVS_MAX_DOFS 32
#define RVS_ABI_VERSION 1

typedef enum {
    RVS_OK = 0,
    RVS_ERR_INVALID_ARGUMENT = 1,
    RVS_ERR_CUDA = 2,
    RVS_ERR_OUT_OF_MEMORY = 3,
    RVS_ERR_NO_DEVICE = 4,
    RVS_ERR_UNSUPPORTED = 5
} rvs_status;

/* Which reference robot class the chain stands for. It selects the skeleton recipe and the
 * enclosing-radii recurrence only; FK always comes from the joint chain in the descriptor. */
typedef enum {
    RVS_ROBOT_XARM6 = 0,   /* robots::xArm6      (src/robots/xArm6.cpp)        */
    RVS_ROBOT_PLANAR = 1,  /* robots::Planar2DOF / Planar10DOF (Planar2DOF.cpp) */
    RVS_ROBOT_SPATIAL = 2  /* robots::Spatial10/14/18/22DOF (Spatial10DOF.cpp)  */
} rvs_robot_kind;

/* Arithmetic variant of the hot path.
 *  REF32  : reference-faithful. FK in FP64 (as orocos-KDL does), skeleton narrowed to FP32, all capsule/box
 *           geometry in FP32 with NO fused multiply-add (the reference is an x86-64 -O3 build without FMA)
 *           and IEEE division / sqrt.  Verdicts and distances are bit-comparable with the oracle<float>.
 *  FAST32 : FP32 throughout (FK included), FMA contraction allowed. Stated tolerance 1e-4 m.
 *  F64    : everything in FP64, outputs through the *_f64 entry points. Tolerance 1e-9 m vs oracle<double>. */
typedef enum {
    RVS_PREC_REF32 = 0,
    RVS_PREC_FAST32 = 1,
    RVS_PREC_F64 = 2
} rvs_precision;

typedef enum {
    RVS_DIST_MIN = 0,      /* d_c only                                         (4 B / configuration)          */
    RVS_DIST_PROFILE = 1,  /* d_c + per-link profile d_c_profile[L]            (4(1+L) B)                     */
    RVS_DIST_FULL = 2      /* + nearest points [M][L][6] = (Rx,Ry,Rz,Ox,Oy,Oz) (+24 L M B), State::nearest_points */
} rvs_distance_mode;

/* Serial revolute chain as parsed from the robot's URDF (kdl_parser semantics):
 *   T_k = T_{k-1} * Trans(joint_xyz[k]) * RPY(joint_rpy[k]) * Rot(joint_axis[k], q_k),   k = 0 .. num_dofs-1
 * followed by one fixed tip transform at index num_dofs (xarm6: joint_eef; chains: the `tool` joint). */
typedef struct {
    int32_t kind;                               /* rvs_robot_kind */
    int32_t num_dofs;                           /* n == number of links L (AbstractRobot::getNumLinks) */
    int32_t ground_included;                    /* AbstractRobot::getGroundIncluded */
    float   gripper_length;                     /* xArm6 only (src/robots/xArm6.cpp:130,231) */
    double  joint_xyz[RVS_MAX_DOFS + 1][3];
    double  joint_rpy[RVS_MAX_DOFS + 1][3];
    double  joint_axis[RVS_MAX_DOFS][3];
    float   capsule_radius[RVS_MAX_DOFS];       /* AbstractRobot::getCapsuleRadius(i) */
    float   limit_lo[RVS_MAX_DOFS];             /* AbstractRobot::getLimits() */
    float   limit_hi[RVS_MAX_DOFS];
} rvs_robot_desc;

/* RealVectorSpaceConfig (src/configurations/RealVectorSpaceConfig.cpp:3-5 and
 * data/configurations/configuration_realvectorspace.yaml) made explicit. */
typedef struct {
    float   equality_threshold;                 /* EQUALITY_THRESHOLD (1e-6 compiled, 1e-4 shipped YAML) */
    float   max_distance;                       /* MAX_DISTANCE (1000) */
    int32_t num_interpolation_validity_checks;  /* NUM_INTERPOLATION_VALIDITY_CHECKS (15 compiled, 10 YAML) */
    int32_t precision;                          /* rvs_precision */
} rvs_params;

/* Parameters of the batched generalized-bur spine primitive (RBTConnectConfig / RGBTConnectConfig,
 * data/configurations/configuration_rbtconnect.yaml:4-8, configuration_rgbtconnect.yaml:4). */
typedef struct {
    int32_t num_layers;            /* RGBTConnectConfig::NUM_LAYERS (5) */
    int32_t num_iter_spine;        /* RBTConnectConfig::NUM_ITER_SPINE (1) */
    int32_t use_expanded_bubble;   /* RBTConnectConfig::USE_EXPANDED_BUBBLE (1) */
    float   d_crit;                /* RBTConnectConfig::D_CRIT (0.05) */
} rvs_spine_params;

typedef struct rvs_engine *rvs_handle;

/* ---- life cycle ---------------------------------------------------------------------------- */
int  rvs_abi_version(void);
int  rvs_device_count(void);
int  rvs_create(const rvs_robot_desc *robot, const rvs_params *params, int device, rvs_handle *out);
int  rvs_destroy(rvs_handle h);
const char *rvs_last_error(rvs_handle h);      /* h may be NULL: last error of a failed rvs_create */
int  rvs_set_params(rvs_handle h, const rvs_params *params);   /* precision < 0: keep the handle's variant */
int  rvs_get_params(rvs_handle h, rvs_params *out);
int  rvs_num_dofs(rvs_handle h);
int  rvs_get_limits(rvs_handle h, float *limit_lo, float *limit_hi);   /* num_dofs floats each */
int  rvs_set_stream(rvs_handle h, void *cuda_stream);  /* adopt an external cudaStream_t (e.g. torch's) */
void *rvs_get_stream(rvs_handle h);
int  rvs_sync(rvs_handle h);
/* kernels launched by this handle since creation (each counts one grid launch) */
int64_t rvs_launch_count(rvs_handle h);

/* Work accounting: pair queries (distanceCapsuleToBox calls) that RealVectorSpace::computeDistance's own loop order
 * (link-major, stop at the first collision, ground pairs skipped) would have made for all configurations passed to
 * rvs_distance* since the last reset. Synchronises the stream. */
int  rvs_pair_count(rvs_handle h, int32_t reset, uint64_t *out);
/* The same counter plus what the work-queue distance kernel actually executed since the last reset:
 * out4[0] = reference-order pair queries (as rvs_pair_count), out4[1] = pairs that survived the broad-phase scan,
 * out4[2] = pairs that went through the reference's case classification, out4[3] = segment-edge evaluations.
 * (The sequential kernel evaluates every visited pair: it only advances out4[0].) Synchronises the stream. */
int  rvs_work_stats(rvs_handle h, int32_t reset, uint64_t *out4);

/* Roofline denominator: achieved FFMA (fp64 = 0) or DFMA (fp64 = 1) throughput of this device in TFLOP/s, measured with a
 * register-only unrolled FMA kernel (best of 4 timed launches, CUDA events). */
int  rvs_measure_fma_peak(rvs_handle h, int32_t fp64, double *tflops);

/* ---- obstacle table (resident in HBM as SoA) -------------------------------------------------- */
/* aabb_min / aabb_max: M x 3 float (x,y,z); is_ground: M bytes (label == "ground"); min_dist_tol: M float
 * or NULL (= max_distance for all). Host pointers. */
int  rvs_set_obstacles(rvs_handle h, const float *aabb_min, const float *aabb_max,
                       const uint8_t *is_ground, const float *min_dist_tol, int32_t M);
int  rvs_update_obstacle_aabbs(rvs_handle h, const float *aabb_min, const float *aabb_max, int32_t M);
int  rvs_num_obstacles(rvs_handle h);

/* ---- hot path, host pointers, synchronous ------------------------------------------------------ */
/* valid[i] = 1 iff RealVectorSpace::isValid(q_i) */
int  rvs_is_valid(rvs_handle h, const float *q, int64_t N, uint8_t *valid);
/* valid[e] = 1 iff RealVectorSpace::isValid(q1_e, q2_e) (discretised edge, far end first) */
int  rvs_is_valid_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *valid);
/* d_c[N]; profile[N x L] (mode >= PROFILE, else may be NULL); nearest[N x M x L x 6] (mode == FULL).
 * After a collision (d_c == 0) the profile follows the reference's early-exit fill
 * (RealVectorSpace.cpp:370-378) and the nearest-point block of that configuration is unspecified
 * (large FULL batches do not touch it at all: csrc/rvs_distance_full.cuh computes the nearest points of the collision-free
 * configurations only; RVS_FULL_SPLIT=0 in the environment at rvs_create time selects the one-kernel path for A/B runs;
 * RVS_FULL_PARTS_TARGET=<blocks>, read per call, overrides how many obstacle ranges share a block of configurations — results do not depend on either). */
int  rvs_distance(rvs_handle h, const float *q, int64_t N, int32_t mode,
                  float *d_c, float *profile, float *nearest);
/* planes: R x M x L x 6 nearest-point sets of R root configurations; plane_idx[N] selects the root
 * per configuration (NULL = root 0 for all). status[i] = 1 when the reference would have returned 0
 * WITHOUT touching the state (a plane distance < 0, RealVectorSpace.cpp:428-429); d_c[i] is 0 then. */
int  rvs_underestimation(rvs_handle h, const float *q, int64_t N, const float *planes,
                         const int32_t *plane_idx, int32_t R, float *d_c, float *profile, uint8_t *status);
/* skeleton[N x (L+1) x 3] */
int  rvs_skeleton(rvs_handle h, const float *q, int64_t N, float *skeleton);
/* radii[N x L x (L+1)] row-major (row = start skeleton point, column = end point) */
int  rvs_enclosing_radii(rvs_handle h, const float *q, int64_t N, float *radii);
/* xArm6 capsule pre-filter of checkSelfCollision(q): bit0 = pair 0-5, bit1 = 1-5, bit2 = 0-4, bit3 = 1-4 set
 * when computeCapsulesDistance < 0 ("maybe": the reference then runs an FCL mesh test, which is out of scope).
 * For every other robot kind the reference's answer is the constant `false` -> all zero. */
int  rvs_self_collision_capsules(rvs_handle h, const float *q, int64_t N, uint8_t *mask);
/* AbstractRobot::setSelfCollisionChecking (src/scenario/Scenario.cpp:101-103; robots::xArm6's constructor default is true,
 * src/robots/xArm6.cpp:67) plus RRTConnectConfig::EPS_STEP, the marching step of xArm6::checkSelfCollision(q1, q2)
 * (src/robots/xArm6.cpp:340-341; data/configurations/configuration_rrtconnect.yaml:4 = 0.1).  A new handle has the check OFF.
 * While it is on, rvs_extend_spines / rvs_store_extend_spines apply the reference's truncation
 * (src/planners/rbt/RBTConnect.cpp:129-131).  Only xArm6 has a self-collision check at all (the other classes return false). */
int  rvs_set_self_collision(rvs_handle h, int32_t enabled, float eps_step);
int  rvs_get_self_collision(rvs_handle h, int32_t *enabled, float *eps_step);
/* xArm6::checkSelfCollision(q1, q2&) (src/robots/xArm6.cpp:284-368) for E motions q1[e] -> q2[e]: up to five enclosing-radii
 * bubble steps, then EPS_STEP marching with the capsule test of the pairs {0-5, 1-5, 0-4, 1-4}.  hit[e] = 1 on a self-collision;
 * q2_out[E x n] = q2 after the call (the last collision-free configuration on a hit, q2 itself otherwise).
 * Capsule stage only: the FCL mesh confirmation (xArm6.cpp:426-464) is host-side and out of contract, so a hit means "maybe"
 * (conservative).  hit = 0 / q2_out = q2 when the check is off or the robot is not an xArm6. */
int  rvs_self_collision_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *hit, float *q2_out);
int  rvs_self_collision_edges_dev(rvs_handle h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_hit, float *d_q2_out);
/* isValid(q1, q2) and robot->checkSelfCollision(q1, q2) in ONE call, reported apart: a caller whose control flow depends on WHICH test failed
 * (RRTx::chooseParent re-seats q_new on a self hit, src/planners/rrtx/RRTx.cpp:466-468) gets both verdicts with one synchronisation. */
int  rvs_check_edges_detailed(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out);
/* Device-resident planner trees: the states of base::Tree kept as a [count][n] array in HBM, answering
 *   rvs_tree_nearest  <- base::Tree::getNearestState (src/state_spaces/Tree.cpp:35-50; nanoflann kd-tree, exact 1-NN in the squared
 *                        L2 metric): Q queries in ONE launch, query k against tree trees[k] (NULL = tree 0) -> idx[k] (-1 for an empty
 *                        tree) and the squared distance; brute force, ties to the lowest index
 *   rvs_tree_radius   <- planning::rrtx::RRTx::findNeighbors (src/planners/rrtx/RRTx.cpp:424-452): every state within `radius` of the
 *                        query in ss->getNorm, in TREE ORDER, except index `exclude`; *count = number of hits (when it exceeds `capacity`
 *                        only the first `capacity` are written: call again with more room)
 * Trees are append-only arrays (rvs_tree_clear empties one, rvs_tree_release frees its memory and lets rvs_tree_create hand the id out again). */
int  rvs_tree_create(rvs_handle h, int32_t *tree);
int  rvs_tree_release(rvs_handle h, int32_t tree);
int  rvs_tree_clear(rvs_handle h, int32_t tree);
int  rvs_tree_size(rvs_handle h, int32_t tree);
int  rvs_tree_append(rvs_handle h, int32_t tree, const float *q, int32_t count);
int  rvs_tree_nearest(rvs_handle h, const int32_t *trees, const float *queries, int32_t Q, int32_t *idx, float *dist2);
int  rvs_tree_radius(rvs_handle h, int32_t tree, const float *query, float radius, int32_t exclude, int32_t capacity, int32_t *idx,
                     float *dist, int32_t *count);
/* The two composite tests every planner of the reference writes out by hand, as ONE call (one upload, one synchronisation):
 *   rvs_check_states: valid[i] = ss->isValid(q_i) && !ss->robot->checkSelfCollision(q_i)
 *                     (src/planners/rrt/RRTConnect.cpp:13-15, trajectory/MotionValidity.cpp:42, rbt_star/RGBMTStar.cpp:202)
 *   rvs_check_edges:  valid[e] = ss->isValid(q1_e, q2_e) && !ss->robot->checkSelfCollision(q1_e, q2_e)
 *                     (src/planners/rrt/RRTConnect.cpp:100, rrtx/RRTx.cpp:467-468, 507-508)
 * Identical to rvs_is_valid / rvs_is_valid_edges while the self-collision check is off (rvs_set_self_collision). */
int  rvs_check_states(rvs_handle h, const float *q, int64_t N, uint8_t *valid);
int  rvs_check_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *valid);
/* Generalized-bur spines from ONE root towards S targets (RGBTConnect::extendGenSpine with
 * underestimation). q_root[L dofs], root profile/planes come from a FULL distance query of the root.
 * q_reached[S x n]; d_c_reached[S]; status[S] (base::State::Status: 1 Advanced, 2 Trapped, 3 Reached). */
int  rvs_extend_spines(rvs_handle h, const float *q_root, const float *root_profile, float root_d_c,
                       const float *root_planes, const float *q_targets, int32_t S,
                       const rvs_spine_params *sp, float *q_reached, float *d_c_reached, uint8_t *status);

/* ---- device-resident State cache (what the reference keeps on base::State: d_c, d_c_profile, nearest_points;
 * include/state_spaces/State.h:24-41) ------------------------------------------------------------------------------
 * The planners grow every spine of a bur from data that computeDistance left on the root State. Keeping those
 * 24*L*M bytes per state in HBM "slots" means a planner iteration moves only configurations and scalars over PCIe.
 * Slots are numbered 0 .. capacity-1 and are owned by the caller (one slot per tree node that served as a bur root).
 * The cache is FP32 (REF32 / FAST32 handles). rvs_set_obstacles with a different M drops the cache; after
 * rvs_update_obstacle_aabbs the cached planes are stale exactly as a reference State's are after
 * Environment::updateEnvironment (the planners then ask again with compute_again = true). */
int  rvs_store_reserve(rvs_handle h, int32_t capacity);          /* grows (contents preserved), never shrinks */
int  rvs_store_capacity(rvs_handle h);
/* computeDistance(q_i) with nearest points (RealVectorSpace.cpp:318-392) for N configurations; everything is
 * written to slot slots[i]; d_c[N] (and profile[N x L] unless NULL) also come back to the host. */
int  rvs_store_distance(rvs_handle h, const float *q, int32_t N, const int32_t *slots, float *d_c, float *profile);
/* Generalized-bur spines whose roots are cached states: spine s starts at slot root_slots[s] towards
 * q_targets[s]. gen2 = 0: RGBTConnect::extendGenSpine (RGBTConnect.cpp:73-92); gen2 = 1: extendGenSpine2
 * (:96-118). q_layers[S x num_layers x n] receives the state after every layer, n_layers[S] how many were made
 * (0 when the root has no clearance: status Trapped), d_c_last[S] the last distance-to-planes, status[S] the
 * base::State::Status of the last extendSpine (1 Advanced, 2 Trapped, 3 Reached). */
int  rvs_store_extend_spines(rvs_handle h, const int32_t *root_slots, const float *q_targets, int32_t S,
                             const rvs_spine_params *sp, int32_t gen2, float *q_layers, int32_t *n_layers,
                             float *d_c_last, uint8_t *status);
/* connectGenSpine's loop (src/planners/rbt/RGBTConnect.cpp:120-150; RBTConnect::connectSpine :152-180 with one layer) for R requests in
 * ONE call: request r starts from the cached state root_slots[r] and repeats, at most `steps` times,
 *     { extendGenSpine2 towards q_targets[r]; computeDistance of the state it reached, cached in new_slots[r][k] }
 * while the spine Advanced and the current state's d_c is above d_crit (below it the reference marches in RRT mode: the caller's
 * business).  new_slots is R x steps (-1 = this request makes fewer steps).  Outputs, R x steps each: q_layers (x num_layers x num_dofs),
 * n_layers (-1: the step was not executed — the loop had ended), d_c_new (real distance of the reached state), status
 * (base::State::Status of the spine).  The 2 x steps kernels gate themselves on device state: one synchronisation per loop. */
int  rvs_store_connect(rvs_handle h, const int32_t *root_slots, const float *q_targets, const int32_t *new_slots, int32_t R, int32_t steps,
                       const rvs_spine_params *sp, float *q_layers, int32_t *n_layers, float *d_c_new, uint8_t *status);
/* copy one slot back to the host (any of the outputs may be NULL); planes[M x L x 6] */
int  rvs_store_read(rvs_handle h, int32_t slot, float *q, float *d_c, float *profile, float *planes);

/* ---- dynamic environment on the device ------------------------------------------------------
 * env::Environment keeps position / velocity / max_vel per object and moves every object labelled "dynamic_obstacle" by one of
 * its motion laws in updateEnvironment(delta_time); MotionValidity::check interleaves K such updates with K validity points of
 * the robot's motion.  Behind a device engine that is K obstacle uploads + K tiny launches per planner iteration.  These entry
 * points keep the obstacle STATES resident and do all K steps in one call. */
typedef enum {
    RVS_MOTION_STRAIGHT = 0,           /* Environment::moveStraight          (Environment.cpp:224-238) */
    RVS_MOTION_CIRCULAR = 1,           /* Environment::moveCircular          (:241-258) */
    RVS_MOTION_TWO_TUNNELS = 2,        /* Environment::moveTwoTunnels        (:261-289) */
    RVS_MOTION_RANDOM_DIRECTIONS = 3,  /* Environment::moveRandomDirections  (:289-333); see rvs_env_set_random_directions */
    RVS_MOTION_LIGHT_DIRECTIONS = 4    /* Environment::moveLightDirections   (:336-377) */
} rvs_motion_type;

typedef struct {
    int32_t motion_type;               /* one of rvs_motion_type: scenario YAML testing.motion_type */
    float   ws_center[3];              /* robot.WS_center */
    float   ws_radius;                 /* robot.WS_radius */
    float   base_radius;               /* Environment::setBaseRadius */
    float   robot_max_vel;             /* Environment::setRobotMaxVel */
} rvs_env_params;

/* Defines the environment AND the resident obstacle table (it replaces rvs_set_obstacles): M boxes with dim / pos (M x 3),
 * rotated[j] != 0 when the box has a non-identity rotation (FCL then reports the bounding-sphere cube; NULL = none),
 * vel (M x 3, NULL = zero), max_vel (M, NULL = zero), is_dynamic (label == "dynamic_obstacle"), is_ground, min_dist_tol as in
 * rvs_set_obstacles.  Environment::sign starts at +1 and path_len at 0 (Environment.cpp:76-77). */
int  rvs_env_set(rvs_handle h, const rvs_env_params *ep, const float *dim, const float *pos, const uint8_t *rotated,
                 const float *vel, const float *max_vel, const uint8_t *is_dynamic, const uint8_t *is_ground,
                 const float *min_dist_tol, int32_t M);
/* Extra state of Environment::moveRandomDirections (Environment.cpp:289-333), set after rvs_env_set:
 *   acc             M x 3 object accelerations (env::Object::acceleration; NULL = keep, rvs_env_set starts them at zero like
 *                   Environment::addObject's default argument)
 *   ground_included Environment::ground_included as Environment::isValid reads it (< 0 = keep; rvs_env_set takes the robot's)
 *   seed            the law draws its new directions from Eigen::Vector3f::Random, i.e. libc rand(), one stream for all objects in index
 *                   order.  The engine replays that stream on the device (glibc's rand(): additive feedback r[i] = r[i-3] + r[i-31]):
 *                   seed >= 0 = srand(seed); < 0 = keep the stream where it is (a fresh handle starts like a process that never
 *                   called srand, i.e. srand(1)).  After srand(seed) on both sides the objects move exactly as in the reference.
 * rvs_env_get_random_directions: current accelerations (may be NULL) and the number of rand() calls consumed since the last seed, so
 * a host that shares its libc stream with the environment can advance its own by as many draws. */
int  rvs_env_set_random_directions(rvs_handle h, const float *acc, int32_t ground_included, int64_t seed);
int  rvs_env_get_random_directions(rvs_handle h, float *acc, uint32_t *draws);
/* `steps` x Environment::updateEnvironment(delta_time) on the device; the resident obstacle table follows. */
int  rvs_env_step(rvs_handle h, float delta_time, int32_t steps);
/* current object positions / velocities (M x 3 each, either may be NULL) */
int  rvs_env_get(rvs_handle h, float *pos, float *vel);
/* MotionValidity::check(q_previous_e, q_current_e) for E motions that share ONE obstacle timeline of K updates of delta_time:
 * validity point k (k = 1..K, interpolateEdge(q_previous, q_current, dist * k / K, dist)) is tested against the environment
 * after k updates.  valid[e] = the reference's return value, first_invalid[e] = k of the first invalid point (0 if valid).
 * commit_edge >= 0: the environment is left exactly as the reference leaves it after checking THAT motion (updated
 * first_invalid times — the reference returns at once — or K times); commit_edge < 0: the environment is not changed
 * (what-if queries for many candidate motions).  Self-collision (xArm6 meshes, FCL) is not part of this check. */
int  rvs_motion_validity(rvs_handle h, const float *q_prev, const float *q_cur, int64_t E, int32_t K, float delta_time,
                         int32_t commit_edge, uint8_t *valid, int32_t *first_invalid);
/* MotionValidity::check(traj_points_current_iter) for E candidate trajectories of K points each (points: E x K x num_dofs): point k
 * is tested, as it is, against the environment after k updates of delta_time.  Same outputs and commit rule as above. */
int  rvs_motion_validity_points(rvs_handle h, const float *points, int64_t E, int32_t K, float delta_time,
                                int32_t commit_edge, uint8_t *valid, int32_t *first_invalid);

/* Batched geometric primitives on K independent (capsule, box) pairs, host pointers.
 * AB: K x 6 (A xyz, B xyz); radius: K; obs: K x 6 (x_min,y_min,z_min,x_max,y_max,z_max).
 * distance -> out: K x 8 rows (d, has_points, P1 xyz, P2 xyz), has_points = 0 where the reference returns nullptr;
 * collision -> hit: K bytes. Uses the handle's precision (REF32 / FAST32 only) and thresholds. */
int  rvs_capsule_box_distance(rvs_handle h, const float *AB, const float *radius, const float *obs, int64_t K, float *out);
int  rvs_capsule_box_collision(rvs_handle h, const float *AB, const float *radius, const float *obs, int64_t K, uint8_t *hit);

/* ---- FP64 variants (handle must have been created with RVS_PREC_F64) ---------------------------- */
int  rvs_distance_f64(rvs_handle h, const float *q, int64_t N, int32_t mode,
                      double *d_c, double *profile, double *nearest);
int  rvs_is_valid_f64(rvs_handle h, const float *q, int64_t N, uint8_t *valid);

/* ---- device-pointer, stream-ordered variants ---------------------------------------------------- */
int  rvs_is_valid_dev(rvs_handle h, const float *d_q, int64_t N, uint8_t *d_valid);
int  rvs_is_valid_edges_dev(rvs_handle h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_valid);
int  rvs_distance_dev(rvs_handle h, const float *d_q, int64_t N, int32_t mode,
                      float *d_d_c, float *d_profile, float *d_nearest);
int  rvs_distance_f64_dev(rvs_handle h, const float *d_q, int64_t N, int32_t mode,
                          double *d_d_c, double *d_profile, double *d_nearest);
int  rvs_underestimation_dev(rvs_handle h, const float *d_q, int64_t N, const float *d_planes,
                             const int32_t *d_plane_idx, int32_t R, float *d_d_c, float *d_profile,
                             uint8_t *d_status);
/* Stateless synthetic configurations: q[i][j] = lo_j + (hi_j-lo_j) * u24(splitmix64(seed ^ (i*32+j)))
 * generated on the device for i in [first, first+N). */
int  rvs_generate_configs_dev(rvs_handle h, uint64_t seed, int64_t first, int64_t N, float *d_q);

/* ---- tuning knobs (do not change results) ------------------------------------------------------- */
/* cull != 0 enables the conservative bounding-box pre-test in MIN/PROFILE distance queries: a pair whose
 * lower bound cannot beat the running per-link minimum skips the exact narrow phase. Results are identical. */
int  rvs_set_culling(rvs_handle h, int32_t cull);
/* on != 0 (default): large MIN/PROFILE distance batches (>= 32768 configurations, 32 <= M < 65536) start every link's running minimum at
 * an upper bound read from a coarse distance field of the obstacle table (128^3 fp16 cells around the robot, rebuilt on the device when
 * the table changed), instead of at MAX_DISTANCE: the first obstacles of the table no longer pass the broad phase for every link.  A link
 * for which no candidate came below its bound is re-scanned without bounds, so results are identical with the switch on or off.
 * rvs_second_pass_warps: how many warps (of 32 configurations) needed that second pass since the last reset. */
int  rvs_set_distance_field(rvs_handle h, int32_t on);
/* on != 0 (default): batches of >= 65536 configurations (distance MIN / PROFILE, isValid, edges) are processed in an order that puts
 * similar configurations into the same warp (a counting sort on the leading joints, on the device; results are written back in the
 * caller's order).  Per-configuration results do not depend on it. */
int  rvs_set_config_sort(rvs_handle h, int32_t on);
int  rvs_second_pass_warps(rvs_handle h, int32_t reset, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* RVS_CUDA_H */
