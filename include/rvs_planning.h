/*
 * rvs_planning.h — C ABI of librvs_planning.so: RPMPLv2's RRT-Connect / RBT-Connect / RGBT-Connect driven through the
 * BATCHED state-space queries of librvs_cuda.so (include/rvs_cuda.h), P independent planner runs in lockstep.
 *
 * Reference interfaces this stands in for (robotics-ETF/RPMPLv2 @ 481d3ffb):
 *   planning::rrt::RRTConnect::solve      src/planners/rrt/RRTConnect.cpp:38-90
 *   planning::rbt::RBTConnect::solve      src/planners/rbt/RBTConnect.cpp:14-65
 *   planning::rbt::RGBTConnect::solve     src/planners/rbt/RGBTConnect.cpp:14-71
 *   the repetition loop of the test apps  apps/test_rgbtconnect.cpp (scenario `testing: max_num` runs of one scenario)
 * Configuration values are those of data/configurations/configuration_{rrtconnect,rbtconnect,rgbtconnect}.yaml.
 * Plain C: pointers and sizes only, status ints, no exception crosses the boundary.
 *
 * Environment switches (read per call): RVS_PLAN_THREADS = host threads that resume the runs between backend calls;
 * RVS_PLAN_CONNECT_CHUNK = K > 0: K steps of connectSpine / connectGenSpine's loop go to the engine as one rvs_store_connect call (default 0 =
 * one spine call and one distance call per step; runs are identical either way, and on the shipped scenes the chain does not pay);
 * RVS_PLAN_PROFILE = print the pool's time split (gather / backend / scatter + resume) to stderr;
 * RVS_PLAN_RNG=libc = draw random states from std::rand() after srand(seed), one call per coordinate like Eigen::VectorXf::Random —
 * a SINGLE run (P = 1) then follows the reference planner's own random stream and reproduces its run state by state
 * (tests/test_planners.py checks exactly that against the reference's compiled solve() loops). Default: a counter-based stream
 * per run, independent of P and of the thread schedule.
 */
#ifndef RVS_PLANNING_H
#define RVS_PLANNING_H

#include <stdint.h>

#include "rvs_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { RVSPLAN_RRT_CONNECT = 0, RVSPLAN_RBT_CONNECT = 1, RVSPLAN_RGBT_CONNECT = 2 } rvsplan_kind;

typedef struct {
    int32_t kind;                  /* rvsplan_kind */
    int64_t max_num_iter;          /* MAX_NUM_ITER */
    int64_t max_num_states;        /* MAX_NUM_STATES */
    float   max_planning_time;     /* MAX_PLANNING_TIME [s] */
    float   eps_step;              /* RRTConnectConfig::EPS_STEP (0.1) */
    int32_t max_extension_steps;   /* RRTConnectConfig::MAX_EXTENSION_STEPS (50) */
    float   delta;                 /* RBTConnectConfig::DELTA (3.14159) */
    int32_t num_spines;            /* RBTConnectConfig::NUM_SPINES (7) */
    rvs_spine_params spine;        /* NUM_LAYERS (5), NUM_ITER_SPINE (1), USE_EXPANDED_BUBBLE (1), D_CRIT (0.05) */
} rvsplan_config;

typedef struct {
    int32_t success;
    int32_t invalid_ends;          /* 1: start or goal in collision (the reference throws std::domain_error, RRTConnect.cpp:13-16) */
    int64_t iterations;
    int64_t states;
    float   planning_time;         /* [s] from the start of the call to the end of this run */
    int32_t path_len;              /* states on the path (0 when no path) */
    int64_t rounds;                /* batched backend calls this run took part in */
    int64_t num_valid, num_edges, num_distances, num_spines;   /* state-space queries it issued */
} rvsplan_result;

void rvsplan_default_config(int32_t kind, rvsplan_config *out);

/* P runs on the scene resident in `h`. q_start / q_goal: P x num_dofs; seeds: P (NULL = 0..P-1).
 * results: P. paths (may be NULL): P x max_path x num_dofs, run p's path in its first results[p].path_len rows
 * (truncated to max_path). pool_stats (may be NULL): [0] lockstep rounds, [1] batched backend calls, [2] kernels launched.
 * Returns an rvs_status; the message of a failure is available through rvs_last_error(h) or rvsplan_last_error(). */
int rvsplan_solve_batch(rvs_handle h, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds,
                        int32_t P, rvsplan_result *results, float *paths, int32_t max_path, int64_t *pool_stats);
/* The same call, also returning both trees of every run the way RBTConnect::outputPlannerData dumps them (`output_file << *trees[0]`,
 * `*trees[1]`; src/planners/rbt/RBTConnect.cpp:203-253, base::Tree's operator<<, src/state_spaces/Tree.cpp:155-162): for run p and tree t
 * (0 = from q_start, 1 = from q_goal) tree_sizes[2p + t] states in tree order, coordinates at tree_coord[(2p + t) * max_tree_states ...]
 * (x num_dofs) and the parent's index inside the same tree (-1 for the root) at tree_parent[(2p + t) * max_tree_states ...]; a tree
 * larger than max_tree_states is truncated (its size is still reported). */
int rvsplan_solve_batch_trees(rvs_handle h, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds,
                              int32_t P, rvsplan_result *results, float *paths, int32_t max_path, int64_t *pool_stats,
                              float *tree_coord, int32_t *tree_parent, int32_t *tree_sizes, int32_t max_tree_states);
const char *rvsplan_last_error(void);

/* ---- RRTx's edge-batch consumers ----------------------------------------------------------------------------
 * Reference: planning::rrtx::RRTx::solve's tree-growth block  src/planners/rrtx/RRTx.cpp:157-190
 *            findNeighbors :424-452, chooseParent :454-491, rewireNeighbors :493-529, propagateCostChanges :531-548,
 *            updatePath :550-572, computePath :575-592, shrinkingBallRadius :636-639; configuration_rrtx.yaml.
 * rvsplan_rrtx_grow builds the tree the reference's loop builds from `q_root` (RRTx roots its single tree at the goal) when its
 * random states are `samples` [S x num_dofs], in that order: per sample  nearest state -> extend by EPS_STEP -> findNeighbors ->
 * chooseParent -> upgradeTree -> rewireNeighbors.  Every candidate edge of an iteration (1 + up to 2 * MAX_NEIGHBORS edges) is ONE
 * batched rvs_check_edges call; the reference's sequential cost logic is replayed over the verdicts, so parents and costs are the
 * reference's, state by state (tests/test_rrtx.py, against RRTx.cpp compiled unmodified).
 * Outputs (capacity max_states rows; the call fails with RVS_ERR_INVALID_ARGUMENT when the tree outgrows it): coord [states x
 * num_dofs], parent [states] (-1: none), cost [states]; new_index [S] = index of the state sample s added, -1 when Trapped.
 * path_start >= 0: also walk updatePath / computePath from that state: path [<= max_states] = state indices from path_start to the
 * root (path_len = 0 when path_start has no parent), *path_invalid = 1 when any edge of the chain is blocked (one batched query). */
typedef struct {
    float   eps_step;                 /* RRTxConfig::EPS_STEP (yaml: 1.0) */
    float   r_rewire;                 /* RRTxConfig::R_REWIRE (yaml: 3.14); 0 = shrinking ball */
    int32_t max_neighbors;            /* RRTxConfig::MAX_NEIGHBORS (1000) */
    int32_t one_call_per_iteration;   /* 1: all candidate edges of an iteration in one engine call; 0: three calls with fewer edges */
    int32_t device_index;             /* 1: nearest / radius queries on the device-resident tree (rvs_tree_*); 0: host scans */
} rvsplan_rrtx_config;

typedef struct {
    int64_t states, iterations, trapped;
    int64_t backend_calls;            /* batched state-space calls issued */
    int64_t edges_checked;            /* edges sent to the state space */
    int64_t edges_reference;          /* edge checks the reference's one-at-a-time loop makes for the same tree */
    int64_t reparented, rewired;      /* q_new that took a parent other than q_near; neighbours re-parented to q_new */
    int64_t reseated;                 /* chooseParent loops in which a self-collision hit re-seated q_new (xArm6.cpp:284-352) */
} rvsplan_rrtx_stats;

void rvsplan_rrtx_default_config(rvsplan_rrtx_config *out);
int rvsplan_rrtx_grow(rvs_handle h, const rvsplan_rrtx_config *cfg, const float *q_root, const float *samples, int32_t S, int32_t max_states,
                      float *coord, int32_t *parent, float *cost, int32_t *new_index, int32_t path_start, int32_t *path, int32_t *path_len,
                      int32_t *path_invalid, rvsplan_rrtx_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* RVS_PLANNING_H */
