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
const char *rvsplan_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RVS_PLANNING_H */
