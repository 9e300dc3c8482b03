// oracle_planning.cpp — the planners of rpmplv2_b200/cpp/planning/ driven by the CPU ORACLE instead of the CUDA engine.
// TEST INFRASTRUCTURE ONLY: it is the checker of tests/test_planners.py and the CPU baseline of bench.py
// (`planning.cpu`); the product library (librvs_planning.so) does not contain or link any of this.
// Same planner sources, same random streams: the only difference between the two arms is who answers the
// state-space queries — orc::is_valid / is_valid_edge / distance / extend_gen_spine_layers here (the restatement of
// RealVectorSpace.cpp:261-446 and RGBTConnect.cpp:73-118), librvs_cuda.so there.
#include <chrono>
#include <cstring>
#include <exception>
#include <mutex>
#include <string>

#include "../rpmplv2_b200/cpp/planning/Planners.h"
#include "../rpmplv2_b200/cpp/planning/planning_capi.h"
#include "rvs_oracle_engine.hpp"

namespace {

thread_local std::string g_error;

class OracleBackend : public rvsplan::SpaceBackend {
public:
    explicit OracleBackend(orc_capi::Engine &e_) : e(e_), cfg(orc_capi::config_of<float>(e_.prm)), n(e_.robot.n), M(e_.sc32.boxes.size()) {}
    const char *name() const override { return "RealVectorSpace (CPU oracle)"; }
    int numDofs() const override { return n; }
    const float *limitsLo() const override { return e.lo.data(); }
    const float *limitsHi() const override { return e.hi.data(); }
    float equalityThreshold() const override { return e.prm.equality_threshold; }
    void reserveSlots(int capacity) override {
        if (size_t(capacity) > slots.size()) slots.resize(size_t(capacity));
    }
    void isValid(const float *q, int N, uint8_t *valid) override {
        std::vector<orc::V3<float>> sk(n + 1);
        for (int i = 0; i < N; i++) {
            orc::skeleton(e.robot, q + size_t(i) * n, sk.data());
            valid[i] = orc::is_valid(e.robot, e.sc32, sk.data(), cfg) ? 1 : 0;
            if (valid[i] && e.robot.self_collision_checking && orc::self_collision_mask(e.robot, sk.data(), cfg)) valid[i] = 0;      // RRTConnect.cpp:13-15
        }
    }
    void isValidEdges(const float *q1, const float *q2, int E, uint8_t *valid) override {
        std::vector<float> end(static_cast<size_t>(n), 0.f);
        for (int i = 0; i < E; i++) {
            valid[i] = orc::is_valid_edge(e.robot, e.sc32, q1 + size_t(i) * n, q2 + size_t(i) * n, cfg) ? 1 : 0;
            if (!valid[i]) continue;
            std::copy_n(q2 + size_t(i) * n, n, end.data());                                   // RRTConnect.cpp:100 (the truncated q2 is discarded)
            if (orc::self_collision_edge(e.robot, cfg, q1 + size_t(i) * n, end.data())) valid[i] = 0;
        }
    }
    bool hasSelfCollision() const override { return e.robot.self_collision_checking && e.robot.kind == orc::XARM6; }
    void isValidEdgesDetailed(const float *q1, const float *q2, int E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out) override {
        for (int i = 0; i < E; i++) {
            obstacle_ok[i] = orc::is_valid_edge(e.robot, e.sc32, q1 + size_t(i) * n, q2 + size_t(i) * n, cfg) ? 1 : 0;
            std::copy_n(q2 + size_t(i) * n, n, q2_out + size_t(i) * n);
            self_hit[i] = orc::self_collision_edge(e.robot, cfg, q1 + size_t(i) * n, q2_out + size_t(i) * n) ? 1 : 0;
        }
    }
    void distanceToSlots(const float *q, int N, const int32_t *slot_idx, float *d_c) override {
        std::vector<orc::V3<float>> sk(n + 1);
        for (int i = 0; i < N; i++) {
            Slot &s = slots[size_t(slot_idx[i])];
            s.q.assign(q + size_t(i) * n, q + size_t(i + 1) * n);
            s.profile.resize(size_t(n));
            s.planes.resize(M * size_t(n) * 6);
            orc::skeleton(e.robot, s.q.data(), sk.data());
            bool hit;
            s.d_c = orc::distance(e.robot, e.sc32, sk.data(), cfg, s.profile.data(), s.planes.data(), &hit);
            d_c[i] = s.d_c;
        }
    }
    void extendGenSpines(const int32_t *root_slots, const float *q_targets, int S, const rvsplan::SpineParams &sp, bool gen2, float *q_layers,
                         int32_t *n_layers, float *d_c_last, uint8_t *status) override {
        orc::SpineConfig sc;
        sc.num_layers = sp.num_layers; sc.num_iter_spine = sp.num_iter_spine; sc.use_expanded_bubble = sp.use_expanded_bubble; sc.d_crit = sp.d_crit;
        for (int s = 0; s < S; s++) {
            const Slot &r = slots[size_t(root_slots[s])];
            int layers = 0;
            status[s] = uint8_t(orc::extend_gen_spine_layers(e.robot, M, cfg, sc, gen2, r.q.data(), r.profile.data(), r.d_c, r.planes.data(),
                                                             q_targets + size_t(s) * n, q_layers + size_t(s) * size_t(sp.num_layers) * n, &layers,
                                                             d_c_last + s));
            n_layers[s] = layers;
        }
    }

private:
    struct Slot { std::vector<float> q, profile, planes; float d_c = -1.f; };
    orc_capi::Engine &e;
    orc::Config<float> cfg;
    int n;
    size_t M;
    std::vector<Slot> slots;
};

}  // namespace

extern "C" {

const char *orcplan_last_error(void) { return g_error.c_str(); }

// The CPU arm. lockstep != 0: ONE pool over all P runs (what the CUDA arm does; single thread). lockstep == 0: the natural CPU way —
// the P runs are split over `threads` std::threads, every run is its own pool, so planning_time is that run's own time.
int orcplan_solve_batch(void *orc_handle, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds, int32_t P,
                        rvsplan_result *results, float *paths, int32_t max_path, int32_t threads, int32_t lockstep, int64_t *pool_stats) {
    if (!orc_handle || !cfg || !q_start || !q_goal || !results || P < 0) { g_error = "orcplan_solve_batch: bad argument"; return 1; }
    auto &eng = *static_cast<orc_capi::Engine *>(orc_handle);
    const int n = eng.robot.n;
    const rvsplan::PlannerConfig pc = rvsplan::config_from_c(*cfg);
    try {
        if (lockstep) {
            OracleBackend backend(eng);
            rvsplan::PlannerPool pool(backend, pc);
            const auto res = pool.solve(q_start, q_goal, seeds, P, true);
            rvsplan::results_to_c(res, n, results, paths, max_path);
            if (pool_stats) { pool_stats[0] = pool.rounds(); pool_stats[1] = pool.backendCalls(); pool_stats[2] = 0; }
            return 0;
        }
        std::string first_error;
        std::mutex mu;
        orc_capi::parallel_for(P, threads, [&](int64_t b, int64_t en) {
            try {
                for (int64_t p = b; p < en; p++) {
                    OracleBackend backend(eng);
                    rvsplan::PlannerPool pool(backend, pc);
                    const uint64_t seed = seeds ? seeds[p] : uint64_t(p);
                    const auto res = pool.solve(q_start + size_t(p) * n, q_goal + size_t(p) * n, &seed, 1, true);
                    rvsplan::results_to_c(res, n, results + p, paths ? paths + size_t(p) * size_t(max_path) * n : nullptr, max_path);
                }
            } catch (const std::exception &ex) {
                std::lock_guard<std::mutex> lock(mu);
                if (first_error.empty()) first_error = ex.what();
            }
        });
        if (!first_error.empty()) { g_error = first_error; return 2; }
        if (pool_stats) { pool_stats[0] = pool_stats[1] = pool_stats[2] = 0; }
    } catch (const std::exception &ex) {
        g_error = ex.what();
        return 2;
    }
    return 0;
}


// one run (P = 1 per pool, single thread) with both trees returned: the CPU twin of rvsplan_solve_batch_trees
int orcplan_solve_batch_trees(void *orc_handle, const rvsplan_config *cfg, const float *q_start, const float *q_goal, const uint64_t *seeds, int32_t P,
                              rvsplan_result *results, float *paths, int32_t max_path, float *tree_coord, int32_t *tree_parent, int32_t *tree_sizes,
                              int32_t max_tree_states) {
    if (!orc_handle || !cfg || !q_start || !q_goal || !results || P < 0 || !tree_coord || !tree_parent || !tree_sizes) { g_error = "orcplan_solve_batch_trees: bad argument"; return 1; }
    auto &eng = *static_cast<orc_capi::Engine *>(orc_handle);
    const int n = eng.robot.n;
    rvsplan::PlannerConfig pc = rvsplan::config_from_c(*cfg);
    pc.keep_trees = true;
    try {
        OracleBackend backend(eng);
        rvsplan::PlannerPool pool(backend, pc);
        const auto res = pool.solve(q_start, q_goal, seeds, P, true);
        rvsplan::results_to_c(res, n, results, paths, max_path);
        rvsplan::trees_to_c(res, n, tree_coord, tree_parent, tree_sizes, max_tree_states);
    } catch (const std::exception &ex) {
        g_error = ex.what();
        return 2;
    }
    return 0;
}

// RRTx's tree growth (rpmplv2_b200/cpp/planning/Rrtx.cpp) over the oracle: the CPU twin of rvsplan_rrtx_grow
int orcplan_rrtx_grow(void *orc_handle, const rvsplan_rrtx_config *cfg, const float *q_root, const float *samples, int32_t S, int32_t max_states, float *coord,
                      int32_t *parent, float *cost, int32_t *new_index, int32_t path_start, int32_t *path, int32_t *path_len, int32_t *path_invalid,
                      rvsplan_rrtx_stats *stats) {
    if (!orc_handle || !cfg || !q_root || !coord || !parent || !cost) { g_error = "orcplan_rrtx_grow: bad argument"; return 1; }
    auto &eng = *static_cast<orc_capi::Engine *>(orc_handle);
    try {
        OracleBackend backend(eng);
        rvsplan::HostNeighborIndex index(eng.robot.n);
        if (rvsplan::rrtx_grow(backend, index, *cfg, q_root, samples, S, max_states, coord, parent, cost, new_index, path_start, path, path_len, path_invalid, stats))
        { g_error = "orcplan_rrtx_grow: the tree outgrew max_states"; return 1; }
    } catch (const std::exception &ex) {
        g_error = ex.what();
        return 2;
    }
    return 0;
}

}
