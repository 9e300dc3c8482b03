// rvs_oracle_engine.hpp — the oracle's handle type, shared by its ctypes entry points (rvs_oracle_capi.cpp) and by the
// oracle-backed planner driver (oracle_planning.cpp). TEST INFRASTRUCTURE ONLY (see rvs_oracle.hpp).
#pragma once
#include <atomic>
#include <thread>

#include "../include/rvs_cuda.h"
#include "rvs_oracle.hpp"

namespace orc_capi {

struct Engine {
    orc::Robot robot;
    std::vector<float> lo, hi;
    rvs_params prm;
    orc::Scene<float> sc32;
    orc::Scene<double> sc64;
    std::atomic<uint64_t> pair_count{0};
};

template <class T> struct Pick;
template <> struct Pick<float> { static const orc::Scene<float> &scene(const Engine &e) { return e.sc32; } };
template <> struct Pick<double> { static const orc::Scene<double> &scene(const Engine &e) { return e.sc64; } };

template <class T>
orc::Config<T> config_of(const rvs_params &p) {
    return {T(p.equality_threshold), T(p.max_distance), int(p.num_interpolation_validity_checks)};
}

// static split of [0, N) over `threads` std::threads (the reference itself is single-threaded)
template <class F>
void parallel_for(int64_t N, int threads, F &&body) {
    if (threads <= 1 || N < 2) { body(int64_t(0), N); return; }
    std::vector<std::thread> pool;
    const int64_t chunk = (N + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const int64_t b = t * chunk, e = std::min<int64_t>(N, b + chunk);
        if (b >= e) break;
        pool.emplace_back([=, &body] { body(b, e); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace orc_capi
