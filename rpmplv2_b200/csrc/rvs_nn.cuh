// rvs_nn.cuh — nearest-neighbour and radius queries over device-resident planner trees.
//
// Reference: base::Tree::getNearestState (src/state_spaces/Tree.cpp:35-50; a nanoflann kd-tree answering EXACT 1-NN queries in the
// squared L2 metric, L2_Simple_Adaptor: a sequential sum of squared differences) and planning::rrtx::RRTx::findNeighbors
// (src/planners/rrtx/RRTx.cpp:424-452: every tree state within `radius` of q in ss->getNorm, tree order, at most MAX_NEIGHBORS nearest).
// Here: brute force.  A tree is a row-major [count][n] float array in HBM (24 B per xArm6 state: 10^5 states = 2.4 MB, L2 resident);
// one CTA answers one query, every thread strides over the states.  Exact, so the answers are those of the kd-tree whenever the
// minimum is unique; ties go to the lowest index (what an exhaustive scan in tree order gives).
#pragma once
#include <cuda_runtime.h>

#include "rvs_kernels.cuh"
#include "rvs_types.h"

namespace rvs {

constexpr int kNnBlock = 256;

// squared L2 distance, sequential sum in index order (nanoflann L2_Simple_Adaptor::evalMetric)
__device__ __forceinline__ float nn_dist2(const float *__restrict__ a, const float *__restrict__ b, int n) {
    float s = 0.f;
    for (int d = 0; d < n; d++) { const float diff = a[d] - b[d]; s += diff * diff; }
    return s;
}

#ifdef RVS_VARIANT_EXACT            // one copy: these kernels are not templates (FP32 without FMA contraction, like the host planners' own sums)
__global__ void __launch_bounds__(kNnBlock) nn_nearest_kernel(const NnArgs a) {
    __shared__ float s_best[kNnBlock / 32];
    __shared__ int s_idx[kNnBlock / 32];
    __shared__ float s_q[kMaxDofs];
    const int k = blockIdx.x;
    const int tree = a.tree_of ? a.tree_of[k] : 0;
    const float *__restrict__ pts = a.points[tree];
    const int count = a.counts[tree];
    if (threadIdx.x < a.n) s_q[threadIdx.x] = a.queries[size_t(k) * a.n + threadIdx.x];
    __syncthreads();
    float best = __int_as_float(0x7f800000);
    int best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < count; i += kNnBlock) {
        const float s = nn_dist2(s_q, pts + size_t(i) * a.n, a.n);
        if (s < best) { best = s; best_i = i; }                      // strided ascending indices: the first minimum of this thread
    }
    // (distance, index) minimum over the block: smaller distance wins, then the smaller index
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kNnBlock / 32; w++)
            if (s_best[w] < best || (s_best[w] == best && s_idx[w] < best_i)) { best = s_best[w]; best_i = s_idx[w]; }
        a.idx[k] = count > 0 ? best_i : -1;
        a.dist2[k] = best;
    }
}

// One CTA walks the tree in chunks of kNnBlock states and appends the hits in tree order (ballot + prefix of the warp counts), so the
// caller sees RRTx::findNeighbors' own neighbour order (chooseParent keeps the FIRST of equal-cost candidates, rewireNeighbors applies
// its changes in that order).  distance = ss->getNorm (Eigen's packet-order norm, coord_norm).
__global__ void __launch_bounds__(kNnBlock) nn_radius_kernel(const NnArgs a) {
    __shared__ int s_warp[kNnBlock / 32];
    __shared__ int s_base;
    __shared__ float s_q[kMaxDofs];
    const float *__restrict__ pts = a.points[a.tree];
    const int count = a.counts[a.tree];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < a.n) s_q[threadIdx.x] = a.queries[threadIdx.x];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < count; c0 += kNnBlock) {
        const int i = c0 + threadIdx.x;
        float d = 0.f;
        bool hit = false;
        if (i < count && i != a.exclude) {
            d = coord_norm(s_q, pts + size_t(i) * a.n, a.n);
            hit = d <= a.radius;
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        const int pos = before + __popc(b & ((1u << lane) - 1u));
        if (hit && pos < a.capacity) { a.out_idx[pos] = i; a.out_dist[pos] = d; }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < kNnBlock / 32; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.out_count = s_base;
}
#endif

}  // namespace rvs
