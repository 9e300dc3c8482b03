// rvs_variant.cu — instantiates the kernels of rvs_kernels.cuh for one arithmetic variant.
// Compiled twice (see Makefile):
//   -DRVS_VARIANT_EXACT  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false   -> REF32 and F64 launch tables
//   -DRVS_VARIANT_FAST   -fmad=true                                             -> FAST32 launch table
#include "rvs_kernels.cuh"
#include "rvs_distance_wq.cuh"
#include "rvs_distance_full.cuh"
#include "rvs_valid_wq.cuh"
#include "rvs_store.cuh"
#include "rvs_env.cuh"
#include "rvs_nn.cuh"
#include "rvs_sort.cuh"
#include "rvs_launch.h"

#include <map>
#include <mutex>
#include <utility>

namespace rvs {
namespace {

inline unsigned grid_for(int64_t n, int block) { return unsigned((n + block - 1) / block); }
constexpr int64_t kSmallValid = 2048;       // validity items (configurations, or edges x checks) up to which one CTA per item is used

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device).  One table per process: (function, device) -> bytes already
// granted; looked up under a mutex because handles on different devices may be driven from different host threads.
inline cudaError_t raise_smem_limit(const void *fn, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = granted[{fn, dev}];
    if (smem <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e == cudaSuccess) have = smem;
    return e;
}

template <class F, class R>
struct Variant {
    template <int MODE, int G>
    static cudaError_t distance_wq(const KArgs &a, cudaStream_t st) {
        const size_t smem = wq_smem_bytes<R, G>(a.robot_n_hint);
        // The dynamic shared-memory limit is an attribute of the function ON THE CURRENT DEVICE: handles may live on several devices
        // (and host threads) of one process, so it is raised per device (a small table indexed by cudaGetDevice, written under a
        // mutex) and not once per process.
        if (smem > 48u * 1024u) {
            const cudaError_t e = raise_smem_limit(reinterpret_cast<const void *>(distance_wq_kernel<F, R, MODE, G>), smem);
            if (e != cudaSuccess) return e;
        }
        distance_wq_kernel<F, R, MODE, G><<<grid_for(a.N, kWqBlock), kWqBlock, smem, st>>>(a);
        return cudaGetLastError();
    }
    // small batches (what a planner sends): one CTA per configuration (rvs_store.cuh), latency instead of throughput
    template <int MODE>
    static cudaError_t distance_cta(const KArgs &a, cudaStream_t st) {
        distance_cta_kernel<F, MODE><<<unsigned(a.N), 32 * kStoreWarps, 0, st>>>(a);
        return cudaGetLastError();
    }
    // FULL mode of a large batch (rvs_distance_full.cuh): PROFILE through the work-queue kernel, then the nearest points of the
    // collision-free configurations only
    static cudaError_t distance_full_split(const KArgs &a, cudaStream_t st) {
        KArgs p = a;
        p.mode = RVS_DIST_PROFILE;
        cudaError_t e = a.robot_n_hint <= 2 ? distance_wq<RVS_DIST_PROFILE, 2>(p, st) : distance_wq<RVS_DIST_PROFILE, kWqG>(p, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.list, 0, 4 * sizeof(int32_t), st);
        if (e != cudaSuccess) return e;
        full_list_kernel<R><<<grid_for(a.N, kListBlock), kListBlock, 0, st>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t smem = full_smem_bytes(a.robot_n_hint, sizeof(R));
        if (smem > 48u * 1024u) {
            e = raise_smem_limit(reinterpret_cast<const void *>(nearest_points_kernel<F, R>), smem);
            if (e != cudaSuccess) return e;
        }
        nearest_points_kernel<F, R><<<dim3(grid_for(a.N, kBlock), unsigned(full_parts(a.N, a.M))), kBlock, smem, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t distance(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        if (sizeof(R) == 4 && a.cull && a.M > 0 && a.N <= distance_small_batch(a.mode, a.M)) {
            if (a.mode == RVS_DIST_MIN) return distance_cta<RVS_DIST_MIN>(a, st);
            if (a.mode == RVS_DIST_PROFILE) return distance_cta<RVS_DIST_PROFILE>(a, st);
            return distance_cta<RVS_DIST_FULL>(a, st);
        }
        if (a.cull && a.mode == RVS_DIST_FULL && a.M > 0 && a.list) return distance_full_split(a, st);
        if (a.cull && a.mode != RVS_DIST_FULL && a.M > 0) {
            // work-queue kernel (rvs_distance_wq.cuh): results identical to the sequential kernel below
            const bool narrow = a.robot_n_hint <= 2;
            if (a.mode == RVS_DIST_MIN) return narrow ? distance_wq<RVS_DIST_MIN, 2>(a, st) : distance_wq<RVS_DIST_MIN, kWqG>(a, st);
            return narrow ? distance_wq<RVS_DIST_PROFILE, 2>(a, st) : distance_wq<RVS_DIST_PROFILE, kWqG>(a, st);
        }
        const unsigned g = grid_for(a.N, kBlock);
        if (a.mode == RVS_DIST_MIN) distance_kernel<F, R, RVS_DIST_MIN><<<g, kBlock, 0, st>>>(a);
        else if (a.mode == RVS_DIST_PROFILE) distance_kernel<F, R, RVS_DIST_PROFILE><<<g, kBlock, 0, st>>>(a);
        else distance_kernel<F, R, RVS_DIST_FULL><<<g, kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    template <int G, bool EDGE>
    static cudaError_t valid_wq(const KArgs &a, int64_t items, cudaStream_t st) {
        const size_t smem = vq_smem_bytes<R, G>(a.robot_n_hint);
        if (smem > 48u * 1024u) {
            const cudaError_t e = raise_smem_limit(reinterpret_cast<const void *>(valid_wq_kernel<F, R, G, EDGE>), smem);
            if (e != cudaSuccess) return e;
        }
        valid_wq_kernel<F, R, G, EDGE><<<dim3(grid_for(items, kWqBlock), unsigned(a.slices > 1 ? a.slices : 1)), kWqBlock, smem, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t is_valid(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        if (sizeof(R) == 4 && a.cull && a.M > 0 && a.N <= kSmallValid) {              // planner-sized batch: one CTA per configuration
            valid_cta_kernel<F, false><<<unsigned(a.N), 32 * kStoreWarps, 0, st>>>(a);
            return cudaGetLastError();
        }
        if (a.cull && a.M > 0 && a.M < (1 << 22)) return a.robot_n_hint <= 2 ? valid_wq<2, false>(a, a.N, st) : valid_wq<6, false>(a, a.N, st);
        is_valid_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t edge_valid(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        if (sizeof(R) == 4 && a.cull && a.M > 0 && a.N * a.num_checks <= kSmallValid) {
            valid_cta_kernel<F, true><<<unsigned(a.N * a.num_checks), 32 * kStoreWarps, 0, st>>>(a);
            return cudaGetLastError();
        }
        if (a.cull && a.M > 0 && a.M < (1 << 22))
            return a.robot_n_hint <= 2 ? valid_wq<2, true>(a, a.N * a.num_checks, st) : valid_wq<6, true>(a, a.N * a.num_checks, st);
        edge_valid_kernel<F, R><<<grid_for(a.N * a.num_checks, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t underestimation(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        underestimation_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t skeleton(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        skeleton_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t enclosing_radii(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        enclosing_radii_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t self_collision(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        self_collision_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t self_collision_edges(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        // planner-sized batches of xArm6 edges: one warp per edge (latency); sweeps: one thread per edge (throughput)
        if (a.robot_n_hint == 6 && a.N <= 8192) self_collision_edge_warp_kernel<F, R><<<grid_for(a.N, kScWarps), 32 * kScWarps, 0, st>>>(a);
        else self_collision_edge_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t primitive(const KArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        primitive_kernel<F, R><<<grid_for(a.N, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t spines(const SpineArgs &a, cudaStream_t st) {
        if (a.S <= 0) return cudaSuccess;
        spine_kernel<F><<<grid_for(a.S, 64), 64, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t store_distance(const StoreArgs &a, cudaStream_t st) {
        if (a.N <= 0) return cudaSuccess;
        store_distance_kernel<F><<<unsigned(a.N), 32 * kStoreWarps, 0, st>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t store_spines(const StoreArgs &a, cudaStream_t st) {
        if (a.S <= 0) return cudaSuccess;
        store_spine_kernel<F><<<grid_for(a.S, kSpineWarps), 32 * kSpineWarps, store_spine_smem_bytes(a.n), st>>>(a);      // <= 21 KB at n = 32
        return cudaGetLastError();
    }
    static cudaError_t motion_valid(const EnvArgs &a, cudaStream_t st) {
        if (a.E <= 0 || a.K <= 0) return cudaSuccess;
        motion_valid_kernel<F, R><<<grid_for(a.E * a.K, kBlock), kBlock, 0, st>>>(a);
        return cudaGetLastError();
    }
    static LaunchTable table() {
        return {distance, is_valid, edge_valid, underestimation, skeleton, enclosing_radii, self_collision, self_collision_edges, primitive, spines, store_distance, store_spines,
                motion_valid};
    }
};

}  // namespace

#ifdef RVS_VARIANT_EXACT
LaunchTable launch_table_ref32() { return Variant<double, float>::table(); }
LaunchTable launch_table_f64() { return Variant<double, double>::table(); }
cudaError_t launch_generate_configs(const KArgs &a, float *out, cudaStream_t st) {
    if (a.N <= 0) return cudaSuccess;
    generate_configs_kernel<<<grid_for(a.N * a.robot_n_hint, 256), 256, 0, st>>>(a, out);
    return cudaGetLastError();
}
cudaError_t launch_env_advance(const EnvArgs &a, cudaStream_t st) {
    if (a.M <= 0 || a.K <= 0) return cudaSuccess;
    if (a.motion_type == RVS_MOTION_RANDOM_DIRECTIONS) env_advance_random_kernel<<<1, 32, 0, st>>>(a);
    else env_advance_kernel<<<grid_for(a.M, 128), 128, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_env_commit(const EnvArgs &a, cudaStream_t st) {
    if (a.M <= 0 || a.K <= 0) return cudaSuccess;
    env_commit_kernel<<<grid_for(a.M, 128), 128, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    env_rank_kernel<<<grid_for(a.M, 128), 128, 0, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_pack_scan(a.obstacles_scan, a.M, a.obstacles_pack, st);
}
cudaError_t launch_nn_nearest(const NnArgs &a, cudaStream_t st) {
    if (a.Q <= 0) return cudaSuccess;
    nn_nearest_kernel<<<unsigned(a.Q), kNnBlock, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_nn_radius(const NnArgs &a, cudaStream_t st) {
    nn_radius_kernel<<<1, kNnBlock, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_sort_configs(const SortArgs &a, cudaStream_t st) {
    if (a.N <= 0) return cudaSuccess;
    sort_hist_kernel<<<grid_for(a.N, 256), 256, 0, st>>>(a);
    sort_scan_kernel<<<1, 1024, 0, st>>>(a);
    sort_scatter_kernel<<<grid_for(a.N, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_df_build(const ObstacleRec *obs, int M, float max_dist, void *df, int dim, const float *origin, float cell, cudaStream_t st) {
    df_build_kernel<<<grid_for(int64_t(dim) * dim * dim, 128), 128, 0, st>>>(obs, M, max_dist, static_cast<__half *>(df), dim, origin[0], origin[1], origin[2], cell);
    return cudaGetLastError();
}
cudaError_t launch_pack_scan(const ObstacleRec *scan, int M, void *pack, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    pack_scan_kernel<<<grid_for((M + 1) / 2, 128), 128, 0, st>>>(scan, M, static_cast<uint4 *>(pack));
    return cudaGetLastError();
}
cudaError_t launch_fma_peak(int fp64, void *sink, int blocks, int iters, cudaStream_t st) {
    if (fp64) fma_peak_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double *>(sink), iters, 1.0);
    else fma_peak_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float *>(sink), iters, 1.0f);
    return cudaGetLastError();
}
#endif
#ifdef RVS_VARIANT_FAST
LaunchTable launch_table_fast32() { return Variant<float, float>::table(); }
#endif

}  // namespace rvs
