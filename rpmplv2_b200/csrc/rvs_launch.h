// rvs_launch.h — launch tables exported by the per-variant translation units (rvs_variant.cu) to the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "rvs_types.h"

namespace rvs {

typedef cudaError_t (*LaunchFn)(const KArgs &, cudaStream_t);

#ifndef RVS_SMALL_BATCH
#define RVS_SMALL_BATCH 4096
#endif
// rvs_distance* calls up to this many configurations take the one-CTA-per-configuration kernel (rvs_store.cuh): measured cross-overs
// (profiles/README.md; profiles/r2/full_small_probe*.txt for FULL against the three-launch split), ~4k configurations for MIN / PROFILE,
// ~8k for FULL, ~1k against long obstacle tables
#ifndef RVS_SMALL_BATCH_FULL
#define RVS_SMALL_BATCH_FULL (2 * RVS_SMALL_BATCH)
#endif
inline int64_t distance_small_batch(int32_t mode, int32_t M) {
    const int64_t b = mode == RVS_DIST_FULL ? RVS_SMALL_BATCH_FULL : RVS_SMALL_BATCH;
    return M > 256 ? RVS_SMALL_BATCH / 4 : b;
}

struct LaunchTable {
    LaunchFn distance;
    LaunchFn is_valid;
    LaunchFn edge_valid;
    LaunchFn underestimation;
    LaunchFn skeleton;
    LaunchFn enclosing_radii;
    LaunchFn self_collision;
    LaunchFn self_collision_edges;
    LaunchFn primitive;
    cudaError_t (*spines)(const SpineArgs &, cudaStream_t);
    cudaError_t (*store_distance)(const StoreArgs &, cudaStream_t);
    cudaError_t (*store_spines)(const StoreArgs &, cudaStream_t);
    cudaError_t (*motion_valid)(const EnvArgs &, cudaStream_t);
};

LaunchTable launch_table_ref32();    // FK FP64, geometry FP32, no FMA contraction, IEEE div/sqrt
LaunchTable launch_table_f64();      // everything FP64, no FMA contraction
LaunchTable launch_table_fast32();   // everything FP32, FMA contraction allowed
cudaError_t launch_generate_configs(const KArgs &a, float *out, cudaStream_t st);
cudaError_t launch_env_advance(const EnvArgs &a, cudaStream_t st);          // K updates of every obstacle -> timeline
cudaError_t launch_env_commit(const EnvArgs &a, cudaStream_t st);           // adopt timeline[kc], rewrite both copies of the obstacle table
cudaError_t launch_nn_nearest(const NnArgs &a, cudaStream_t st);
cudaError_t launch_nn_radius(const NnArgs &a, cudaStream_t st);
struct SortArgs;
cudaError_t launch_sort_configs(const SortArgs &a, cudaStream_t st);   // histogram (zeroed by the caller) -> scan -> scatter: a.perm
cudaError_t launch_df_build(const ObstacleRec *obs, int M, float max_dist, void *df, int dim, const float *origin, float cell, cudaStream_t st);   // coarse distance field
cudaError_t launch_pack_scan(const ObstacleRec *scan, int M, void *pack, cudaStream_t st);   // fp16 copy of the scan-order obstacle table
cudaError_t launch_fma_peak(int fp64, void *sink, int blocks, int iters, cudaStream_t st);

}  // namespace rvs
