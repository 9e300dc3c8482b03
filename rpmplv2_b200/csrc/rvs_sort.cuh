// rvs_sort.cuh — order a large batch of configurations so that the 32 configurations of a warp are geometrically alike.
//
// The work-queue distance kernel (rvs_distance_wq.cuh) gives one warp 32 configurations; everything it decides per warp — which links
// are still open, where the sorted obstacle table can be cut off, when the queues hold a full warp of one kind of work — is as good as
// those 32 arms are similar.  A batch arrives in the caller's order (random for sweeps and dataset generation), so the kernel is given
// a permutation: a counting sort on a key made of the three leading joints (they place most of the chain; two joints or four measured
// slower), quantised so that a bin holds ~128 configurations (measured optimum: finer bins cost more in the sort than they return in
// the kernel; profiles/r2).  Three small kernels (histogram, scan of <= 2^21 bins, scatter); results are written back through the same
// permutation, so the caller sees its own order and per-configuration results do not depend on it.
#pragma once
#include <cuda_runtime.h>

#include "rvs_types.h"

namespace rvs {

struct SortArgs {
    const float *q;          // [N][n]
    int64_t N;
    int32_t n;
    int32_t joints;          // leading joints in the key (<= 6)
    int32_t bits[6];         // bits per joint, most significant joint first
    float lo[6], inv[6];     // joint range: u = (q - lo) * inv in [0, 1)
    uint32_t *hist;          // [1 << total bits] zeroed by the caller; becomes the write cursor of the scatter
    int32_t nbins;
    int32_t *perm;           // [N] out: perm[k] = index of the configuration processed k-th
};

__device__ __forceinline__ uint32_t sort_key(const SortArgs &a, int64_t i) {
    uint32_t key = 0;
    for (int j = 0; j < a.joints; j++) {
        const float u = (a.q[i * a.n + j] - a.lo[j]) * a.inv[j];
        const int cells = 1 << a.bits[j];
        int b = int(u * float(cells));
        b = b < 0 ? 0 : (b >= cells ? cells - 1 : b);                    // (NaN -> 0)
        key = (key << a.bits[j]) | uint32_t(b);
    }
    return key;
}

#ifdef RVS_VARIANT_EXACT
__global__ void __launch_bounds__(256) sort_hist_kernel(const SortArgs a) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < a.N) atomicAdd(a.hist + sort_key(a, i), 1u);
}
// exclusive scan of the histogram by ONE block of 1024 threads (<= 2^21 bins): every thread owns a contiguous run of bins and walks it
// with 16-byte loads; the 1024 partial sums are scanned by warp shuffles
__global__ void __launch_bounds__(1024) sort_scan_kernel(const SortArgs a) {
    __shared__ uint32_t warp_sum[32];
    const int per = ((a.nbins + 1023) / 1024 + 3) & ~3, b0 = threadIdx.x * per, b1 = min(b0 + per, a.nbins);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t sum = 0;
    if (a.nbins >= 4096) {
        for (int b = b0; b < b1; b += 4) { const uint4 v = *reinterpret_cast<const uint4 *>(a.hist + b); sum += v.x + v.y + v.z + v.w; }
    } else
        for (int b = b0; b < b1; b++) sum += a.hist[b];
    uint32_t incl = sum;
    for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane];
        for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += v; }
        warp_sum[lane] = w;
    }
    __syncthreads();
    uint32_t run = incl - sum + (warp > 0 ? warp_sum[warp - 1] : 0u);
    if (a.nbins >= 4096) {
        for (int b = b0; b < b1; b += 4) {
            uint4 v = *reinterpret_cast<const uint4 *>(a.hist + b);
            const uint32_t c0 = v.x, c1 = v.y, c2 = v.z, c3 = v.w;
            v.x = run; v.y = run + c0; v.z = v.y + c1; v.w = v.z + c2;
            run = v.w + c3;
            *reinterpret_cast<uint4 *>(a.hist + b) = v;
        }
    } else
        for (int b = b0; b < b1; b++) { const uint32_t c = a.hist[b]; a.hist[b] = run; run += c; }
}
__global__ void __launch_bounds__(256) sort_scatter_kernel(const SortArgs a) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < a.N) a.perm[atomicAdd(a.hist + sort_key(a, i), 1u)] = int32_t(i);
}
#endif

}  // namespace rvs
