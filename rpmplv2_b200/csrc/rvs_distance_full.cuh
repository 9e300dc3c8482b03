// rvs_distance_full.cuh — FULL mode of computeDistance (RealVectorSpace.cpp:318-392: d_c, the per-link profile AND the nearest points of
// every (link, obstacle) pair) for large batches, as three launches:
//   1. distance_wq_kernel<PROFILE> (rvs_distance_wq.cuh): d_c, the profile with the reference's early-exit fill, the pair accounting;
//   2. full_list_kernel: the configurations with d_c > 0, in the caller's order (the coherent-warp permutation of step 1 is NOT used here:
//      measured, profiles/r2 — neighbouring threads then write result rows megabytes apart and step 3 takes 1.5x as long);
//   3. nearest_points_kernel: one thread per LISTED configuration walks all L x M pairs through the reference's case analysis and
//      writes the 24-byte nearest-point pieces.
// Why: the nearest-point block of a configuration in collision is unspecified (include/rvs_cuda.h; the reference leaves the loop at
// the first hit, :370-378), and in the one-thread-per-configuration kernel (distance_kernel<FULL>, rvs_kernels.cuh) those
// configurations are lanes that sit idle until the free ones of their warp have walked every pair (ncu, profiles/r2: 9.7 of 32 lanes).
// With the list every lane of step 3 is a free configuration and does the same L x M trips; nothing is culled, nothing exits early.
// Step 3 also differs from distance_kernel<FULL> in its memory behaviour: the skeleton sits in a shared-memory column instead of a
// dynamically indexed local array, the loop is obstacle-major / link-minor so that the L pieces of one obstacle (24 L contiguous bytes
// of the [M][L][6] block) are written back to back and meet in L2 before they are evicted, a piece goes out as one 16-byte and one
// 8-byte store, and the edge evaluations are deferred onto per-thread rings (below).  Same per-pair routines in the same order per
// pair (classify_pair, then plan_edge_distance over the planned edges = capsule_box_distance<R, true>), so the points are those of
// distance_kernel<FULL> bit for bit (tests/test_gpu_parity.py compares the two on every precision).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "rvs_kernels.cuh"

namespace rvs {

constexpr int kFullTile = 256;          // obstacles staged per shared-memory tile (8 KB)
constexpr int kListBlock = 256;

// list[0] = count (zeroed by the launcher), entries from list + 4
template <class R>
__global__ void __launch_bounds__(kListBlock) full_list_kernel(const KArgs a) {
    __shared__ unsigned warp_base[kListBlock / 32 + 1];
    const int64_t slot = int64_t(blockIdx.x) * kListBlock + threadIdx.x;
    const bool live = slot < a.N;
    const int32_t cfg = int32_t(slot);
    const bool keep = live && static_cast<const R *>(a.d_c)[cfg] > R(0);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_base[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned total = 0;
        for (int w = 0; w < kListBlock / 32; w++) { const unsigned c = warp_base[w]; warp_base[w] = total; total += c; }
        warp_base[kListBlock / 32] = total ? atomicAdd(reinterpret_cast<unsigned *>(a.list), total) : 0u;
    }
    __syncthreads();
    if (keep) a.list[4 + warp_base[kListBlock / 32] + warp_base[warp] + __popc(mask & ((1u << lane) - 1u))] = cfg;
}

// ranges of the obstacle table per batch of kBlock configurations: enough blocks for ~7 waves of 4 CTAs on 148 SMs if every configuration
// were listed (2220 / 4440 measured, profiles/r2), at least 8 obstacles per range
#ifndef RVS_FULL_PARTS_TARGET
#define RVS_FULL_PARTS_TARGET 4440
#endif
inline int full_parts(int64_t N, int M) {
    const char *e = std::getenv("RVS_FULL_PARTS_TARGET");              // (read per launch: tests use 1 to walk a long table in one range, tile after tile)
    const int64_t target = e ? std::max(1, std::atoi(e)) : RVS_FULL_PARTS_TARGET;
    const int64_t blocks = (N + kBlock - 1) / kBlock;
    const int64_t want = (target + blocks - 1) / blocks;
    return int(std::max<int64_t>(1, std::min<int64_t>(want, M / 8)));
}

template <class R> struct Pair2;
template <> struct Pair2<float> { typedef float2 type; static __device__ __forceinline__ float2 make(float x, float y) { return make_float2(x, y); } };
template <> struct Pair2<double> { typedef double2 type; static __device__ __forceinline__ double2 make(double x, double y) { return make_double2(x, y); } };

// (101 registers -> 4 CTAs per SM; 5 / 6 CTAs measured slower: the store path is the limit, profiles/r2/ncu_details_nearest_points_1m.txt)
#ifndef RVS_FULL_MINBLOCKS
#define RVS_FULL_MINBLOCKS 4
#endif
// The case analysis of a pair (classify_pair, rvs_geom.cuh) ends most pairs on the spot; the others leave 1, 2, 4 or 12 box edges for
// distanceLineSegToLineSeg.  Run pair by pair, that edge loop is where the warp falls apart (ncu, profiles/r2: seg_edge and its helpers
// are two thirds of the instructions at 5 of 32 lanes, classify_pair runs at 29).  So the edge work is DEFERRED: a thread pushes the
// pairs that need edges onto its own ring in shared memory (kFullRing entries — 4 / 8 / 16 / 32 measured, profiles/r2 — each the plan,
// the pair and the face candidate) and goes on
// classifying; when some lane's ring is full, every lane with work evaluates ONE edge per trip — one converged seg_edge per lane
// whatever pair or edge count it belongs to — until all rings are at most half full.  A pair's edges are still evaluated in the
// reference's order by the one thread that owns the pair, so the nearest points are those of the sequential loop bit for bit.
#ifndef RVS_FULL_RING
#define RVS_FULL_RING 8
#endif
constexpr int kFullRing = RVS_FULL_RING;       // a power of two

inline size_t full_smem_bytes(int n, size_t scalar) {
    return size_t(kFullTile) * sizeof(ObstacleRec) + size_t(3 * (n + 1)) * kBlock * scalar + size_t(kFullRing) * kBlock * (2 * sizeof(uint32_t) + scalar);
}

// One 24-byte piece (6 floats) = one 16-byte and one 8-byte store, whichever order keeps the 16-byte one aligned: two requests per piece
// instead of three (every lane of a request writes into a different sector — the rows of a warp's configurations are kilobytes apart —
// so the sector count of this kernel's stores is requests x 32).  FP64 pieces: three 16-byte stores.
template <class R> __device__ __forceinline__ void store_points(R *dst, bool vec, const V3<R> &P1, const V3<R> &P2) {
    if (vec) {
        typename Pair2<R>::type *d2 = reinterpret_cast<typename Pair2<R>::type *>(dst);
        d2[0] = Pair2<R>::make(P1.x, P1.y); d2[1] = Pair2<R>::make(P1.z, P2.x); d2[2] = Pair2<R>::make(P2.y, P2.z);
    } else {
        dst[0] = P1.x; dst[1] = P1.y; dst[2] = P1.z; dst[3] = P2.x; dst[4] = P2.y; dst[5] = P2.z;
    }
}
#ifndef RVS_FULL_STORE16
#define RVS_FULL_STORE16 1
#endif

template <> __device__ __forceinline__ void store_points<float>(float *dst, bool vec, const V3<float> &P1, const V3<float> &P2) {
    if (vec && RVS_FULL_STORE16) {
        if ((reinterpret_cast<size_t>(dst) & 15) == 0) {
            *reinterpret_cast<float4 *>(dst) = make_float4(P1.x, P1.y, P1.z, P2.x);
            *reinterpret_cast<float2 *>(dst + 4) = make_float2(P2.y, P2.z);
        } else {
            *reinterpret_cast<float2 *>(dst) = make_float2(P1.x, P1.y);
            *reinterpret_cast<float4 *>(dst + 2) = make_float4(P1.z, P2.x, P2.y, P2.z);
        }
    } else if (vec) {
        float2 *d2 = reinterpret_cast<float2 *>(dst);
        d2[0] = make_float2(P1.x, P1.y); d2[1] = make_float2(P1.z, P2.x); d2[2] = make_float2(P2.y, P2.z);
    } else {
        dst[0] = P1.x; dst[1] = P1.y; dst[2] = P1.z; dst[3] = P2.x; dst[4] = P2.y; dst[5] = P2.z;
    }
}

template <class F, class R>
__global__ void __launch_bounds__(kBlock, sizeof(R) == 4 ? RVS_FULL_MINBLOCKS : 1) nearest_points_kernel(const KArgs a) {
    extern __shared__ float4 full_smem4[];
    ObstacleRec *tile = reinterpret_cast<ObstacleRec *>(full_smem4);
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n;
    R *sk = reinterpret_cast<R *>(tile + kFullTile) + threadIdx.x;        // component c of point k at sk[(3k + c) * kBlock]
    R *ring_d = sk - threadIdx.x + size_t(3 * (L + 1)) * kBlock + threadIdx.x;                                   // entry e: candidate distance at ring_d[e * kBlock]
    uint32_t *ring_plan = reinterpret_cast<uint32_t *>(ring_d - threadIdx.x + size_t(kFullRing) * kBlock) + threadIdx.x;   // edge codes (24) | count (4) | candidate (4)
    uint32_t *ring_pair = ring_plan + size_t(kFullRing) * kBlock;                                                // link | obstacle-in-tile << 8
    const unsigned listed = *reinterpret_cast<const unsigned *>(a.list);
    if (unsigned(blockIdx.x) * kBlock >= listed) return;                 // (block-uniform)
    const unsigned k = unsigned(blockIdx.x) * kBlock + threadIdx.x;
    const bool live = k < listed;
    const int64_t cfg = live ? int64_t(a.list[4 + k]) : 0;
    const R eq_thr = R(a.eq_thr), max_dist = R(a.max_dist);
    if (live) fk_skeleton<F, R>(rb, a.q + cfg * L, sk, kBlock);
    R *nearest = static_cast<R *>(a.nearest) + size_t(cfg) * size_t(a.M) * size_t(L) * 6;
    const bool vec = (reinterpret_cast<size_t>(a.nearest) & (2 * sizeof(R) - 1)) == 0;      // a piece is 6 R: pairs stay aligned when the block is
    const int ground_included = rb->ground_included;

    auto capsule_of = [&](int i) {
        Capsule<R> cp;
        cp.set({sk[(3 * i) * kBlock], sk[(3 * i + 1) * kBlock], sk[(3 * i + 2) * kBlock]},
               {sk[(3 * i + 3) * kBlock], sk[(3 * i + 4) * kBlock], sk[(3 * i + 5) * kBlock]}, RobotScalars<R>::radius(rb, i), R(0));
        return cp;
    };

    // gridDim.y blocks share a batch of configurations, each walking its own range of the obstacle table: batches that would not fill the
    // GPU with one thread per configuration (full_parts below) get that many times the warps
    const int m_lo = int(int64_t(a.M) * blockIdx.y / gridDim.y), m_hi = int(int64_t(a.M) * (blockIdx.y + 1) / gridDim.y);
    for (int base = m_lo; base < m_hi; base += kFullTile) {
        const int count = min(kFullTile, m_hi - base);
        __syncthreads();
        stage_tile(tile, a.obstacles + base, count);
        __syncthreads();
        R *tile_dst = nearest + size_t(base) * size_t(L) * 6;
        int head = 0, queued = 0;          // my ring: entries [head, head + queued)
        int edge_k = 0;                    // edges of the front entry already evaluated
        R best = 0;                        // ... and the best candidate so far
        V3<R> B1 = {0, 0, 0}, B2 = {0, 0, 0};
        // one edge of the front entry of every lane that has one
        auto edge_trip = [&]() {
            if (queued > 0) {
                const uint32_t plan = ring_plan[head * kBlock], pair = ring_pair[head * kBlock];
                const int i = int(pair & 255u), j = int(pair >> 8);
                const Capsule<R> cp = capsule_of(i);
                const Box<R> bx = load_box<R>(tile[j]);
                const int edges = int((plan >> 24) & 15u);
                if (edge_k == 0) {                                      // the candidate classify_pair left (face projection, or none)
                    best = ring_d[head * kBlock];
                    plan_candidate_points(cp, bx, plan >> 28, B1, B2);
                }
                PairPlan<R> pl;
                pl.list = plan & 0xffffffu; pl.count = edges; pl.twelve = edges == 12;
                V3<R> q1, q2;
                const R de = plan_edge_distance<R, true>(cp, bx, plan_edge(pl, edge_k), eq_thr, max_dist, q1, q2);
                if (de > R(0) && de < best) { best = de; B1 = q1; B2 = q2; }
                if (++edge_k == edges) {
                    store_points(tile_dst + (size_t(j) * L + i) * 6, vec, B1, B2);
                    head = (head + 1) & (kFullRing - 1); queued--; edge_k = 0;
                }
            }
        };
#pragma unroll 1
        for (int j = 0; j < count; j++) {
            const ObstacleRec &o = tile[j];
            const bool ground = (o.flags & 1u) != 0u;
            const Box<R> bx = load_box<R>(o);
            R *dst = tile_dst + size_t(j) * size_t(L) * 6;
#pragma unroll 1
            for (int i = 0; i < L; i++, dst += 6) {
                if (live) {
                    V3<R> P1 = {0, 0, 0}, P2 = {0, 0, 0};
                    bool done = true;
                    if (ground && i < ground_included) {                   // ground rule (:336-341)
                        P2.z = -max_dist;
                    } else {
                        const Capsule<R> cp = capsule_of(i);
                        const PairPlan<R> pl = classify_pair<R, true>(cp, bx, eq_thr, max_dist, P1, P2);
                        if (!pl.done) {
                            done = false;
                            const int e = (head + queued) & (kFullRing - 1);
                            ring_plan[e * kBlock] = (pl.twelve ? 0u : pl.list) | (uint32_t(pl.count) << 24) | (pl.cand << 28);
                            ring_pair[e * kBlock] = uint32_t(i) | (uint32_t(j) << 8);
                            ring_d[e * kBlock] = pl.d;
                            queued++;
                        }
                    }
                    if (done) store_points(dst, vec, P1, P2);
                }
                if (__any_sync(0xffffffffu, queued == kFullRing)) {
                    do edge_trip(); while (__any_sync(0xffffffffu, queued > kFullRing / 2));
                }
            }
        }
        while (__any_sync(0xffffffffu, queued > 0)) edge_trip();          // the tile is restaged next: nothing may stay queued
    }
}

}  // namespace rvs
