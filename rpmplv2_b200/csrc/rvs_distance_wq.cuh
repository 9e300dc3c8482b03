// rvs_distance_wq.cuh — the MIN / PROFILE distance kernel of the RealVectorSpaceCUDA engine (sm_100a):
// computeDistance (RealVectorSpace.cpp:318-392) restructured as   scan -> warp work queues -> compacted phases.
//
// Why: the one-thread-per-configuration kernel (rvs_kernels.cuh, still used for FULL mode) is issue-bound with only
// 12.8 of 32 lanes active (ncu, profiles/r1a): every lane walks the reference's 27-region case tree for every
// (link, obstacle) pair although ~97 % of the pairs cannot change the per-link minimum.
//
// How (one warp = 32 configurations, lane = configuration during the scan, lane = work item afterwards):
//   phase A  scan      each lane tests every obstacle of the shared-memory tile against its links' axis-aligned
//                      bounding boxes inflated by (running minimum + radius): 6 compares per pair, no divergence.
//                      Survivors are appended to the warp's pair queue Q1 in shared memory.
//   phase B  classify  whenever Q1 holds >= 32 pairs, every lane pops ONE pair (of any configuration of the warp),
//                      re-tests it against the now tighter minimum with the Euclidean box gap, and runs the
//                      reference's endpoint-projection / region classification (classify_pair).  Immediate results
//                      are merged; the box edges the reference would hand to distanceLineSegToLineSeg are appended
//                      to the edge queue Q2.
//   phase C  edges     whenever Q2 holds >= 32 edges, every lane pops ONE edge and evaluates seg_edge().
// All three phases run with (nearly) full warps.  Results are merged with shared-memory atomicMin on the IEEE bit
// pattern (distances that reach the minimum are > 0, so integer order == float order).
//
// Exactness (bit-identical outputs to the sequential kernel / the reference loop, in any processing order):
//   * culling is conservative: a pair is skipped only if a lower bound of ANY candidate distance the reference
//     could compute for it exceeds (current minimum of that link + radius) by a safety margin (0.1 % + 1e-4 m),
//     so a skipped pair can neither lower the minimum nor be a collision;
//   * min(c_k) - r == min(c_k - r) and the min_dist_tol filter d > tol is monotone, so candidates can be merged one
//     by one instead of per pair;
//   * first-collision semantics (RealVectorSpace.cpp:370-378) are order dependent in the reference: we keep the
//     minimum of key = link * M + obstacle over ALL colliding pairs (links below the current hit are never
//     dropped), then replay exactly that one pair sequentially to obtain the value the reference stores in
//     d_c_profile[link]; the pair count the reference would have visited follows from the same key.
#pragma once
#include <cuda_runtime.h>

#include "rvs_fk.cuh"
#include "rvs_geom.cuh"
#include "rvs_kernels.cuh"
#include "rvs_types.h"

namespace rvs {

constexpr int kWqBlock = 128;
constexpr int kWqWarps = kWqBlock / 32;
constexpr int kWqStride = kWqBlock + 1;           // row pitch of the per-thread shared arrays: row r of thread t sits in bank (r + t) % 32, so
                                                  // popped work items of ONE configuration (same t, different links) do not collide
constexpr int kWqTile = 1024;                     // obstacles per shared-memory tile (32 B each)
constexpr int kWqChunk = 32;                      // obstacles scanned between two refreshes of the inflated link boxes
constexpr unsigned kNoHit = 0xFFFFFFFFu;
constexpr unsigned kNoEntry = 0xFFFFFFFFu;

template <int G> struct WqCaps {
    static constexpr int q1 = 32 * G + 32;        // one scan iteration can push 32*G pairs on top of a leftover < 32
    static constexpr int q2 = 32 * 12 + 32;       // one classify round can push 32*12 edges on top of a leftover < 32
};

template <class R> struct WqBits;
template <> struct WqBits<float> {
    typedef int type;
    static __device__ __forceinline__ int to(float x) { return __float_as_int(x); }
    static __device__ __forceinline__ float from(int b) { return __int_as_float(b); }
};
template <> struct WqBits<double> {
    typedef long long type;
    static __device__ __forceinline__ long long to(double x) { return __double_as_longlong(x); }
    static __device__ __forceinline__ double from(long long b) { return __longlong_as_double(b); }
};

// bytes of dynamic shared memory for L links, a tile of T obstacles and scan group width G
template <class R, int G>
__host__ __device__ inline size_t wq_smem_bytes(int L, int T) {
    size_t b = size_t(T) * sizeof(ObstacleRec);
    b += size_t(3 * (L + 1)) * kWqStride * sizeof(R);                       // skeleton
    b += size_t(L) * kWqStride * sizeof(typename WqBits<R>::type);         // per-link minima
    b += size_t(L) * sizeof(R);                                           // radii
    b = (b + 15) & ~size_t(15);
    b += size_t(kWqBlock) * sizeof(unsigned);                             // first-collision keys
    b += size_t(kWqWarps) * (4 + WqCaps<G>::q1 + WqCaps<G>::q2) * sizeof(unsigned);
    return b;
}

template <class R>
struct WqShared {
    ObstacleRec *tile;
    R *skel;                                  // [3*(L+1)][kWqStride]
    typename WqBits<R>::type *best;           // [L][kWqStride]
    R *radius;                                // [L]
    unsigned *hit;                            // [kWqBlock]
    volatile unsigned *tails;                 // this warp: [0] = Q1 tail, [1] = Q2 tail
    unsigned *q1, *q2;                        // this warp's queues
};

template <class R>
struct WqCtx {
    WqShared<R> s;
    int L, M, tile_base, wbase;               // wbase = first thread of this warp in the block
    R eq_thr, max_dist;
};

template <class R> __device__ __forceinline__ R wq_inf() { return R(__int_as_float(0x7f800000)); }

template <class R> __device__ __forceinline__ R wq_threshold(R best, R radius, R eq_thr) {
    const R t = best + radius;
    return (t > eq_thr ? t : eq_thr) * R(1.001) + R(1e-4);
}

// merge one finished candidate distance (radius already subtracted) of pair `key` into configuration slot `t`
template <class R>
__device__ __forceinline__ void wq_merge(const WqCtx<R> &c, int t, int link, unsigned key, R d, R tol) {
    if (d > tol) return;                                                   // min_dist_tol: contributes MAX_DISTANCE (:367-368)
    if (d <= R(0)) atomicMin(&c.s.hit[t], key);                            // a collision of this pair (:370)
    else if (d > R(0)) atomicMin(&c.s.best[link * kWqStride + t], WqBits<R>::to(d));
}

template <class R>
__device__ __forceinline__ void wq_load_pair(const WqCtx<R> &c, int t, int link, int jt, Capsule<R> &cp, Box<R> &bx, R &tol) {
    const R *sk = c.s.skel + (3 * link) * kWqStride + t;
    const V3<R> A = {sk[0], sk[kWqStride], sk[2 * kWqStride]}, B = {sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride]};
    cp.set(A, B, c.s.radius[link], R(0));
    const float4 *t4 = reinterpret_cast<const float4 *>(c.s.tile) + 2 * jt;
    const float4 w0 = t4[0], w1 = t4[1];
    bx = {{R(w0.x), R(w0.y), R(w0.z)}, {R(w0.w), R(w1.x), R(w1.y)}};
    tol = R(w1.z);
}

// phase B: one surviving pair
template <class R, int G>
__device__ __forceinline__ void wq_classify(const WqCtx<R> &c, unsigned e) {
    const int t = c.wbase + int(e & 31u), link = int((e >> 5) & 31u), jt = int(e >> 10);
    const unsigned key = unsigned(link) * unsigned(c.M) + unsigned(c.tile_base + jt);
    if (key > c.s.hit[t]) return;                                          // an earlier pair (reference order) already collides
    Capsule<R> cp; Box<R> bx; R tol;
    wq_load_pair(c, t, link, jt, cp, bx, tol);
    // Euclidean gap between the segment's bounding box and the obstacle, against the current minimum
    const R thr = wq_threshold(WqBits<R>::from(c.s.best[link * kWqStride + t]), cp.radius, c.eq_thr);
    const R lox = cp.A.x < cp.B.x ? cp.A.x : cp.B.x, hix = cp.A.x < cp.B.x ? cp.B.x : cp.A.x;
    const R loy = cp.A.y < cp.B.y ? cp.A.y : cp.B.y, hiy = cp.A.y < cp.B.y ? cp.B.y : cp.A.y;
    const R loz = cp.A.z < cp.B.z ? cp.A.z : cp.B.z, hiz = cp.A.z < cp.B.z ? cp.B.z : cp.A.z;
    const R gx = fmax(fmax(bx.lo.x - hix, lox - bx.hi.x), R(0)), gy = fmax(fmax(bx.lo.y - hiy, loy - bx.hi.y), R(0)),
            gz = fmax(fmax(bx.lo.z - hiz, loz - bx.hi.z), R(0));
    if (gx * gx + gy * gy + gz * gz > thr * thr) return;
    V3<R> u, v;
    const PairPlan<R> pl = classify_pair<R, false>(cp, bx, c.eq_thr, c.max_dist, u, v);
    if (pl.done) { wq_merge(c, t, link, key, pl.d, tol); return; }
    wq_merge(c, t, link, key, pl.d - cp.radius, tol);                      // face distance (2a) or MAX_DISTANCE (2b)
    const unsigned pos = atomicAdd(const_cast<unsigned *>(&c.s.tails[1]), unsigned(pl.count));
#pragma unroll 1
    for (int k = 0; k < pl.count; k++) c.s.q2[pos + k] = (e & 0xFFFFFu) | (plan_edge(pl, k) << 20);
}

// phase C: one box edge of one pair
template <class R>
__device__ __forceinline__ void wq_edge(const WqCtx<R> &c, unsigned e) {
    const int t = c.wbase + int(e & 31u), link = int((e >> 5) & 31u), jt = int((e >> 10) & 1023u);
    const unsigned key = unsigned(link) * unsigned(c.M) + unsigned(c.tile_base + jt);
    if (key > c.s.hit[t]) return;
    Capsule<R> cp; Box<R> bx; R tol;
    wq_load_pair(c, t, link, jt, cp, bx, tol);
    V3<R> u, v;
    const R de = plan_edge_distance<R, false>(cp, bx, e >> 20, c.eq_thr, c.max_dist, u, v);
    wq_merge(c, t, link, key, de <= R(0) ? R(0) - cp.radius : de - cp.radius, tol);     // Q14: a touching edge gives -radius
}

// run compacted phases while a queue holds a full warp of work (or anything at all when flushing)
template <class R, int G>
__device__ __forceinline__ void wq_drain(const WqCtx<R> &c, bool flush) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        __syncwarp();
        const unsigned t2 = c.s.tails[1], t1 = c.s.tails[0];
        if (t2 >= 32u || (flush && t2 > 0u && t1 == 0u)) {
            const unsigned n = t2 < 32u ? t2 : 32u, base = t2 - n;
            const unsigned e = unsigned(lane) < n ? c.s.q2[base + lane] : kNoEntry;
            __syncwarp();
            if (lane == 0) c.s.tails[1] = base;
            __syncwarp();
            if (e != kNoEntry) wq_edge<R>(c, e);
        } else if (t1 >= 32u || (flush && t1 > 0u)) {
            const unsigned n = t1 < 32u ? t1 : 32u, base = t1 - n;
            const unsigned e = unsigned(lane) < n ? c.s.q1[base + lane] : kNoEntry;
            __syncwarp();
            if (lane == 0) c.s.tails[0] = base;
            __syncwarp();
            if (e != kNoEntry) wq_classify<R, G>(c, e);
        } else
            break;
    }
}

// F = FK scalar, R = geometry scalar, MODE = RVS_DIST_MIN | RVS_DIST_PROFILE, G = links scanned per obstacle pass.
template <class F, class R, int MODE, int G>
__global__ void __launch_bounds__(kWqBlock) distance_wq_kernel(const KArgs a, const int tile_cap) {
    extern __shared__ __align__(16) unsigned char wq_smem[];
    typedef typename WqBits<R>::type Bits;
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    WqCtx<R> c;
    {
        unsigned char *p = wq_smem;
        c.s.tile = reinterpret_cast<ObstacleRec *>(p); p += size_t(tile_cap) * sizeof(ObstacleRec);
        c.s.skel = reinterpret_cast<R *>(p); p += size_t(3 * (L + 1)) * kWqStride * sizeof(R);
        c.s.best = reinterpret_cast<Bits *>(p); p += size_t(L) * kWqStride * sizeof(Bits);
        c.s.radius = reinterpret_cast<R *>(p); p += size_t(L) * sizeof(R);
        p = wq_smem + ((size_t(p - wq_smem) + 15) & ~size_t(15));
        c.s.hit = reinterpret_cast<unsigned *>(p); p += size_t(kWqBlock) * sizeof(unsigned);
        unsigned *wq = reinterpret_cast<unsigned *>(p) + size_t(warp) * (4 + WqCaps<G>::q1 + WqCaps<G>::q2);
        c.s.tails = wq; c.s.q1 = wq + 4; c.s.q2 = wq + 4 + WqCaps<G>::q1;
    }
    c.L = L; c.M = a.M; c.wbase = warp * 32; c.tile_base = 0;
    c.eq_thr = R(a.eq_thr); c.max_dist = R(a.max_dist);

    const int64_t cfg = int64_t(blockIdx.x) * kWqBlock + threadIdx.x;
    const bool live = cfg < a.N;
    const int t = threadIdx.x;
    if (live) fk_skeleton<F, R>(rb, a.q + cfg * L, c.s.skel + t, kWqStride);
    else for (int k = 0; k < 3 * (L + 1); k++) c.s.skel[k * kWqStride + t] = R(0);
    for (int i = 0; i < L; i++) c.s.best[i * kWqStride + t] = WqBits<R>::to(c.max_dist);
    c.s.hit[t] = kNoHit;
    if (t < L) c.s.radius[t] = RobotScalars<R>::radius(rb, t);
    if (lane == 0) { c.s.tails[0] = 0; c.s.tails[1] = 0; }
    const int ground_included = rb->ground_included;

    for (int base = 0; base < a.M; base += tile_cap) {
        const int count = min(tile_cap, a.M - base);
        __syncthreads();
        stage_tile(c.s.tile, a.obstacles + base, count);
        __syncthreads();
        c.tile_base = base;
        const float4 *t4 = reinterpret_cast<const float4 *>(c.s.tile);
#pragma unroll 1
        for (int g0 = 0; g0 < L; g0 += G) {
            const int gskip = ground_included - g0;                       // links g0 .. g0+gskip-1 ignore ground objects (:336)
            const unsigned ground_keep = gskip <= 0 ? ~0u : (gskip >= G ? 0u : ~((1u << gskip) - 1u));
#pragma unroll 1
            for (int j0 = 0; j0 < count; j0 += kWqChunk) {
                // inflated link boxes from the current minima; links at or above a known collision are dead
                R lo[G][3], hi[G][3];
                const unsigned hk = c.s.hit[t];
                const int limit = !live ? 0 : (hk == kNoHit ? L : int(hk / unsigned(a.M)) + 1);
#pragma unroll
                for (int l = 0; l < G; l++) {
                    const int i = g0 + l;
                    if (i < limit && i < L) {
                        const R *sk = c.s.skel + (3 * i) * kWqStride + t;
                        const R thr = wq_threshold(WqBits<R>::from(c.s.best[i * kWqStride + t]), c.s.radius[i], c.eq_thr);
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const R p = sk[k * kWqStride], q = sk[(3 + k) * kWqStride];
                            lo[l][k] = (p < q ? p : q) - thr;
                            hi[l][k] = (p < q ? q : p) + thr;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; k++) { lo[l][k] = wq_inf<R>(); hi[l][k] = -wq_inf<R>(); }          // overlaps nothing
                    }
                }
                const int jend = min(j0 + kWqChunk, count);
#pragma unroll 1
                for (int j = j0; j < jend; j++) {
                    const float4 w0 = t4[2 * j], w1 = t4[2 * j + 1];      // (lo.xyz, hi.x) (hi.y, hi.z, tol, flags): warp-wide broadcast
                    const R ox = R(w0.x), oy = R(w0.y), oz = R(w0.z), px = R(w0.w), py = R(w1.x), pz = R(w1.y);
                    unsigned mask = 0;
#pragma unroll
                    for (int l = 0; l < G; l++) {
                        const bool ov = ox <= hi[l][0] && px >= lo[l][0] && oy <= hi[l][1] && py >= lo[l][1] && oz <= hi[l][2] && pz >= lo[l][2];
                        mask |= ov ? (1u << l) : 0u;
                    }
                    if (__float_as_uint(w1.w) & 1u) mask &= ground_keep;
                    if (mask) {
                        unsigned pos = atomicAdd(const_cast<unsigned *>(&c.s.tails[0]), unsigned(__popc(mask)));
                        do {
                            const int l = __ffs(int(mask)) - 1;
                            mask &= mask - 1u;
                            c.s.q1[pos++] = unsigned(lane) | (unsigned(g0 + l) << 5) | (unsigned(j) << 10);
                        } while (mask);
                    }
                    __syncwarp();
                    if (c.s.tails[0] >= 32u) wq_drain<R, G>(c, false);
                }
            }
        }
        wq_drain<R, G>(c, true);                                            // queue entries index this tile: empty them before it is replaced
    }
    __syncwarp();

    // ---- per-configuration epilogue (lane = configuration again) ---------------------------------------------------
    const unsigned hk = c.s.hit[t];
    const bool collided = live && hk != kNoHit;
    const int limit = collided ? int(hk / unsigned(a.M)) : L;
    const int hit_j = collided ? int(hk % unsigned(a.M)) : 0;
    R hit_d = 0;
    if (collided) {
        // the value the reference stores for the colliding link: replay exactly that pair sequentially
        Capsule<R> cp;
        const R *sk = c.s.skel + (3 * limit) * kWqStride + t;
        cp.set({sk[0], sk[kWqStride], sk[2 * kWqStride]}, {sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride]}, c.s.radius[limit], R(0));
        V3<R> u, v;
        hit_d = capsule_box_distance<R, false>(cp, load_box<R>(a.obstacles[hit_j]), c.eq_thr, c.max_dist, u, v);
    }
    if (a.pair_count) {
        // pair queries the REFERENCE's loop (link-major, stop at the first collision) would have made
        unsigned visited = 0;
        if (live) {
            for (int i = 0; i < limit; i++) visited += unsigned(a.M - (i < ground_included ? a.num_ground : 0));
            if (collided) {
                unsigned upto = unsigned(hit_j + 1);
                if (limit < ground_included && a.num_ground > 0)
                    for (int j = 0; j <= hit_j; j++) upto -= (a.obstacles[j].flags & 1u);
                visited += upto;
            }
        }
        for (int off = 16; off > 0; off >>= 1) visited += __shfl_down_sync(0xffffffffu, visited, off);
        if (lane == 0 && visited) atomicAdd(a.pair_count, (unsigned long long)visited);
    }
    if (!live) return;
    R d_c = c.max_dist;
    for (int i = 0; i < L; i++) { const R b = WqBits<R>::from(c.s.best[i * kWqStride + t]); if (b < d_c) d_c = b; }
    static_cast<R *>(a.d_c)[cfg] = collided ? R(0) : d_c;
    if (MODE >= RVS_DIST_PROFILE) {
        R *out = static_cast<R *>(a.profile) + cfg * L;
        for (int i = 0; i < L; i++) {
            const R b = WqBits<R>::from(c.s.best[i * kWqStride + t]);
            out[i] = !collided ? b : (i < limit ? b : (i == limit ? hit_d : R(0)));
        }
    }
}

}  // namespace rvs
