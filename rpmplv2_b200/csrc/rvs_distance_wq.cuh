// rvs_distance_wq.cuh — the MIN / PROFILE distance kernel of the RealVectorSpaceCUDA engine (sm_100a):
// computeDistance (RealVectorSpace.cpp:318-392) restructured as   scan -> warp work queues -> compacted phases.
//
// Why: the one-thread-per-configuration kernel (rvs_kernels.cuh, still used for FULL mode) is issue-bound with only
// 12.8 of 32 lanes active (ncu, profiles/r1a): every lane walks the reference's 27-region case tree for every
// (link, obstacle) pair although ~97 % of the pairs cannot change the per-link minimum.
//
// How (one warp = 32 configurations, lane = configuration during the scan, lane = work item afterwards):
//   phase A  scan      each lane tests every obstacle (warp-uniform 16-byte loads of the HBM-resident table, served by
//                      L1) against its links' axis-aligned bounding boxes inflated by (running minimum + radius):
//                      6 predicate-chained compares per pair, no divergence.  A lane with any hit stores ONE hit word
//                      (slot mask | lane | obstacle pair) at a ballot-computed position of Q0.
//                      The table is sorted by distance d from the robot base and every link knows its reach (farthest
//                      axis point from the base): once d - reach exceeds a link's threshold for all 32 configurations,
//                      the link is dropped for the rest of the table.  Live links are packed into warp-uniform slots,
//                      so an iteration only pays for the links that are still open (2-3 of 6 on the sweep workload).
//                      The link of the current first collision only searches for EARLIER collisions (threshold =
//                      radius): its minimum is not an output, and without this a colliding configuration would push
//                      every obstacle of that link through the queues.
//   phase 0  bootstrap the running minimum of every link starts at an UPPER BOUND read from a coarse distance field of the obstacle table
//                      (df_build_kernel: 128^3 cells around the robot, fp16, distance from the cell centre to the nearest obstacle,
//                      rounded up): min over three points of the link axis of (field + half a cell diagonal) - radius.  Without it the
//                      first thresholds are infinite and the first obstacles of the table go through the narrow phase for every link.
//   phase A' unpack    Q0 >= 32: every lane turns one hit word into pair entries of Q1 (warp prefix sum).
//   phase B1 re-cull   Q1 >= 32: every lane pops ONE pair (of any configuration of the warp) and re-tests it with the
//                      Euclidean box gap against the now tighter minimum; survivors are ballot-compacted into Q1b.
//   phase B2 classify  Q1b >= 32: every lane runs the reference's endpoint-projection / region classification
//                      (classify_pair) for ONE pair.  Immediate results are merged; the box edges the reference would
//                      hand to distanceLineSegToLineSeg are appended to the edge queue Q2.
//   phase C  edges     Q2 >= 32: every lane pops ONE edge and evaluates seg_edge().
// All phases run with (nearly) full warps.  Results are merged with shared-memory atomicMin on the IEEE bit pattern
// (distances that reach the minimum are > 0, so integer order == float order).  Thresholds are refreshed after 4, 8, 16, ...,
// 128 and then every 128 obstacles; every refresh interval ends with a full flush of the queues, so the next thresholds
// (and the cut-off) are exact for the scanned prefix.
//
// Exactness (bit-identical outputs to the sequential kernel / the reference loop, in any processing order):
//   * culling is conservative: a pair is skipped only if a lower bound of ANY candidate distance the reference
//     could compute for it exceeds (current minimum of that link + radius) by a safety margin (0.1 % + 1e-4 m),
//     so a skipped pair can neither lower the minimum nor be a collision; the base-distance cut-off uses the triangle
//     inequality dist(x, box) >= dist(base, box) - |x - base| with dist(base, box) rounded DOWN on the host;
//   * min(c_k) - r == min(c_k - r) and the min_dist_tol filter d > tol is monotone, so candidates can be merged one
//     by one instead of per pair;
//   * the bootstrap bound is only a THRESHOLD: it is >= the exact distance to the obstacle that defines it (a point of the axis against
//     that box), so that obstacle survives every cull and its candidate replaces the bound.  Should the reference's routine return MORE
//     than the geometric distance for it (its case tree has slips) and nothing else come below the bound, the minimum still
//     equals the bound at the end: such a link is re-scanned without bounds (second pass of the scan loop), so outputs never depend on
//     the field.  Obstacles with a min_dist_tol or the ground label are left out of the field (they do not count for every link);
//   * first-collision semantics (RealVectorSpace.cpp:370-378) are order dependent in the reference: we keep the
//     minimum of key = link * M + obstacle over ALL colliding pairs (links below the current hit are never
//     dropped), then replay exactly that one pair sequentially to obtain the value the reference stores in
//     d_c_profile[link]; the pair count the reference would have visited follows from the same key.
#pragma once
#include <cuda_fp16.h>
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
constexpr int kWqSegment = 1 << 16;               // queue entries carry a 16-bit obstacle index: the table is walked in segments of 65536
#ifndef RVS_WQ_INTERVAL
#define RVS_WQ_INTERVAL 128
#endif
#ifndef RVS_WQ_MINBLOCKS
#define RVS_WQ_MINBLOCKS 6                        // 80 registers: with coherent warps 6 CTAs/SM without the spills of 72 registers beat 7 (profiles/r2)
#endif
#ifndef RVS_WQ_FLUSH
#define RVS_WQ_FLUSH 2
#endif
#ifndef RVS_WQ_EXPAND
#define RVS_WQ_EXPAND 32
#endif
#ifndef RVS_WQ_G
#define RVS_WQ_G 3
#endif
#ifndef RVS_WQ_HALF
#define RVS_WQ_HALF 1
#endif
#ifndef RVS_WQ_DFPTS
#define RVS_WQ_DFPTS 3                            // points of every link axis looked up in the distance field (ends + evenly spaced interior points)
#endif
#ifndef RVS_WQ_FIRST
#define RVS_WQ_FIRST 4                            // first refresh interval when the minima start from the distance-field bounds (without them: 4)
#endif
constexpr int kWqExpand = RVS_WQ_EXPAND;          // hit words unpacked per expansion round (bounds the pair queue: kWqExpand * 2G + 32 entries)
constexpr int kWqG = RVS_WQ_G;                    // links scanned per pass over the obstacle table (slots of the scan loop)
constexpr int kWqInterval = RVS_WQ_INTERVAL;      // obstacles between two threshold refreshes (after the first doublings 4, 8, 16, ...)
constexpr int kWqUnroll = 2;                      // obstacles tested per scan iteration
constexpr unsigned kNoHit = 0xFFFFFFFFu;
constexpr unsigned kNoEntry = 0xFFFFFFFFu;

// Q0 hit word (one per lane and scan iteration with any hit): slot mask (2*G bits: slot s x first obstacle, then slot s x second
// obstacle) | lane (5) << 12 | obstacle-pair-in-segment (15) << 17
// Q1 / Q1b / Q2 entry: lane (5) | link (5) << 5 | obstacle-in-segment (16) << 10 | edge code (6) << 26
template <int G> struct WqCaps {
    static constexpr int q0 = 64;                         // one scan iteration pushes <= 32 hit words on top of a leftover < 32
    static constexpr int q1 = kWqExpand * G * kWqUnroll + 32;   // one expansion unpacks <= kWqExpand*G*unroll pairs on top of a leftover < 32
    static constexpr int q1b = 64;                        // one re-cull round pushes <= 32 on top of a leftover < 32
    static constexpr int q2 = 32 * 12 + 32;               // one classify round pushes <= 32*12 edges on top of a leftover < 32
    static constexpr int words = 8 + q0 + q1 + q1b + q2;  // 8 control words first: 3 tails, 3 work counters
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

// bytes of dynamic shared memory for L links and scan group width G
template <class R, int G>
__host__ __device__ inline size_t wq_smem_bytes(int L) {
    size_t b = size_t(3 * (L + 1)) * kWqStride * sizeof(R);               // skeleton
    b += size_t(L) * kWqStride * sizeof(typename WqBits<R>::type);        // per-link minima
    b += size_t(L) * kWqStride * sizeof(R);                               // per-link reach (farthest axis point from the base)
    b += size_t(L) * sizeof(R);                                           // radii
    b = (b + 15) & ~size_t(15);
    b += size_t(kWqBlock) * sizeof(unsigned);                             // first-collision keys
    b += size_t(kWqWarps) * WqCaps<G>::words * sizeof(unsigned);
    return b;
}

template <class R>
struct WqCtx {
    const float4 *obs4;                       // scan-ordered obstacle table in HBM as pairs of 16-byte words (read through L1)
    R *skel;                                  // [3*(L+1)][kWqStride]
    typename WqBits<R>::type *best;           // [L][kWqStride]
    R *reach;                                 // [L][kWqStride] max(|A - base|, |B - base|) of the link's axis
    R *radius;                                // [L]
    unsigned *hit;                            // [kWqBlock]
    volatile unsigned *tails;                 // this warp: [0] Q1, [1] Q1b, [2] Q2 tails; [4], [5], [6] entries popped from each
    unsigned *q0, *q1, *q1b, *q2;             // this warp's queues
    int M, seg_base, wbase;                   // wbase = first thread of this warp in the block
    R eq_thr, max_dist;
};

// carve the block's dynamic shared memory; cheap enough to redo inside the out-of-line drain (which gets scalars in registers
// instead of a context object through local memory)
template <class R, int G>
__device__ __forceinline__ WqCtx<R> wq_make_ctx(unsigned char *smem, const float4 *obs4, int L, int M, int seg_base, R eq_thr, R max_dist) {
    typedef typename WqBits<R>::type Bits;
    const int warp = threadIdx.x >> 5;
    WqCtx<R> c;
    unsigned char *p = smem;
    c.skel = reinterpret_cast<R *>(p); p += size_t(3 * (L + 1)) * kWqStride * sizeof(R);
    c.best = reinterpret_cast<Bits *>(p); p += size_t(L) * kWqStride * sizeof(Bits);
    c.reach = reinterpret_cast<R *>(p); p += size_t(L) * kWqStride * sizeof(R);
    c.radius = reinterpret_cast<R *>(p); p += size_t(L) * sizeof(R);
    p = smem + ((size_t(p - smem) + 15) & ~size_t(15));
    c.hit = reinterpret_cast<unsigned *>(p); p += size_t(kWqBlock) * sizeof(unsigned);
    unsigned *wq = reinterpret_cast<unsigned *>(p) + size_t(warp) * WqCaps<G>::words;
    c.tails = wq; c.q0 = wq + 8; c.q1 = c.q0 + WqCaps<G>::q0; c.q1b = c.q1 + WqCaps<G>::q1; c.q2 = c.q1b + WqCaps<G>::q1b;
    c.obs4 = obs4;
    c.M = M; c.wbase = warp * 32; c.seg_base = seg_base;
    c.eq_thr = eq_thr; c.max_dist = max_dist;
    return c;
}

// scan boxes are FP32 in every variant (the obstacle table is FP32): an F64 bound is rounded OUTWARDS, which keeps the test conservative
__device__ __forceinline__ float wq_round_down(float v) { return v; }
__device__ __forceinline__ float wq_round_up(float v) { return v; }
__device__ __forceinline__ float wq_round_down(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float wq_round_up(double v) { return __double2float_ru(v); }

template <class R> __device__ __forceinline__ R wq_inf() { return R(__int_as_float(0x7f800000)); }

template <class R> __device__ __forceinline__ R wq_threshold(R best, R radius, R eq_thr) {
    const R t = best + radius;
    return (t > eq_thr ? t : eq_thr) * R(1.001) + R(1e-4);
}

// mask |= bit  iff the obstacle box [o, p] overlaps the inflated link box [lo, hi] (all comparisons false on NaN)
template <class R>
__device__ __forceinline__ void wq_overlap(unsigned &mask, unsigned bit, R ox, R oy, R oz, R px, R py, R pz, const R (&lo)[3], const R (&hi)[3]) {
    const bool ov = ox <= hi[0] && px >= lo[0] && oy <= hi[1] && py >= lo[1] && oz <= hi[2] && pz >= lo[2];
    mask |= ov ? bit : 0u;
}
// FP32: one predicate chain (6 FSETP feeding each other) and a predicated OR — 7 issue slots per pair
template <>
__device__ __forceinline__ void wq_overlap<float>(unsigned &mask, unsigned bit, float ox, float oy, float oz, float px, float py, float pz,
                                                  const float (&lo)[3], const float (&hi)[3]) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.f32 p, %1, %7;\n\t"
        "setp.ge.and.f32 p, %4, %10, p;\n\t"
        "setp.le.and.f32 p, %2, %8, p;\n\t"
        "setp.ge.and.f32 p, %5, %11, p;\n\t"
        "setp.le.and.f32 p, %3, %9, p;\n\t"
        "setp.ge.and.f32 p, %6, %12, p;\n\t"
        "@p or.b32 %0, %0, %13;\n\t}"
        : "+r"(mask)
        : "f"(ox), "f"(oy), "f"(oz), "f"(px), "f"(py), "f"(pz), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "r"(bit));
}

// ---- half-precision scan boxes -----------------------------------------------------------------------------------------------------
// The scan is a CONSERVATIVE filter (every survivor is re-tested exactly in phase B1), so its boxes may be rounded OUTWARDS to fp16:
// obstacle lo / link lo down, obstacle hi / link hi up (<= 1 mm at 1 m, 2 mm at 2-4 m: noise next to thresholds of 5-30 cm).  Two
// obstacles are packed per half2, so ONE packed compare (HSET2, on the FMA pipe) tests two pairs: 6 HSET2 + ~4 LOP3 per slot and
// obstacle pair instead of 14 FSETP / LOP, and the compare work moves off the ALU pipe, which the FP32 scan saturated (FSETP issues at
// half rate: the FP32 scan could not exceed 0.5 instructions per cycle and scheduler).
// Packed table: two uint4 per obstacle PAIR (2p, 2p+1) of the scan-ordered table:
//   w0 = (lo.x | lo.y | lo.z | hi.x)   w1 = (hi.y | hi.z | flags | 0)     each 32-bit word = half2 (obstacle 2p in the low half)
//   flags: bit 0 / bit 16 = label "ground" of obstacle 2p / 2p+1.  A missing second obstacle (odd M) is an empty box (lo = +inf, hi = -inf).
__device__ __forceinline__ unsigned wq_h2(__half a, __half b) { const __half2 v = __halves2half2(a, b); return *reinterpret_cast<const unsigned *>(&v); }
__device__ __forceinline__ unsigned wq_dup_down(float x) { const __half v = __float2half_rd(x); return wq_h2(v, v); }
__device__ __forceinline__ unsigned wq_dup_up(float x) { const __half v = __float2half_ru(x); return wq_h2(v, v); }
__device__ __forceinline__ unsigned wq_le2(unsigned a, unsigned b) { return __hle2_mask(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b)); }
__device__ __forceinline__ unsigned wq_ge2(unsigned a, unsigned b) { return __hge2_mask(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b)); }

// ---- coarse distance field (phase 0) ---------------------------------------------------------------------------------------------------
// upper bound of dist(p, nearest eligible obstacle): field value of p's cell + half a cell diagonal; +inf outside the grid
__device__ __forceinline__ float wq_df_point(const KArgs &a, float x, float y, float z) {
    const float fx = (x - a.df_origin[0]) * a.df_inv_cell, fy = (y - a.df_origin[1]) * a.df_inv_cell, fz = (z - a.df_origin[2]) * a.df_inv_cell;
    const float lim = float(a.df_dim);
    if (!(fx >= 0.f && fx < lim && fy >= 0.f && fy < lim && fz >= 0.f && fz < lim)) return __int_as_float(0x7f800000);
    const int idx = (int(fz) * a.df_dim + int(fy)) * a.df_dim + int(fx);
    return __half2float(__ldg(static_cast<const __half *>(a.df) + idx)) + a.df_slack;
}
// bound for the capsule (A, B, radius): a value >= the exact distance of the link to the obstacle that defines it, widened so that it
// stays above that distance through the roundings of the narrow phase
template <class R>
__device__ __forceinline__ R wq_df_bound(const KArgs &a, R ax, R ay, R az, R bx, R by, R bz, R radius, R max_dist) {
    const float x0 = float(ax), y0 = float(ay), z0 = float(az), x1 = float(bx), y1 = float(by), z1 = float(bz);
    float u = wq_df_point(a, x0, y0, z0);
    u = fminf(u, wq_df_point(a, x1, y1, z1));
#pragma unroll
    for (int k = 1; k < RVS_WQ_DFPTS - 1; k++) {                                   // interior points of the axis, evenly spaced
        const float w = float(k) / float(RVS_WQ_DFPTS - 1);
        u = fminf(u, wq_df_point(a, x0 + w * (x1 - x0), y0 + w * (y1 - y0), z0 + w * (z1 - z0)));
    }
    const float b = fmaxf(u - float(radius), 0.f) * 1.001f + 1e-4f;
    return b < float(max_dist) ? R(b) : max_dist;
}

#ifdef RVS_VARIANT_EXACT
// one thread per cell: distance from the cell centre to the nearest obstacle that counts for EVERY link of every configuration (no
// ground label, no min_dist_tol below MAX_DISTANCE), rounded up to fp16
__global__ void __launch_bounds__(128) df_build_kernel(const ObstacleRec *__restrict__ obs, int M, float max_dist, __half *__restrict__ df, int dim,
                                                       float ox, float oy, float oz, float cell) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dim * dim * dim) return;
    const int ix = idx % dim, iy = (idx / dim) % dim, iz = idx / (dim * dim);
    const float cx = ox + (float(ix) + 0.5f) * cell, cy = oy + (float(iy) + 0.5f) * cell, cz = oz + (float(iz) + 0.5f) * cell;
    float best = __int_as_float(0x7f800000);
#pragma unroll 4
    for (int j = 0; j < M; j++) {
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(obs + j)), w1 = __ldg(reinterpret_cast<const float4 *>(obs + j) + 1);
        if ((__float_as_uint(w1.w) & 1u) || !(w1.z >= max_dist)) continue;
        const float gx = fmaxf(fmaxf(w0.x - cx, cx - w0.w), 0.f), gy = fmaxf(fmaxf(w0.y - cy, cy - w1.x), 0.f), gz = fmaxf(fmaxf(w0.z - cz, cz - w1.y), 0.f);
        const float d2 = gx * gx + gy * gy + gz * gz;
        best = fminf(best, d2);
    }
    df[idx] = __float2half_ru(sqrtf(best) * 1.00001f + 1e-6f);
}

// one thread per obstacle pair: scan-ordered ObstacleRec table -> packed half2 table (launched after every upload / re-ranking of the table)
__global__ void __launch_bounds__(128) pack_scan_kernel(const ObstacleRec *__restrict__ scan, int M, uint4 *__restrict__ pack) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * p >= M) return;
    const float inf = __int_as_float(0x7f800000);
    const ObstacleRec a = scan[2 * p];
    ObstacleRec b;
    if (2 * p + 1 < M) b = scan[2 * p + 1];
    else { b.lo[0] = b.lo[1] = b.lo[2] = inf; b.hi[0] = b.hi[1] = b.hi[2] = -inf; b.tol = 0.f; b.flags = 0u; }
    uint4 w0, w1;
    w0.x = wq_h2(__float2half_rd(a.lo[0]), __float2half_rd(b.lo[0]));
    w0.y = wq_h2(__float2half_rd(a.lo[1]), __float2half_rd(b.lo[1]));
    w0.z = wq_h2(__float2half_rd(a.lo[2]), __float2half_rd(b.lo[2]));
    w0.w = wq_h2(__float2half_ru(a.hi[0]), __float2half_ru(b.hi[0]));
    w1.x = wq_h2(__float2half_ru(a.hi[1]), __float2half_ru(b.hi[1]));
    w1.y = wq_h2(__float2half_ru(a.hi[2]), __float2half_ru(b.hi[2]));
    w1.z = (a.flags & 1u) | ((b.flags & 1u) << 16);
    w1.w = 0u;
    pack[2 * p] = w0;
    pack[2 * p + 1] = w1;
}
#endif

// merge one finished candidate distance (radius already subtracted) of pair `key` into configuration slot `t`
template <class R>
__device__ __forceinline__ void wq_merge(const WqCtx<R> &c, int t, int link, unsigned key, R d, R tol) {
    if (d > tol) return;                                                   // min_dist_tol: contributes MAX_DISTANCE (:367-368)
    if (d <= R(0)) atomicMin(&c.hit[t], key);                              // a collision of this pair (:370)
    else if (d > R(0)) atomicMin(&c.best[link * kWqStride + t], WqBits<R>::to(d));
}

// entry -> configuration slot, link, capsule axis, obstacle box / tolerance and the pair's key in REFERENCE order
// (link-major, obstacle index in the caller's order: the scan table is sorted, its records carry the original index)
template <class R>
struct WqPair {
    int t, link;
    unsigned key;
    V3<R> A, B;
    Box<R> bx;
    R tol;
};
template <class R>
__device__ __forceinline__ WqPair<R> wq_load_pair(const WqCtx<R> &c, unsigned e) {
    WqPair<R> p;
    p.t = c.wbase + int(e & 31u);
    p.link = int((e >> 5) & 31u);
    const int j = c.seg_base + int((e >> 10) & 0xFFFFu);
    const R *sk = c.skel + (3 * p.link) * kWqStride + p.t;
    p.A = {sk[0], sk[kWqStride], sk[2 * kWqStride]};
    p.B = {sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride]};
    const float4 w0 = __ldg(c.obs4 + 2 * j), w1 = __ldg(c.obs4 + 2 * j + 1);
    p.bx = {{R(w0.x), R(w0.y), R(w0.z)}, {R(w0.w), R(w1.x), R(w1.y)}};
    p.tol = R(w1.z);
    p.key = unsigned(p.link) * unsigned(c.M) + (__float_as_uint(w1.w) >> 8);
    return p;
}

// phase B1: does this pair still matter?  Euclidean gap between the segment's bounding box and the obstacle against the
// CURRENT minimum of its link (the scan used the Chebyshev gap and a possibly older minimum).
template <class R>
__device__ __forceinline__ bool wq_recull(const WqCtx<R> &c, unsigned e) {
    const WqPair<R> p = wq_load_pair(c, e);
    const unsigned hk = c.hit[p.t];
    if (p.key > hk) return false;                                          // an earlier pair (reference order) already collides
    // on the link of the current first collision only an earlier COLLISION matters (its minimum is not an output)
    const bool hit_link = hk < unsigned(p.link + 1) * unsigned(c.M);
    const R thr = wq_threshold(hit_link ? R(0) : WqBits<R>::from(c.best[p.link * kWqStride + p.t]), c.radius[p.link], c.eq_thr);
    const R gx = fmax(fmax(p.bx.lo.x - fmax(p.A.x, p.B.x), fmin(p.A.x, p.B.x) - p.bx.hi.x), R(0));
    const R gy = fmax(fmax(p.bx.lo.y - fmax(p.A.y, p.B.y), fmin(p.A.y, p.B.y) - p.bx.hi.y), R(0));
    const R gz = fmax(fmax(p.bx.lo.z - fmax(p.A.z, p.B.z), fmin(p.A.z, p.B.z) - p.bx.hi.z), R(0));
    return !(gx * gx + gy * gy + gz * gz > thr * thr);
}

// phase B2: one surviving pair through the reference's classification
template <class R>
__device__ __forceinline__ void wq_classify(const WqCtx<R> &c, unsigned e) {
    const WqPair<R> p = wq_load_pair(c, e);
    if (p.key > c.hit[p.t]) return;
    Capsule<R> cp;
    cp.set(p.A, p.B, c.radius[p.link], R(0));
    V3<R> u, v;
    const PairPlan<R> pl = classify_pair<R, false>(cp, p.bx, c.eq_thr, c.max_dist, u, v);
    if (pl.done) { wq_merge(c, p.t, p.link, p.key, pl.d, p.tol); return; }
    wq_merge(c, p.t, p.link, p.key, pl.d - cp.radius, p.tol);              // face distance (2a) or MAX_DISTANCE (2b)
    const unsigned pos = atomicAdd(const_cast<unsigned *>(&c.tails[2]), unsigned(pl.count));
#pragma unroll 1
    for (int k = 0; k < pl.count; k++) c.q2[pos + k] = (e & 0x3FFFFFFu) | (plan_edge(pl, k) << 26);
}

// phase C: one box edge of one pair
template <class R>
__device__ __forceinline__ void wq_edge(const WqCtx<R> &c, unsigned e) {
    const WqPair<R> p = wq_load_pair(c, e);
    if (p.key > c.hit[p.t]) return;
    Capsule<R> cp;
    cp.set(p.A, p.B, c.radius[p.link], R(0));
    V3<R> u, v;
    const R de = plan_edge_distance<R, false>(cp, p.bx, e >> 26, c.eq_thr, c.max_dist, u, v);
    wq_merge(c, p.t, p.link, p.key, de <= R(0) ? R(0) - cp.radius : de - cp.radius, p.tol);     // Q14: a touching edge gives -radius
}

// pop up to 32 entries from the top of a queue (warp-uniform); lanes beyond the count get kNoEntry
__device__ __forceinline__ unsigned wq_pop(volatile unsigned *tail, const unsigned *q, unsigned have, int lane) {
    const unsigned n = have < 32u ? have : 32u, base = have - n;
    const unsigned e = unsigned(lane) < n ? q[base + lane] : kNoEntry;
    __syncwarp();
    if (lane == 0) { *tail = base; tail[4] += n; }
    return e;
}

// phase A': unpack up to 32 hit words of the scan (one lane each) into pair entries of Q1.  The scan pushes ONE word per lane and
// iteration (a predicated store at a ballot-computed position); turning its bits into pairs here, with every lane busy, is ~4x
// cheaper than a divergent per-lane loop inside the scan.  `n0` (Q0 fill, warp-uniform, kept in a register) is returned reduced.
template <class R, int G>
__device__ __forceinline__ unsigned wq_expand(const WqCtx<R> &c, unsigned n0, unsigned slotmap, int lane) {
    const unsigned n = n0 < unsigned(kWqExpand) ? n0 : unsigned(kWqExpand), base0 = n0 - n;
    const unsigned w = unsigned(lane) < n ? c.q0[base0 + lane] : 0u;
    unsigned m = w & ((1u << (2 * G)) - 1u);
    const unsigned cnt = unsigned(__popc(m));
    unsigned incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned t1 = c.tails[0];
    unsigned pos = t1 + incl - cnt;
    const unsigned from = (w >> 12) & 31u, j = (w >> 17) << 1;
    while (m) {
        const int b = __ffs(int(m)) - 1;
        m &= m - 1u;
        const int sl = b >= G ? b - G : b;
        c.q1[pos++] = from | (((slotmap >> (5 * sl)) & 31u) << 5) | ((j + (b >= G ? 1u : 0u)) << 10);
    }
    __syncwarp();
    if (lane == 0) c.tails[0] = t1 + total;
    return base0;
}

// run compacted phases while a queue holds a full warp of work.  flush = 1: also unpack a partly filled Q0 (its words refer to
// the CURRENT slot -> link map, so the scan empties it before the map changes); flush = 2: empty every queue.
// Inlined at exactly ONE call site (the scan loop is arranged for that): the compacted phases are ~1500 instructions, and several
// copies made the kernel outgrow the instruction cache (ncu: 31 % of the stall samples were `no_instruction`); an out-of-line
// copy costs ~45 register spills around every call instead.
template <class R, int G>
__device__ __forceinline__ unsigned wq_drain(const WqCtx<R> &c, int flush, unsigned n0, unsigned slotmap) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        __syncwarp();
        const unsigned t1 = c.tails[0], t1b = c.tails[1], t2 = c.tails[2];
        if (t2 >= 32u || (flush == 2 && t2 > 0u && n0 == 0u && t1 == 0u && t1b == 0u)) {
            const unsigned e = wq_pop(&c.tails[2], c.q2, t2, lane);
            if (e != kNoEntry) wq_edge<R>(c, e);
        } else if (t1b >= 32u || (flush == 2 && t1b > 0u && n0 == 0u && t1 == 0u)) {
            const unsigned e = wq_pop(&c.tails[1], c.q1b, t1b, lane);
            if (e != kNoEntry) wq_classify<R>(c, e);
        } else if (t1 >= 32u || (flush == 2 && t1 > 0u && n0 == 0u)) {
            const unsigned e = wq_pop(&c.tails[0], c.q1, t1, lane);
            const bool keep = e != kNoEntry && wq_recull<R>(c, e);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (keep) c.q1b[t1b + __popc(b & ((1u << lane) - 1u))] = e;
            if (lane == 0) c.tails[1] = t1b + __popc(b);
        } else if (n0 >= 32u || (flush && n0 > 0u)) {
            n0 = wq_expand<R, G>(c, n0, slotmap, lane);
        } else
            break;
    }
    return n0;
}

// the first-collision replay of the epilogue runs once per colliding configuration: out of line as well
template <class R>
__device__ __noinline__ R wq_replay_pair(R ax, R ay, R az, R bx_, R by, R bz, R radius, const ObstacleRec *o, R eq_thr, R max_dist) {
    Capsule<R> cp;
    cp.set({ax, ay, az}, {bx_, by, bz}, radius, R(0));
    V3<R> u, v;
    return capsule_box_distance<R, false>(cp, load_box<R>(*o), eq_thr, max_dist, u, v);
}

// F = FK scalar, R = geometry scalar, MODE = RVS_DIST_MIN | RVS_DIST_PROFILE, G = links scanned per obstacle pass.
template <class F, class R, int MODE, int G>
__global__ void __launch_bounds__(kWqBlock, sizeof(R) == 4 ? RVS_WQ_MINBLOCKS : 3) distance_wq_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char wq_smem[];
    static_assert(G >= 1 && G <= 6, "the slot -> link map packs at most six 5-bit link indices");
    typedef typename WqBits<R>::type Bits;
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n;
    const int lane = threadIdx.x & 31;

    WqCtx<R> c = wq_make_ctx<R, G>(wq_smem, reinterpret_cast<const float4 *>(a.obstacles_scan), L, a.M, 0, R(a.eq_thr), R(a.max_dist));

    const int64_t slot_k = int64_t(blockIdx.x) * kWqBlock + threadIdx.x;
    const bool live = slot_k < a.N;
    const int64_t cfg = live && a.perm ? int64_t(__ldg(a.perm + slot_k)) : slot_k;     // coherent warps: see rvs_sort.cuh
    const int t = threadIdx.x;
    if (live) fk_skeleton<F, R>(rb, a.q + cfg * L, c.skel + t, kWqStride);
    else for (int k = 0; k < 3 * (L + 1); k++) c.skel[k * kWqStride + t] = R(0);
    // reach of every link: no point of its axis is farther from the base than this, so an obstacle at distance d from the base
    // is at least d - reach away from the link (the scan table is sorted by d: see the cut-off below)
    {
        const R bx = R(rb->xyz[0][0]), by = R(rb->xyz[0][1]), bz = R(rb->xyz[0][2]);
        R prev = 0;
        for (int i = 0; i <= L; i++) {
            const R x = c.skel[(3 * i) * kWqStride + t] - bx, y = c.skel[(3 * i + 1) * kWqStride + t] - by, z = c.skel[(3 * i + 2) * kWqStride + t] - bz;
            const R nrm = sqrt(x * x + y * y + z * z);
            if (i > 0) c.reach[(i - 1) * kWqStride + t] = fmax(prev, nrm);
            prev = nrm;
        }
    }
    // phase 0: every link's running minimum starts at the distance-field bound (an upper bound of the answer), or at MAX_DISTANCE
    for (int i = 0; i < L; i++) {
        R b0 = c.max_dist;
        if (a.df && live) {
            const R *sk = c.skel + (3 * i) * kWqStride + t;
            b0 = wq_df_bound<R>(a, sk[0], sk[kWqStride], sk[2 * kWqStride], sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride], RobotScalars<R>::radius(rb, i), c.max_dist);
        }
        c.best[i * kWqStride + t] = WqBits<R>::to(b0);
    }
    c.hit[t] = kNoHit;
    if (t < L) c.radius[t] = RobotScalars<R>::radius(rb, t);
    if (lane < 8) c.tails[lane] = 0;
    const int ground_included = rb->ground_included;
    __syncthreads();                                                       // radius[] is shared by the block; everything else is per warp

    // pass 0 scans with the bootstrap bounds in place; pass 1 (normally never entered) re-scans the links whose minimum still EQUALS its
    // bound, i.e. for which no candidate came below it, without bounds — see "Exactness" in the header
    bool scanning = live;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
    for (int base = 0; base < a.M; base += kWqSegment) {
        const int count = min(kWqSegment, a.M - base);
        c.seg_base = base;
        const float4 *o4 = c.obs4 + 2 * size_t(base);
        const uint4 *op = static_cast<const uint4 *>(a.obstacles_pack) + size_t(base);                 // two uint4 per obstacle pair; base is even
        const float *sd = a.scan_dist + base;
#pragma unroll 1
        for (int g0 = 0; g0 < L; g0 += G) {
            // links of this group that some configuration of the warp can still improve (warp-uniform, only ever shrinks)
            unsigned alive = (1u << min(G, L - g0)) - 1u;
            unsigned slotmap = 0, n0 = 0;                                  // slot -> link map (5 bits per slot); hit words waiting in Q0
#pragma unroll 1
            for (int j0 = 0; j0 < count;) {
                // Cut-off: the table is sorted by distance from the base, so once  d(j0) - reach > threshold  for a link, neither
                // obstacle j0 nor any later one can lower its minimum.  Links at or above a known collision are dead as well.
                const unsigned hk = c.hit[t];
                const int hit_link = hk == kNoHit ? -1 : int(hk / unsigned(a.M));
                const int limit = !scanning ? 0 : (hk == kNoHit ? L : hit_link + 1);
                const R dnext = R(__ldg(sd + j0));
                unsigned mine = 0;
#pragma unroll
                for (int l = 0; l < G; l++) {
                    const int i = g0 + l;
                    if (i < limit && i < L) {
                        // the link of the current first collision only looks for EARLIER collisions (its minimum is not an output)
                        const R thr = wq_threshold(i == hit_link ? R(0) : WqBits<R>::from(c.best[i * kWqStride + t]), c.radius[i], c.eq_thr);
                        if (!(dnext > (c.reach[i * kWqStride + t] + thr) * R(1.00001))) mine |= 1u << l;
                    }
                }
                alive &= __reduce_or_sync(0xffffffffu, mine);
                // No live link left: one empty iteration over the tail of the table, so that the end-of-segment flush below (the only
                // inlined copy of the compacted phases) still runs.
                if (!alive) j0 = (count - 1) & ~1;
                // live links are packed into slots 0..K-1 (K, the slot -> link map and the ground filter are warp-uniform);
                // slot boxes = link bounding boxes inflated by the current threshold, empty for lanes that are done with the link
                const int K = __popc(alive);
                unsigned ground_keep = 0;
                slotmap = 0;                                                // (Q0 is empty here: every interval ends with a flush)
#if RVS_WQ_HALF
                unsigned lo[G][3], hi[G][3];                                // half2, the link box value in both halves (rounded outwards)
#else
                float lo[G][3], hi[G][3];
#endif
                {
                    unsigned rem = alive;
#pragma unroll
                    for (int s = 0; s < G; s++) {
                        const int l = __ffs(int(rem)) - 1, i = g0 + l;
                        rem &= rem - 1u;
                        if (s < K && ((mine >> l) & 1u)) {
                            const R *sk = c.skel + (3 * i) * kWqStride + t;
                            const R thr = wq_threshold(i == hit_link ? R(0) : WqBits<R>::from(c.best[i * kWqStride + t]), c.radius[i], c.eq_thr);
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const R p = sk[k * kWqStride], q = sk[(3 + k) * kWqStride];
#if RVS_WQ_HALF
                                lo[s][k] = wq_dup_down(wq_round_down(fmin(p, q) - thr));
                                hi[s][k] = wq_dup_up(wq_round_up(fmax(p, q) + thr));
#else
                                lo[s][k] = wq_round_down(fmin(p, q) - thr);
                                hi[s][k] = wq_round_up(fmax(p, q) + thr);
#endif
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 3; k++) {                                      // overlaps nothing
#if RVS_WQ_HALF
                                lo[s][k] = 0x7C007C00u; hi[s][k] = 0xFC00FC00u;                // (+inf, +inf) / (-inf, -inf) in half2
#else
                                lo[s][k] = wq_inf<float>(); hi[s][k] = -wq_inf<float>();
#endif
                            }
                        }
                        if (s < K) {
                            slotmap |= unsigned(i) << (5 * s);
                            if (i >= ground_included) ground_keep |= 1u << s;                 // links below ground_included ignore ground objects (:336)
                        }
                    }
#if !RVS_WQ_HALF
                    ground_keep *= (1u << G) + 1u;                                            // replicated for both unrolled obstacles
#endif
                }
                // the minima settle quickly (the table is sorted nearest-first): refresh after 4, 8, 16, 32, 64, 128, 256, 384, ... obstacles
                const int first = a.df ? RVS_WQ_FIRST : 4;
                const int jend = alive ? min(j0 + (j0 < first ? first : min(j0, kWqInterval)), count) : count;
                const bool seg_end = jend == count && g0 + G >= L;         // queue entries index this segment: empty every queue before moving on
#pragma unroll 1
                for (int j = j0; j < jend; j += kWqUnroll) {
#if RVS_WQ_HALF
                    // two obstacles per packed record (see pack_scan_kernel): two warp-uniform 16-byte loads, 6 packed compares per slot
                    const uint4 w0 = __ldg(op + (j)), w1 = __ldg(op + (j) + 1);
                    unsigned m2 = 0;                                                          // bit S: obstacle j x slot S, bit 16 + S: obstacle j + 1 x slot S
#define RVS_WQ_SLOT(S)                                                                                                             \
    if (S < G) {                                                                                                                   \
        const unsigned ax = wq_le2(w0.x, hi[S < G ? S : 0][0]) & wq_ge2(w0.w, lo[S < G ? S : 0][0]) & wq_le2(w0.y, hi[S < G ? S : 0][1]);      \
        const unsigned bx = wq_ge2(w1.x, lo[S < G ? S : 0][1]) & wq_le2(w0.z, hi[S < G ? S : 0][2]) & wq_ge2(w1.y, lo[S < G ? S : 0][2]);      \
        m2 |= ax & bx & (0x00010001u << (S));                                                                                      \
    }
                    switch (K) {                                                              // warp-uniform: only live slots are tested
                        case 0: break;
                        default: RVS_WQ_SLOT(5) [[fallthrough]];
                        case 5: RVS_WQ_SLOT(4) [[fallthrough]];
                        case 4: RVS_WQ_SLOT(3) [[fallthrough]];
                        case 3: RVS_WQ_SLOT(2) [[fallthrough]];
                        case 2: RVS_WQ_SLOT(1) [[fallthrough]];
                        case 1: RVS_WQ_SLOT(0)
                    }
#undef RVS_WQ_SLOT
                    if (w1.z) m2 &= ((w1.z & 1u) ? (ground_keep | 0xFFFF0000u) : ~0u) & ((w1.z & 0x10000u) ? ((ground_keep << 16) | 0x0000FFFFu) : ~0u);
                    const unsigned mask = (m2 & 0xFFFFu) | ((m2 >> 16) << G);                 // hit-word layout: slot bits of obstacle j, then of obstacle j + 1
#else
                    // (lo.xyz, hi.x) (hi.y, hi.z, tol, flags) of two obstacles: warp-uniform 16-byte loads
                    const int j1 = j + 1 < jend ? j + 1 : j;
                    const float4 *pj = o4 + 2 * j, *pj1 = pj + (j1 - j) * 2;
                    const float4 w0 = __ldg(pj), w1 = __ldg(pj + 1), v0 = __ldg(pj1), v1 = __ldg(pj1 + 1);
                    unsigned mask = 0;
#define RVS_WQ_SLOT(S)                                                                                                             \
    if (S < G) {                                                                                                                   \
        wq_overlap<float>(mask, 1u << (S), w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, lo[S < G ? S : 0], hi[S < G ? S : 0]); \
        wq_overlap<float>(mask, 1u << (G + (S)), v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, lo[S < G ? S : 0], hi[S < G ? S : 0]); \
    }
                    switch (K) {                                                              // warp-uniform: only live slots are tested
                        case 0: break;
                        default: RVS_WQ_SLOT(5) [[fallthrough]];
                        case 5: RVS_WQ_SLOT(4) [[fallthrough]];
                        case 4: RVS_WQ_SLOT(3) [[fallthrough]];
                        case 3: RVS_WQ_SLOT(2) [[fallthrough]];
                        case 2: RVS_WQ_SLOT(1) [[fallthrough]];
                        case 1: RVS_WQ_SLOT(0)
                    }
#undef RVS_WQ_SLOT
                    if (j1 == j) mask &= (1u << G) - 1u;                                    // odd tail: the second slot is a duplicate
                    if ((__float_as_uint(w1.w) | __float_as_uint(v1.w)) & 1u)
                        mask &= ((__float_as_uint(w1.w) & 1u) ? (ground_keep | ~((1u << G) - 1u)) : ~0u) &
                                ((__float_as_uint(v1.w) & 1u) ? (ground_keep | ((1u << G) - 1u)) : ~0u);
#endif
                    // one hit word per lane with any hit, at a ballot-computed position of Q0 (no divergence in the scan)
                    const unsigned hits = __ballot_sync(0xffffffffu, mask != 0u);
                    if (mask) c.q0[n0 + __popc(hits & ((1u << lane) - 1u))] = mask | (unsigned(lane) << 12) | (unsigned(j >> 1) << 17);
                    n0 += unsigned(__popc(hits));
                    // hit words refer to the current slot -> link map: the last iteration of an interval unpacks all of them
                    const bool last = j + kWqUnroll >= jend;
                    if (n0 >= 32u || last) n0 = wq_drain<R, G>(c, last ? (seg_end ? 2 : RVS_WQ_FLUSH) : 0, n0, slotmap);      // interval end: every queue is emptied, so the next thresholds are exact
                }
                j0 = jend;
            }
        }
    }
    __syncwarp();
    if (pass == 1 || !a.df) break;
    bool redo = false;
    if (live) {
        const unsigned hk0 = c.hit[t];
        const int outputs = hk0 == kNoHit ? L : int(hk0 / unsigned(a.M));     // links below the first collision: their minima are outputs
        for (int i = 0; i < outputs; i++) {
            const R *sk = c.skel + (3 * i) * kWqStride + t;
            const R b0 = wq_df_bound<R>(a, sk[0], sk[kWqStride], sk[2 * kWqStride], sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride], RobotScalars<R>::radius(rb, i), c.max_dist);
            if (b0 < c.max_dist && WqBits<R>::from(c.best[i * kWqStride + t]) == b0) { c.best[i * kWqStride + t] = WqBits<R>::to(c.max_dist); redo = true; }
        }
    }
    if (!__any_sync(0xffffffffu, redo)) break;
    if (a.pair_count && lane == 0) atomicAdd(a.pair_count + 4, 1ull);            // (work statistics: warps that took the second pass)
    scanning = redo;
    __syncwarp();
    }

    // ---- per-configuration epilogue (lane = configuration again) ---------------------------------------------------
    const unsigned hk = c.hit[t];
    const bool collided = live && hk != kNoHit;
    const int limit = collided ? int(hk / unsigned(a.M)) : L;
    const int hit_j = collided ? int(hk % unsigned(a.M)) : 0;
    R hit_d = 0;
    if (collided) {
        // the value the reference stores for the colliding link: replay exactly that pair sequentially
        const R *sk = c.skel + (3 * limit) * kWqStride + t;
        hit_d = wq_replay_pair<R>(sk[0], sk[kWqStride], sk[2 * kWqStride], sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride], c.radius[limit],
                                  a.obstacles + hit_j, c.eq_thr, c.max_dist);
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
        if (lane >= 1 && lane <= 3 && c.tails[3 + lane]) atomicAdd(a.pair_count + lane, (unsigned long long)c.tails[3 + lane]);
    }
    if (!live) return;
    R d_c = c.max_dist;
    for (int i = 0; i < L; i++) { const R b = WqBits<R>::from(c.best[i * kWqStride + t]); if (b < d_c) d_c = b; }
    static_cast<R *>(a.d_c)[cfg] = collided ? R(0) : d_c;
    if (MODE >= RVS_DIST_PROFILE) {
        R *out = static_cast<R *>(a.profile) + cfg * L;
        for (int i = 0; i < L; i++) {
            const R b = WqBits<R>::from(c.best[i * kWqStride + t]);
            out[i] = !collided ? b : (i < limit ? b : (i == limit ? hit_d : R(0)));
        }
    }
}

}  // namespace rvs
