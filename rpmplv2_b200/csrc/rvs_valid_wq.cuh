// rvs_valid_wq.cuh — isValid(q) and isValid(q1,q2) (RealVectorSpace.cpp:261-309) as  scan -> warp work queue -> compacted predicate.
//
// The reference's predicate collisionCapsuleToBox (CollisionAndDistance.cpp:4-42) starts with a slab test: a pair is rejected when
// both capsule endpoints lie beyond the box by more than the radius on one axis. The scan applies a slightly LOOSER version of that
// test (link bounding box inflated by radius * 1.001 + margin, six predicate-chained compares per pair, lane = configuration, no
// divergence); the ~1 % of pairs that survive are queued and evaluated with the exact reference predicate capsule_hits_box() — which
// contains the exact slab test — by full warps (lane = queued pair of any configuration of the warp). A configuration is invalid as
// soon as one of its pairs hits: the verdict is order free, so it is identical to the sequential kernel's.
// EDGE = true: lane = one interpolated configuration of an edge (k = N..1, RealVectorSpace.cpp:267-269); verdict[edge] &= valid.
#pragma once
#include "rvs_distance_wq.cuh"

namespace rvs {

template <int G> struct VqCaps {
    static constexpr int q1 = 32 * G * kWqUnroll + 32;   // pairs: one scan iteration pushes <= 32*G*unroll on top of a leftover < 32
    static constexpr int q2 = 32 * 12 + 32;               // box edges: one classification round pushes <= 32*12 on top of a leftover < 32
    static constexpr int words = 4 + q1 + q2;             // control words first: [0] pair tail, [1] edge tail
};

template <class R, int G>
__host__ __device__ inline size_t vq_smem_bytes(int L) {
    size_t b = size_t(3 * (L + 1)) * kWqStride * sizeof(R);               // skeleton
    b += size_t(L) * kWqStride * sizeof(R);                               // per-link cut-off distance from the base (reach + slab threshold)
    b += 2 * size_t(L) * sizeof(R);                                       // radius, radius * sqrt(3) / 3
    b = (b + 15) & ~size_t(15);
    b += size_t(kWqBlock) * sizeof(unsigned);                             // invalid flags
    b += size_t(kWqWarps) * VqCaps<G>::words * sizeof(unsigned);
    return b;
}

#ifndef RVS_VQ_MINBLOCKS
#define RVS_VQ_MINBLOCKS 5
#endif
template <class F, class R, int G, bool EDGE>
__global__ void __launch_bounds__(kWqBlock, sizeof(R) == 4 ? RVS_VQ_MINBLOCKS : 3) valid_wq_kernel(const KArgs a) {
    extern __shared__ __align__(16) unsigned char vq_smem[];
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n, n = L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = threadIdx.x, wbase = warp * 32;
    unsigned char *p = vq_smem;
    R *skel = reinterpret_cast<R *>(p); p += size_t(3 * (L + 1)) * kWqStride * sizeof(R);
    R *cut = reinterpret_cast<R *>(p); p += size_t(L) * kWqStride * sizeof(R);
    R *radius = reinterpret_cast<R *>(p); p += size_t(L) * sizeof(R);
    R *radius_in = reinterpret_cast<R *>(p); p += size_t(L) * sizeof(R);
    p = vq_smem + ((size_t(p - vq_smem) + 15) & ~size_t(15));
    unsigned *invalid = reinterpret_cast<unsigned *>(p); p += size_t(kWqBlock) * sizeof(unsigned);
    unsigned *wq = reinterpret_cast<unsigned *>(p) + size_t(warp) * VqCaps<G>::words;
    volatile unsigned *tail = wq;
    unsigned *q1 = wq + 4, *q2 = q1 + VqCaps<G>::q1;
    const float4 *obs4 = reinterpret_cast<const float4 *>(a.obstacles_scan);   // nearest to the base first (the verdict is order free)
    const R eq_thr = R(a.eq_thr), max_dist = R(a.max_dist);
    int seg_base = 0;

    // a.perm (rvs_sort.cuh): thread k works on configuration / edge perm[k], so that the configurations of a warp are alike
    const int64_t slot_k = int64_t(blockIdx.x) * kWqBlock + threadIdx.x;
    int64_t item = slot_k, edge = 0;
    bool live;
    if (!EDGE) {
        live = slot_k < a.N;
        if (live && a.perm) item = int64_t(__ldg(a.perm + slot_k));
        if (live) fk_skeleton<F, R>(rb, a.q + item * n, skel + t, kWqStride);
    } else {
        edge = slot_k / a.num_checks;
        const int kidx = int(slot_k - edge * a.num_checks);
        live = edge < a.N;
        if (live && a.perm) edge = int64_t(__ldg(a.perm + edge));
        if (live) {
            const float *q1p = a.q + edge * n, *q2p = a.q2 + edge * n;
            const float dist = coord_norm(q1p, q2p, n);
            const float k = float(a.num_checks - kidx);                 // k = N .. 1
            const float step = k / float(a.num_checks) * dist;
            float qn[kMaxDofs];
            if (dist > 0.f) {
                const float f = step / dist;
                for (int i = 0; i < n; i++) qn[i] = q1p[i] + (q2p[i] - q1p[i]) * f;
            } else
                for (int i = 0; i < n; i++) qn[i] = q2p[i];
            fk_skeleton<F, R>(rb, qn, skel + t, kWqStride);
        }
    }
    if (!live) for (int k = 0; k < 3 * (L + 1); k++) skel[k * kWqStride + t] = R(0);
    {
        // slab threshold (looser than the reference's `+- radius`) and the distance from the base beyond which no obstacle can
        // touch the link: reach (farthest axis point from the base) + threshold.  The scan table is sorted by that distance.
        const R bx = R(rb->xyz[0][0]), by = R(rb->xyz[0][1]), bz = R(rb->xyz[0][2]);
        R prev = 0;
        for (int i = 0; i <= L; i++) {
            const R x = skel[(3 * i) * kWqStride + t] - bx, y = skel[(3 * i + 1) * kWqStride + t] - by, z = skel[(3 * i + 2) * kWqStride + t] - bz;
            const R nrm = sqrt(x * x + y * y + z * z);
            if (i > 0) {
                const R *sk = skel + (3 * (i - 1)) * kWqStride + t;
                R mag = R(0);
#pragma unroll
                for (int k = 0; k < 3; k++) { const R u = sk[k * kWqStride], v = sk[(3 + k) * kWqStride]; mag = fmax(mag, fmax(fabs(u), fabs(v))); }
                const R thr = RobotScalars<R>::radius(rb, i - 1) * R(1.001) + R(1e-5) + R(4e-6) * mag;
                cut[(i - 1) * kWqStride + t] = (fmax(prev, nrm) + fmax(R(1.7320509) * thr, eq_thr * R(1.001) + R(1e-5))) * R(1.00001);
            }
            prev = nrm;
        }
    }
    invalid[t] = 0u;
    if (t < L) { radius[t] = RobotScalars<R>::radius(rb, t); radius_in[t] = RobotScalars<R>::radius_in(rb, t); }
    if (lane < 4) tail[lane] = 0;
    const int ground_included = rb->ground_included;
    __syncthreads();

    // The exact reference predicate in two compacted phases (lane = queued item of ANY configuration of the warp):
    //   pairs  hit_box_plan(): endpoint accept, slab reject, crossing / projection tests; decided collisions are recorded, the box
    //          edges the reference would measure go to the edge queue
    //   edges  one segment-edge distance per lane: closer than the radius = collision
    // (the one-piece predicate ran its segment-edge evaluations with ~2 of 32 lanes active and, inlined a dozen times, made the
    // kernel 126 KB of SASS: 71 % of the stall samples were instruction-cache misses on dense scenes, profiles/README.md)
    auto load_item = [&](unsigned e, int &tt, int &link, Capsule<R> &cp, Box<R> &bx) {
        tt = wbase + int(e & 31u);
        link = int((e >> 5) & 31u);
        const int j = seg_base + int((e >> 10) & 0xFFFFu);
        const R *sk = skel + (3 * link) * kWqStride + tt;
        cp.set({sk[0], sk[kWqStride], sk[2 * kWqStride]}, {sk[3 * kWqStride], sk[4 * kWqStride], sk[5 * kWqStride]}, radius[link], radius_in[link]);
        const float4 w0 = __ldg(obs4 + 2 * j), w1 = __ldg(obs4 + 2 * j + 1);
        bx = {{R(w0.x), R(w0.y), R(w0.z)}, {R(w0.w), R(w1.x), R(w1.y)}};
    };
    auto drain = [&](bool flush) {
        for (;;) {
            __syncwarp();
            const unsigned t1 = tail[0], t2 = tail[1];
            if (t2 >= 32u || (flush && t2 > 0u && t1 == 0u)) {
                const unsigned cnt = t2 < 32u ? t2 : 32u, base = t2 - cnt;
                const unsigned e = unsigned(lane) < cnt ? q2[base + lane] : kNoEntry;
                __syncwarp();
                if (lane == 0) tail[1] = base;
                if (e != kNoEntry) {
                    int tt, link; Capsule<R> cp; Box<R> bx;
                    load_item(e, tt, link, cp, bx);
                    if (!atomicOr(&invalid[tt], 0u)) {                            // (the flag only goes 0 -> 1; atomics keep the polling well defined)
                        V3<R> u, v;
                        if (plan_edge_distance<R, false>(cp, bx, e >> 26, eq_thr, max_dist, u, v) < cp.radius) atomicOr(&invalid[tt], 1u);
                    }
                }
            } else if (t1 >= 32u || (flush && t1 > 0u)) {
                const unsigned cnt = t1 < 32u ? t1 : 32u, base = t1 - cnt;
                const unsigned e = unsigned(lane) < cnt ? q1[base + lane] : kNoEntry;
                __syncwarp();
                if (lane == 0) tail[0] = base;
                if (e != kNoEntry) {
                    int tt, link; Capsule<R> cp; Box<R> bx;
                    load_item(e, tt, link, cp, bx);
                    if (!atomicOr(&invalid[tt], 0u)) {                            // (the flag only goes 0 -> 1; atomics keep the polling well defined)
                        // Cheap exact-safe reject in front of the predicate: every `true` of collisionCapsuleToBox implies an axis point
                        // within sqrt(2) * radius of the box (endpoint in the box inflated by r/sqrt(3); crossing point inside the rectangle
                        // oversized by r; an edge or face distance < r), hence within sqrt(2) r + (half diagonal) of the box centre.  A
                        // diagonal link's bounding box passes the slab test for many small boxes that are nowhere near the segment.
                        const R cx = (bx.lo.x + bx.hi.x) * R(0.5), cy = (bx.lo.y + bx.hi.y) * R(0.5), cz = (bx.lo.z + bx.hi.z) * R(0.5);
                        const R hx = (bx.hi.x - bx.lo.x) * R(0.5), hy = (bx.hi.y - bx.lo.y) * R(0.5), hz = (bx.hi.z - bx.lo.z) * R(0.5);
                        const R ux = cp.B.x - cp.A.x, uy = cp.B.y - cp.A.y, uz = cp.B.z - cp.A.z;
                        const R vx = cx - cp.A.x, vy = cy - cp.A.y, vz = cz - cp.A.z;
                        const R uu = ux * ux + uy * uy + uz * uz, uv = ux * vx + uy * vy + uz * vz, vv = vx * vx + vy * vy + vz * vz;
                        R d2;                                                       // squared distance of the box centre from the segment
                        if (!(uv > R(0))) d2 = vv;
                        else if (uv >= uu) { const R wx = cx - cp.B.x, wy = cy - cp.B.y, wz = cz - cp.B.z; d2 = wx * wx + wy * wy + wz * wz; }
                        else d2 = vv - uv * uv / uu;
                        const R reach = sqrt(hx * hx + hy * hy + hz * hz) + fmax(R(1.41422) * radius[link], eq_thr) * R(1.001) + R(1e-4);      // (a segment-edge distance below EQUALITY_THRESHOLD snaps to 0 = hit)
                        const bool far = d2 - R(4e-6) * (vv + R(1)) > reach * reach * R(1.002);     // (NaN anywhere: not far, the predicate decides)
                        if (!far) {
                            HitEdges edges;
                            if (hit_box_plan(cp, bx, max_dist, edges)) atomicOr(&invalid[tt], 1u);
                            else if (edges.count) {
                                const unsigned pos = atomicAdd(const_cast<unsigned *>(&tail[1]), unsigned(edges.count));
#pragma unroll 1
                                for (int k = 0; k < edges.count; k++) q2[pos + k] = (e & 0x3FFFFFFu) | (edges.at(k) << 26);
                            }
                        }
                    }
                }
            } else
                break;
        }
    };

    // Mid-size batches (a planner's edge batch against a thousand boxes) leave most SMs idle while every warp walks the whole table:
    // the table is split over gridDim.y slices (multiples of 64 obstacles); the verdict is the AND over the slices (pre-set to 1 by
    // the host, cleared by any slice that finds a collision).  The table is sorted, so the far slices end at their first cut-off test.
    const int slices = int(gridDim.y);
    const int per = (((a.M + slices - 1) / slices) + 63) & ~63;
    const int j_lo = int(blockIdx.y) * per, j_hi = min(a.M, j_lo + per);
    constexpr int kSeg = kWqSegment;                                       // queue entries carry a 16-bit obstacle index: lane | link << 5 | obstacle << 10 | edge code << 26
    for (int base = j_lo; base < j_hi; base += kSeg) {
        const int count = min(kSeg, j_hi - base);
        seg_base = base;                                                   // (every interval ends with a flush, so no entry outlives its segment)
#pragma unroll 1
        for (int g0 = 0; g0 < L; g0 += G) {
            unsigned alive = (1u << min(G, L - g0)) - 1u;                 // links some configuration of the warp still has to test (warp-uniform)
            const float4 *o4 = obs4 + 2 * size_t(base);
            const float *sd = a.scan_dist + base;
#pragma unroll 1
            for (int j0 = 0; j0 < count; j0 += 64) {      // (an interval = 64 obstacles)
                // every 64 obstacles: drop configurations that are already invalid and links no later obstacle can reach; pack the
                // live links into slots 0..K-1 (K and the slot -> link map are warp-uniform)
                const bool dead = !live || invalid[t] != 0u;
                const R dnext = R(__ldg(sd + j0));
                unsigned mine = 0;
#pragma unroll
                for (int l = 0; l < G; l++) {
                    const int i = g0 + l;
                    if (!dead && i < L && !(dnext > cut[i * kWqStride + t])) mine |= 1u << l;
                }
                alive &= __reduce_or_sync(0xffffffffu, mine);
                // nothing left to test: one empty iteration over the tail of the table, so that the flush at the end of every
                // interval (the only inlined copy of the exact predicate) still runs for this group
                const int jfirst = alive ? j0 : (count - 1) & ~1;
                const int K = __popc(alive);
                unsigned slotmap = 0, ground_keep = 0;
                float lo[G][3], hi[G][3];                                 // FP32 in every variant (the table is FP32); F64 bounds rounded outwards
                {
                    unsigned rem = alive;
#pragma unroll
                    for (int s = 0; s < G; s++) {
                        const int l = __ffs(int(rem)) - 1, i = g0 + l;
                        rem &= rem - 1u;
                        if (s < K && ((mine >> l) & 1u)) {
                            const R *sk = skel + (3 * i) * kWqStride + t;
                            R mag = R(0);
#pragma unroll
                            for (int k = 0; k < 3; k++) { const R u = sk[k * kWqStride], v = sk[(3 + k) * kWqStride]; mag = fmax(mag, fmax(fabs(u), fabs(v))); }
                            const R thr = radius[i] * R(1.001) + R(1e-5) + R(4e-6) * mag;      // looser than the reference's `+- radius` slab
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const R u = sk[k * kWqStride], v = sk[(3 + k) * kWqStride];
                                lo[s][k] = wq_round_down(fmin(u, v) - thr);
                                hi[s][k] = wq_round_up(fmax(u, v) + thr);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 3; k++) { lo[s][k] = wq_inf<float>(); hi[s][k] = -wq_inf<float>(); }
                        }
                        if (s < K) {
                            slotmap |= unsigned(i) << (5 * s);
                            if (i >= ground_included) ground_keep |= 1u << s;                 // links below ground_included ignore ground objects (:285-286)
                        }
                    }
                    ground_keep *= (1u << G) + 1u;
                }
                const int jend = alive ? min(j0 + 64, count) : count;
#pragma unroll 1
                for (int j = jfirst; j < jend; j += kWqUnroll) {
                    const int j1 = j + 1 < jend ? j + 1 : j;
                    const float4 *pj = o4 + 2 * j, *pj1 = pj + (j1 - j) * 2;
                    const float4 w0 = __ldg(pj), w1 = __ldg(pj + 1), v0 = __ldg(pj1), v1 = __ldg(pj1 + 1);
                    unsigned mask = 0;
#define RVS_VQ_SLOT(S)                                                                                                             \
    if (S < G) {                                                                                                                   \
        wq_overlap<float>(mask, 1u << (S), w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, lo[S < G ? S : 0], hi[S < G ? S : 0]); \
        wq_overlap<float>(mask, 1u << (G + (S)), v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, lo[S < G ? S : 0], hi[S < G ? S : 0]); \
    }
                    switch (K) {
                        case 0: break;
                        default: RVS_VQ_SLOT(5) [[fallthrough]];
                        case 5: RVS_VQ_SLOT(4) [[fallthrough]];
                        case 4: RVS_VQ_SLOT(3) [[fallthrough]];
                        case 3: RVS_VQ_SLOT(2) [[fallthrough]];
                        case 2: RVS_VQ_SLOT(1) [[fallthrough]];
                        case 1: RVS_VQ_SLOT(0)
                    }
#undef RVS_VQ_SLOT
                    if (j1 == j) mask &= (1u << G) - 1u;
                    if ((__float_as_uint(w1.w) | __float_as_uint(v1.w)) & 1u)
                        mask &= ((__float_as_uint(w1.w) & 1u) ? (ground_keep | ~((1u << G) - 1u)) : ~0u) &
                                ((__float_as_uint(v1.w) & 1u) ? (ground_keep | ((1u << G) - 1u)) : ~0u);
                    if (mask) {
                        unsigned pos = atomicAdd(const_cast<unsigned *>(&tail[0]), unsigned(__popc(mask)));
                        do {
                            const int b = __ffs(int(mask)) - 1;
                            mask &= mask - 1u;
                            const int sl = b >= G ? b - G : b, jj = b >= G ? j1 : j;
                            q1[pos++] = unsigned(lane) | (((slotmap >> (5 * sl)) & 31u) << 5) | (unsigned(jj) << 10);
                        } while (mask);
                    }
                    __syncwarp();
                    // Every interval ends with a flush: a configuration whose only colliding pairs wait in a partly filled queue would
                    // otherwise stay "valid" and keep its links open (and the whole warp scanning) to the end of the table.
                    const bool last = j + kWqUnroll >= jend;
                    if (tail[0] >= 32u || last) drain(last);
                }
                if (!alive) break;
            }
        }
    }
    __syncwarp();
    if (!live) return;
    if (!EDGE && slices == 1) a.verdict[item] = invalid[t] ? 0 : 1;
    else if (invalid[t]) a.verdict[EDGE ? edge : item] = 0;               // verdict[] is pre-set to 1 by the host
}

}  // namespace rvs
