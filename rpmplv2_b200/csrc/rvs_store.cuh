// rvs_store.cuh — device-resident State cache and the planner-shaped kernels that work on it.
//
// The reference keeps, on every base::State, what computeDistance found for it: d_c, d_c_profile and the
// per-(obstacle, link) nearest points (include/state_spaces/State.h:24-41), and the bur planners read them back for
// every spine they grow from that state (RBTConnect.cpp:85-150, RGBTConnect.cpp:73-118). Here that cache lives in HBM
// as fixed-size SLOTS, so that the 24*L*M bytes of nearest points never cross PCIe:
//
//     slot s = [ q : n ][ d_c : 1 ][ profile : L ][ pad to a multiple of 4 floats ][ planes : M x L x 6 ]
//
//   store_distance_kernel   RealVectorSpace::computeDistance (RealVectorSpace.cpp:318-392) of a FEW configurations:
//                           one CTA per configuration, one warp per link, one lane per obstacle — the opposite
//                           decomposition of the sweep kernels (one lane per configuration), because planner batches are
//                           tens to thousands of configurations and latency, not throughput, is what the planner waits on.
//   store_spine_kernel      RGBTConnect::extendGenSpine / extendGenSpine2 (RGBTConnect.cpp:73-118) for S spines whose
//                           roots are slots: one WARP per spine; the spine recurrence is replicated in the 32 lanes and
//                           the distance-to-planes of every layer (L*M planes) is split over the lanes.
// Both produce results bit-identical to the sequential kernels (min is order independent; every value that is stored is
// computed by exactly the same instruction sequence).
#pragma once
#include "rvs_kernels.cuh"

namespace rvs {

constexpr int kStoreWarps = 8;

template <class F>
__global__ void __launch_bounds__(32 * kStoreWarps) store_distance_kernel(const StoreArgs a) {
    __shared__ float s_sk[3 * (kMaxDofs + 1)];
    __shared__ float s_prof[kMaxDofs];
    __shared__ float s_hit_d[kMaxDofs];
    __shared__ int s_hit_j[kMaxDofs];
    const RobotDev *__restrict__ rb = a.robot;
    const int L = a.n, M = a.M;
    const int cfg = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *q = a.q + size_t(cfg) * L;
    if (a.layers_n) {
        // step of a connect chain: the configuration is the last layer the spine kernel of this step wrote (a.q = its q_layers)
        const int nl = a.layers_n[cfg];
        if (nl < 0) { if (threadIdx.x == 0) a.d_c[cfg] = -1.f; return; }                   // (the whole CTA leaves: nothing to do)
        q = a.q + (size_t(cfg) * a.num_layers + size_t(nl > 1 ? nl - 1 : 0)) * L;
    }
    float *slot = a.store + size_t(a.slots[cfg]) * store_slot_floats(L, M);
    float *planes = slot + store_header_floats(L);
    if (warp == 0) fk_skeleton_warp<F, float>(rb, q, s_sk);            // lane k evaluates sincos(q_k); every lane stores the same skeleton (as distance_cta_kernel)
    if (threadIdx.x < L) slot[threadIdx.x] = q[threadIdx.x];
    __syncthreads();
    const int ground_included = rb->ground_included;
    const float eq_thr = a.eq_thr, max_dist = a.max_dist;
    for (int i = warp; i < L; i += kStoreWarps) {
        Capsule<float> cp;
        cp.set({s_sk[3 * i], s_sk[3 * i + 1], s_sk[3 * i + 2]}, {s_sk[3 * i + 3], s_sk[3 * i + 4], s_sk[3 * i + 5]}, rb->radius[i], 0.f);
        float best = max_dist, hit_d = 0.f;
        int hit_j = 0x7fffffff;
        for (int j = lane; j < M; j += 32) {
            const ObstacleRec o = a.obstacles[j];
            float d;
            V3<float> P1 = {0, 0, 0}, P2 = {0, 0, 0};
            if ((o.flags & 1u) && i < ground_included) {               // ground rule (RealVectorSpace.cpp:336-341)
                d = max_dist;
                P2.z = -max_dist;
            } else
                d = capsule_box_distance<float, true>(cp, load_box<float>(o), eq_thr, max_dist, P1, P2);
            if (d > o.tol) d = max_dist;                               // min_dist_tol (:367-368)
            if (d <= 0.f) { if (j < hit_j) { hit_j = j; hit_d = d; } }
            else if (d < best) best = d;
            float2 *dst = reinterpret_cast<float2 *>(planes + (size_t(j) * L + i) * 6);
            dst[0] = make_float2(P1.x, P1.y); dst[1] = make_float2(P1.z, P2.x); dst[2] = make_float2(P2.y, P2.z);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, hit_j, off);
            const float od = __shfl_xor_sync(0xffffffffu, hit_d, off);
            if (ob < best) best = ob;
            if (oj < hit_j) { hit_j = oj; hit_d = od; }
        }
        if (lane == 0) { s_prof[i] = best; s_hit_j[i] = hit_j; s_hit_d[i] = hit_d; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int limit = L;                                                  // first colliding link (link-major early exit, :371-378)
        for (int i = 0; i < L; i++) if (s_hit_j[i] != 0x7fffffff) { limit = i; break; }
        float d_c = max_dist;
        for (int i = 0; i < L; i++) if (s_prof[i] < d_c) d_c = s_prof[i];
        if (limit < L) d_c = 0.f;
        slot[L] = d_c;
        a.d_c[cfg] = d_c;
        for (int i = 0; i < L; i++) {
            const float p = limit == L ? s_prof[i] : (i < limit ? s_prof[i] : (i == limit ? s_hit_d[i] : 0.f));
            slot[L + 1 + i] = p;
            if (a.profile) a.profile[size_t(cfg) * L + i] = p;
        }
        if (a.chain_state && a.status[cfg] != 1) a.chain_state[cfg] = 0;                    // the loop ends unless the spine Advanced
    }
}

// computeDistance for a SMALL batch through the host / device-pointer entry points (rvs_distance*): the same decomposition as
// store_distance_kernel — one CTA per configuration, one warp per link, one lane per obstacle — because what a planner waits for is
// the latency of a handful of configurations, not throughput: a single xArm6 query against 51 boxes with nearest points takes the
// one-thread-per-configuration kernel 256 us (306 pairs in sequence), this one ~45 us.  MODE as in distance_kernel; FP32 geometry.
// Results are bit-identical to distance_kernel (same per-pair routine; minima and the first-collision rule are order free).
template <class F, int MODE>
__global__ void __launch_bounds__(32 * kStoreWarps) distance_cta_kernel(const KArgs a) {
    __shared__ float s_sk[3 * (kMaxDofs + 1)];
    __shared__ float s_prof[kMaxDofs];
    __shared__ float s_hit_d[kMaxDofs];
    __shared__ int s_hit_j[kMaxDofs];
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n, M = a.M;
    const int64_t cfg = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) fk_skeleton_warp<F, float>(rb, a.q + cfg * L, s_sk);      // lane k evaluates sincos(q_k); every lane stores the same skeleton
    __syncthreads();
    const int ground_included = rb->ground_included;
    const float eq_thr = a.eq_thr, max_dist = a.max_dist;
    float *nearest = MODE == RVS_DIST_FULL ? static_cast<float *>(a.nearest) + size_t(cfg) * size_t(M) * size_t(L) * 6 : nullptr;
    for (int i = warp; i < L; i += kStoreWarps) {
        Capsule<float> cp;
        cp.set({s_sk[3 * i], s_sk[3 * i + 1], s_sk[3 * i + 2]}, {s_sk[3 * i + 3], s_sk[3 * i + 4], s_sk[3 * i + 5]}, rb->radius[i], 0.f);
        float best = max_dist, hit_d = 0.f;
        int hit_j = 0x7fffffff;
        for (int j = lane; j < M; j += 32) {
            const ObstacleRec o = a.obstacles[j];
            float d;
            V3<float> P1 = {0, 0, 0}, P2 = {0, 0, 0};
            if ((o.flags & 1u) && i < ground_included) {               // ground rule (RealVectorSpace.cpp:336-341)
                d = max_dist;
                P2.z = -max_dist;
            } else
                d = capsule_box_distance<float, MODE == RVS_DIST_FULL>(cp, load_box<float>(o), eq_thr, max_dist, P1, P2);
            if (d > o.tol) d = max_dist;                               // min_dist_tol (:367-368)
            if (d <= 0.f) { if (j < hit_j) { hit_j = j; hit_d = d; } }
            else if (d < best) best = d;
            if (MODE == RVS_DIST_FULL) {
                float2 *dst = reinterpret_cast<float2 *>(nearest + (size_t(j) * L + i) * 6);
                dst[0] = make_float2(P1.x, P1.y); dst[1] = make_float2(P1.z, P2.x); dst[2] = make_float2(P2.y, P2.z);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, hit_j, off);
            const float od = __shfl_xor_sync(0xffffffffu, hit_d, off);
            if (ob < best) best = ob;
            if (oj < hit_j) { hit_j = oj; hit_d = od; }
        }
        if (lane == 0) { s_prof[i] = best; s_hit_j[i] = hit_j; s_hit_d[i] = hit_d; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int limit = L;                                                  // first colliding link (link-major early exit, :371-378)
        for (int i = 0; i < L; i++) if (s_hit_j[i] != 0x7fffffff) { limit = i; break; }
        float d_c = max_dist;
        for (int i = 0; i < L; i++) if (s_prof[i] < d_c) d_c = s_prof[i];
        static_cast<float *>(a.d_c)[cfg] = limit < L ? 0.f : d_c;
        if (MODE >= RVS_DIST_PROFILE) {
            float *out = static_cast<float *>(a.profile) + cfg * L;
            for (int i = 0; i < L; i++) out[i] = limit == L ? s_prof[i] : (i < limit ? s_prof[i] : (i == limit ? s_hit_d[i] : 0.f));
        }
        if (a.pair_count) {                                             // pair queries the reference's loop would have made
            unsigned visited = 0;
            for (int i = 0; i < limit; i++) visited += unsigned(M - (i < ground_included ? a.num_ground : 0));
            if (limit < L) {
                unsigned upto = unsigned(s_hit_j[limit] + 1);
                if (limit < ground_included && a.num_ground > 0)
                    for (int j = 0; j <= s_hit_j[limit]; j++) upto -= (a.obstacles[j].flags & 1u);
                visited += upto;
            }
            atomicAdd(a.pair_count, (unsigned long long)visited);
        }
    }
}

// isValid(q) / isValid(q1,q2) for a SMALL batch: one CTA per configuration (EDGE: per interpolated configuration of an edge), one warp
// per link, one lane per obstacle (slab reject per lane, survivors through the reference predicate 32 at a time), first hit ends the CTA.  In the work-queue kernel 32
// configurations share one warp's queues, so a planner-sized batch waits for ~10 rounds of the compacted phases per interval
// (64 configurations against 1000 boxes: 147 us); here the latency is one pass over M / 32 obstacles per lane.  Verdicts are order free.
template <class F, bool EDGE>
__global__ void __launch_bounds__(32 * kStoreWarps) valid_cta_kernel(const KArgs a) {
    __shared__ float s_sk[3 * (kMaxDofs + 1)];
    __shared__ float s_q[kMaxDofs];
    __shared__ int s_list[kStoreWarps][64];                              // per warp: obstacles that passed the slab test (leftover < 32 + 32 new)
    __shared__ int s_hit;
    const RobotDev *__restrict__ rb = a.robot;
    const int L = rb->n, n = L, M = a.M;
    const int64_t item = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t edge = 0;
    if (threadIdx.x == 0) s_hit = 0;
    if (warp == 0) {
        const float *q = a.q + item * n;
        if (EDGE) {                                                       // RealVectorSpace.cpp:267-269: k = N .. 1, interpolateEdge(q1, q2, k / N * dist, dist)
            edge = item / a.num_checks;
            const int kidx = int(item - edge * a.num_checks);
            const float *q1p = a.q + edge * n, *q2p = a.q2 + edge * n;
            const float dist = coord_norm(q1p, q2p, n);
            const float step = float(a.num_checks - kidx) / float(a.num_checks) * dist;
            if (lane < n) s_q[lane] = dist > 0.f ? q1p[lane] + (q2p[lane] - q1p[lane]) * (step / dist) : q2p[lane];
            __syncwarp();
            q = s_q;
        }
        fk_skeleton_warp<F, float>(rb, q, s_sk);
    }
    if (EDGE) edge = item / a.num_checks;
    __syncthreads();
    const int ground_included = rb->ground_included;
    const float eq_thr = a.eq_thr, max_dist = a.max_dist;
    // Per link: every lane applies the predicate's own slab reject (CollisionAndDistance.cpp:14-17, the exit ~98 % of the pairs take)
    // to its obstacle; the survivors are ballot-compacted into a small list and go through the full predicate 32 at a time, so the
    // warp does not pay the face / edge tests once per 32 obstacles because ONE lane needs them.  (An endpoint inside the inflated
    // box — the predicate's first, accepting test — is never slab-rejected, so nothing that collides is filtered.)
    // the flag only ever goes 0 -> 1 and is polled: atomics keep the inter-warp traffic well defined (and racecheck quiet)
    auto stopped = [&]() { return __shfl_sync(0xffffffffu, lane == 0 ? atomicOr(&s_hit, 0) : 0, 0) != 0; };
    for (int i = warp; i < L; i += kStoreWarps) {
        Capsule<float> cp;
        cp.set({s_sk[3 * i], s_sk[3 * i + 1], s_sk[3 * i + 2]}, {s_sk[3 * i + 3], s_sk[3 * i + 4], s_sk[3 * i + 5]}, rb->radius[i], rb->radius_in[i]);
        const V3<float> &A = cp.A, &B = cp.B;
        const float radius = cp.radius;
        int have = 0;
        bool done = false;
        auto exact = [&](int count) {                                     // the top `count` list entries through the reference predicate
            bool hit = false;
            if (lane < count) hit = capsule_hits_box<float>(cp, load_box<float>(a.obstacles[s_list[warp][have + lane]]), eq_thr, max_dist);
            if (__any_sync(0xffffffffu, hit)) { if (lane == 0) atomicOr(&s_hit, 1); done = true; }
            __syncwarp();
        };
        for (int j0 = 0; j0 < M && !done; j0 += 32) {
            const int j = j0 + lane;
            bool pass = false;
            if (j < M) {
                const ObstacleRec o = a.obstacles[j];
                if (!((o.flags & 1u) && i < ground_included)) {
                    const Box<float> bx = load_box<float>(o);
                    pass = !((A.x < bx.lo.x - radius && B.x < bx.lo.x - radius) || (A.x > bx.hi.x + radius && B.x > bx.hi.x + radius) ||
                             (A.y < bx.lo.y - radius && B.y < bx.lo.y - radius) || (A.y > bx.hi.y + radius && B.y > bx.hi.y + radius) ||
                             (A.z < bx.lo.z - radius && B.z < bx.lo.z - radius) || (A.z > bx.hi.z + radius && B.z > bx.hi.z + radius));
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, pass);
            if (pass) s_list[warp][have + __popc(b & ((1u << lane) - 1u))] = j;
            have += __popc(b);
            __syncwarp();
            if (have >= 32) { have -= 32; exact(32); }
            if (!done && stopped()) done = true;                         // another link already collides
        }
        if (!done && have > 0) { const int count = have; have = 0; exact(count); }
        if (done || stopped()) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!EDGE) a.verdict[item] = s_hit ? 0 : 1;
        else if (s_hit) a.verdict[edge] = 0;                              // verdict[] is pre-set to 1 by the host
    }
}

// distance-to-planes (RealVectorSpace.cpp:402-446) with the L*M planes of one layer split over the lanes of a warp;
// every lane returns the same d_c / prof / rejected. Values are those of underestimation_of() (same per-plane arithmetic).
__device__ __forceinline__ float underestimation_warp(const RobotDev *__restrict__ rb, int L, int M, const float *sk, const float *__restrict__ planes,
                                                      float max_dist, float *prof, bool &rejected, int lane) {
    float d_c = max_dist;
    bool rej = false;
#pragma unroll 1
    for (int i = 0; i < L; i++) {
        const V3<float> A = {sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, B = {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]};
        const float radius = rb->radius[i];
        float best = max_dist;
#pragma unroll 1
        for (int j = lane; j < M; j += 32) {
            const float2 *src = reinterpret_cast<const float2 *>(planes + (size_t(j) * L + i) * 6);
            const float2 w0 = src[0], w1 = src[1], w2 = src[2];
            const V3<float> Rp = {w0.x, w0.y, w1.x}, Op = {w1.y, w2.x, w2.y};
            if (rsqrt_of(dotF(Op, Op)) == max_dist) continue;
            V3<float> nrm = sub(Rp, Op);
            const float z = dotF(nrm, nrm);
            if (z > 0.f) { const float len = rsqrt_of(z); nrm = {nrm.x / len, nrm.y / len, nrm.z / len}; }
            const V3<float> ao = sub(A, Op), bo = sub(B, Op);
            const float da = (ao.x * nrm.x + ao.y * nrm.y) + ao.z * nrm.z;
            const float db = (bo.x * nrm.x + bo.y * nrm.y) + bo.z * nrm.z;
            const float d = (db < da ? db : da) - radius;
            if (d < 0.f) rej = true;
            if (d < best) best = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            if (ob < best) best = ob;
        }
        prof[i] = best;
        if (best < d_c) d_c = best;
    }
    rejected = __any_sync(0xffffffffu, rej);
    return rejected ? 0.f : d_c;
}

// One warp per spine, LANE i OWNS JOINT / LINK i (n <= 32): q_i, the profile entry of link i, its rho and its bubble ratio live in
// that lane's registers; the skeletons, the enclosing-radii matrix (row i written by lane i) and |q_e - q| sit in shared memory.
// A 22-DOF layer is ~23 000 dependent instructions when one thread walks it (profiles/README.md: a lone warp issues one
// instruction every ~6 cycles, so that is 70 us per layer); spread over the lanes it is a few thousand.  Same arithmetic per
// element, in the same order, as spine_kernel / the oracle; the minima over links are order-free (NaN ratios are skipped as
// `std::min(step, v)` skips them).
// gen2 = 0: extendGenSpine (stop when d_c < D_CRIT or status != Advanced, RGBTConnect.cpp:73-92); gen2 = 1: extendGenSpine2 (stop
// when d_c < D_CRIT or status == Reached, :96-118). Every layer's state is written to q_layers.
constexpr int kSpineWarps = 4;
__host__ __device__ inline size_t store_spine_smem_bytes(int n) {
    return size_t(kSpineWarps) * (size_t(2 * 3 * (n + 1)) + size_t(n) * (n + 1) + size_t(n)) * sizeof(float);
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const float o = __shfl_xor_sync(0xffffffffu, v, off); if (o < v) v = o; }
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const float o = __shfl_xor_sync(0xffffffffu, v, off); if (v < o) v = o; }
    return v;
}

// distance-to-planes of link `lane` (RealVectorSpace.cpp:402-446): lanes over links when there are at least as many links as
// obstacles, otherwise lanes over obstacles link by link.  Returns this lane's profile entry; d_c / rejected are warp-uniform.
__device__ __forceinline__ float underestimation_lanes(const RobotDev *__restrict__ rb, int L, int M, const float *sk, const float *__restrict__ planes,
                                                       float max_dist, float &d_c, bool &rejected, int lane) {
    auto plane_distance = [&](int i, int j, bool &skip) {
        const float2 *src = reinterpret_cast<const float2 *>(planes + (size_t(j) * L + i) * 6);
        const float2 w0 = src[0], w1 = src[1], w2 = src[2];
        const V3<float> Rp = {w0.x, w0.y, w1.x}, Op = {w1.y, w2.x, w2.y};
        skip = rsqrt_of(dotF(Op, Op)) == max_dist;
        V3<float> nrm = sub(Rp, Op);
        const float z = dotF(nrm, nrm);
        if (z > 0.f) { const float len = rsqrt_of(z); nrm = {nrm.x / len, nrm.y / len, nrm.z / len}; }
        const V3<float> A = {sk[3 * i], sk[3 * i + 1], sk[3 * i + 2]}, B = {sk[3 * i + 3], sk[3 * i + 4], sk[3 * i + 5]};
        const V3<float> ao = sub(A, Op), bo = sub(B, Op);
        const float da = (ao.x * nrm.x + ao.y * nrm.y) + ao.z * nrm.z;
        const float db = (bo.x * nrm.x + bo.y * nrm.y) + bo.z * nrm.z;
        return (db < da ? db : da) - rb->radius[i];
    };
    float mine = max_dist;
    bool rej = false;
    if (L >= M) {
        if (lane < L) {
#pragma unroll 1
            for (int j = 0; j < M; j++) {
                bool skip;
                const float d = plane_distance(lane, j, skip);
                if (skip) continue;
                if (d < 0.f) rej = true;
                if (d < mine) mine = d;
            }
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < L; i++) {
            float best = max_dist;
#pragma unroll 1
            for (int j = lane; j < M; j += 32) {
                bool skip;
                const float d = plane_distance(i, j, skip);
                if (skip) continue;
                if (d < 0.f) rej = true;
                if (d < best) best = d;
            }
            best = warp_min_f(best);
            if (lane == i) mine = best;
        }
    }
    rejected = __any_sync(0xffffffffu, rej);
    d_c = rejected ? 0.f : warp_min_f(lane < L ? mine : max_dist);
    return mine;
}

template <class F>
__global__ void __launch_bounds__(32 * kSpineWarps) store_spine_kernel(const StoreArgs a) {
    extern __shared__ __align__(16) float spine_smem[];
    const RobotDev *__restrict__ rb = a.robot;
    const int n = a.n, L = n, W = n + 1, M = a.M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kSpineWarps + warp;
    if (s >= a.S) return;                                              // (whole warps leave together; no block-level barrier below)
    float *sk = spine_smem + size_t(warp) * (2 * 3 * W + n * W + n), *sk_new = sk + 3 * W, *Rm = sk_new + 3 * W, *delta = Rm + n * W;
    const float *slot = a.store + size_t(a.root_slot[s]) * store_slot_floats(n, M);
    if (a.chain_state) {
        // step of a connect chain (connectGenSpine's loop, RGBTConnect.cpp:120-150): it runs while the chain is alive, the caller left
        // room for its result (a.slots[s] >= 0) and the root still has clearance above D_CRIT (below it the caller marches in RRT mode)
        const int alive = a.chain_state[s];
        if (!alive || a.slots[s] < 0 || !(slot[n] > a.d_crit)) {
            __syncwarp();
            if (lane == 0) { a.n_layers[s] = -1; a.status[s] = 0; a.d_c_last[s] = slot[n]; a.chain_state[s] = 0; }
            return;
        }
    }
    const float *root_planes = slot + store_header_floats(n);
    float *q_out = a.q_layers + size_t(s) * a.num_layers * n;
    const bool own = lane < n;
    const float q_e = own ? a.q_targets[size_t(s) * n + lane] : 0.f;
    float q_cur = own ? slot[lane] : 0.f, q_new = q_cur, prof_cur = own ? slot[n + 1 + lane] : 0.f;
    const float root_d_c = slot[n];
    float dc_cur = root_d_c, d_out = root_d_c;
    int status = 2, layers = 0;                                        // Trapped when the root itself has no clearance (RBTConnect.cpp:91-92)
    const float inf = __int_as_float(0x7f800000);
    // FK of the state whose joint i sits in lane i's register
    auto fk_of = [&](float q_i, float *out) {
        F s_mine = 0, c_mine = 1;
        if (own) sincos_of(F(q_i), &s_mine, &c_mine);
        fk_skeleton_with<F, float>(rb, [=](int k, F *sn, F *cs) { *sn = shfl_f(s_mine, k); *cs = shfl_f(c_mine, k); }, out);
        __syncwarp();
    };
    if (root_d_c > 0.f) {
        for (int layer = 0; layer < a.num_layers; layer++) {
            status = 1;
            if (layer == 0) fk_of(q_cur, sk);                          // later layers start where the previous one ended: its skeleton is kept
            float rho = 0.f, rho_profile = 0.f, q_temp = q_cur;
            for (int counter = 1;; counter++) {
                const float *sk_use = sk;
                if (counter > 1) { fk_of(q_temp, sk_new); sk_use = sk_new; }
                if (own) { enclosing_radii_row<float>(rb, sk_use, lane, Rm + lane * W); delta[lane] = fabsf(q_e - q_temp); }
                __syncwarp();
                float step;
                if (a.use_expanded_bubble) {
                    float v = inf;
                    if (own) { v = (prof_cur - rho_profile) / dot_dyn(Rm + (lane + 1), W, delta, n); if (!(v == v)) v = inf; }
                    step = warp_min_f(v);
                } else {
                    float dot = 0.f;
                    if (lane == n - 1) dot = dot_dyn(Rm + n, W, delta, n);
                    step = (dc_cur - rho) / __shfl_sync(0xffffffffu, dot, n - 1);
                }
                if (step > 1.f) { q_new = q_e; status = 3; }
                else q_new = q_temp + step * (q_e - q_temp);
                // RBTConnect.cpp:129-131: self-collision on [q_temp, q_new] (xArm6 only, n == 6).  The first bubble iteration of
                // xArm6::checkSelfCollision needs exactly the skeleton and enclosing radii of q_temp that are already in shared
                // memory: lane 0 evaluates its four bubble steps, and `step > 1` (no self-collision possible, by far the common
                // answer) ends the check there.  Otherwise lane 0 walks the whole routine (bubble steps + EPS_STEP marching).
                bool self_collision = false;
                if (rb->self_collision && rb->kind == RVS_ROBOT_XARM6 && n == 6) {
                    __syncwarp();
                    if (own) delta[lane] = fabsf(q_new - q_temp);
                    __syncwarp();
                    float q1v[6], q2v[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) { q1v[k] = __shfl_sync(0xffffffffu, q_temp, k); q2v[k] = __shfl_sync(0xffffffffu, q_new, k); }
                    // the four bubble steps spread over eight lanes; the rare path (`step > 1` fails): the whole routine, by the whole warp
                    // (cooperative FK in its bubble iterations, marching points 32 at a time)
                    float s04, s05, s14, s15;
                    const bool need = !(self_collision_steps_warp(rb, sk_use, Rm, delta, a.eq_thr, a.max_dist, s04, s05, s14, s15) > 1.f);
                    const int hit = need && self_collision_edge_warp<F>(rb, q1v, q2v, a.eq_thr, a.max_dist) ? 1 : 0;
                    if (hit) {
                        self_collision = true;
                        float mine = q_new;
#pragma unroll
                        for (int k = 0; k < 6; k++) { const float v = __shfl_sync(0xffffffffu, q2v[k], 0); if (lane == k) mine = v; }
                        q_new = mine;
                        float moved = 0.f;
                        if (lane == 0) moved = coord_norm(q1v, q2v, 6);
                        moved = __shfl_sync(0xffffffffu, moved, 0);
                        status = !(moved < a.eq_thr) ? 1 : 2;
                    }
                }
                if (counter == a.num_iter_spine || status != 1 || self_collision) break;
                fk_of(q_new, sk_new);
                // rho_k = |sk_k - sk_new_k| for k = 1..L (lane k-1); rho_profile[k-1] = max(rho_{k-1}, rho_k); rho = max(rho, rho_k)
                float rho_k = 0.f;
                if (own) {
                    const int k = lane + 1;
                    rho_k = sqrtf(sqD(sk[3 * k] - sk_new[3 * k], sk[3 * k + 1] - sk_new[3 * k + 1], sk[3 * k + 2] - sk_new[3 * k + 2]));
                }
                float rho_prev = __shfl_up_sync(0xffffffffu, rho_k, 1);
                if (lane == 0) rho_prev = 0.f;
                rho_profile = rho_prev < rho_k ? rho_k : rho_prev;
                const float m = warp_max_f(rho_k);
                rho = rho < m ? m : rho;
                q_temp = q_new;
                __syncwarp();
            }
            if (own) q_out[size_t(layer) * n + lane] = q_new;
            layers = layer + 1;
            __syncwarp();
            fk_of(q_new, sk_new);
            bool rejected;
            const float prof_new = underestimation_lanes(rb, L, M, sk_new, root_planes, a.max_dist, d_out, rejected, lane);
            if (d_out < a.d_crit || (a.gen2 ? status == 3 : status != 1)) break;
            q_cur = q_new; prof_cur = prof_new; dc_cur = d_out;
            for (int k = lane; k < 3 * W; k += 32) sk[k] = sk_new[k];  // FK(q_cur) of the next layer == FK(q_new) of this one
            __syncwarp();
        }
    }
    if (layers == 0 && own) q_out[lane] = q_new;
    if (lane == 0) {
        a.n_layers[s] = layers;
        a.d_c_last[s] = d_out;
        a.status[s] = uint8_t(status);
    }
}

}  // namespace rvs
