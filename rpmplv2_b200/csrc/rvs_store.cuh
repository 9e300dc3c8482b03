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
    float *slot = a.store + size_t(a.slots[cfg]) * store_slot_floats(L, M);
    float *planes = slot + store_header_floats(L);
    if (threadIdx.x == 0) {
        float sk[3 * (kMaxDofs + 1)];
        fk_skeleton<F, float>(rb, q, sk);
        for (int k = 0; k < 3 * (L + 1); k++) s_sk[k] = sk[k];
    }
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

// One warp per spine. gen2 = 0: extendGenSpine (stop when d_c < D_CRIT or status != Advanced, RGBTConnect.cpp:73-92);
// gen2 = 1: extendGenSpine2 (stop when d_c < D_CRIT or status == Reached, :96-118). Every layer's state is written to q_layers.
template <class F>
__global__ void __launch_bounds__(128) store_spine_kernel(const StoreArgs a) {
    const RobotDev *__restrict__ rb = a.robot;
    const int n = a.n, L = n, W = n + 1, M = a.M;
    const int s = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= a.S) return;
    const float *slot = a.store + size_t(a.root_slot[s]) * store_slot_floats(n, M);
    const float *root_planes = slot + store_header_floats(n);
    const float *q_e = a.q_targets + size_t(s) * n;
    float *q_out = a.q_layers + size_t(s) * a.num_layers * n;
    float q_cur[kMaxDofs], q_new[kMaxDofs], prof_cur[kMaxDofs], prof_new[kMaxDofs], delta[kMaxDofs], rho_profile[kMaxDofs];
    float sk[3 * (kMaxDofs + 1)], sk_new[3 * (kMaxDofs + 1)];
    float Rm[kMaxDofs * (kMaxDofs + 1)];
    for (int i = 0; i < n; i++) { q_cur[i] = slot[i]; q_new[i] = slot[i]; prof_cur[i] = slot[n + 1 + i]; }
    const float root_d_c = slot[n];
    float dc_cur = root_d_c, d_out = root_d_c;
    int status = 2, layers = 0;                                        // Trapped when the root itself has no clearance (RBTConnect.cpp:91-92)
    if (root_d_c > 0.f) {
        for (int layer = 0; layer < a.num_layers; layer++) {
            status = 1;
            fk_skeleton<F, float>(rb, q_cur, sk);
            float rho = 0.f;
            for (int i = 0; i < L; i++) rho_profile[i] = 0.f;
            float q_temp[kMaxDofs];
            for (int i = 0; i < n; i++) q_temp[i] = q_cur[i];
            for (int counter = 1;; counter++) {
                if (counter == 1) enclosing_radii_of<float>(rb, sk, Rm);
                else { fk_skeleton<F, float>(rb, q_temp, sk_new); enclosing_radii_of<float>(rb, sk_new, Rm); }
                for (int i = 0; i < n; i++) delta[i] = fabsf(q_e[i] - q_temp[i]);
                float step;
                if (a.use_expanded_bubble) {
                    step = __int_as_float(0x7f800000);
                    for (int i = 0; i < n; i++) {
                        const float v = (prof_cur[i] - rho_profile[i]) / dot_dyn(Rm + (i + 1), W, delta, n);
                        if (v < step) step = v;
                    }
                } else
                    step = (dc_cur - rho) / dot_dyn(Rm + n, W, delta, n);
                if (step > 1.f) {
                    for (int i = 0; i < n; i++) q_new[i] = q_e[i];
                    status = 3;
                } else
                    for (int i = 0; i < n; i++) q_new[i] = q_temp[i] + step * (q_e[i] - q_temp[i]);
                if (counter == a.num_iter_spine || status != 1) break;
                fk_skeleton<F, float>(rb, q_new, sk_new);
                float rho_k_prev = 0.f;
                for (int k = 1; k <= L; k++) {
                    const float rho_k = sqrtf(sqD(sk[3 * k] - sk_new[3 * k], sk[3 * k + 1] - sk_new[3 * k + 1], sk[3 * k + 2] - sk_new[3 * k + 2]));
                    rho_profile[k - 1] = rho_k_prev < rho_k ? rho_k : rho_k_prev;
                    rho_k_prev = rho_k;
                    rho = rho < rho_k ? rho_k : rho;
                }
                for (int i = 0; i < n; i++) q_temp[i] = q_new[i];
            }
            if (lane == 0) for (int i = 0; i < n; i++) q_out[size_t(layer) * n + i] = q_new[i];
            layers = layer + 1;
            fk_skeleton<F, float>(rb, q_new, sk_new);
            bool rejected;
            d_out = underestimation_warp(rb, L, M, sk_new, root_planes, a.max_dist, prof_new, rejected, lane);
            if (d_out < a.d_crit || (a.gen2 ? status == 3 : status != 1)) break;
            for (int i = 0; i < n; i++) { q_cur[i] = q_new[i]; prof_cur[i] = prof_new[i]; }
            dc_cur = d_out;
        }
    }
    if (lane == 0) {
        if (layers == 0) { for (int i = 0; i < n; i++) q_out[i] = q_new[i]; }
        a.n_layers[s] = layers;
        a.d_c_last[s] = d_out;
        a.status[s] = uint8_t(status);
    }
}

}  // namespace rvs
