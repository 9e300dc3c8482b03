// rvs_capi.cu — C ABI of librvs_cuda.so (include/rvs_cuda.h): handle, HBM-resident robot + obstacle tables,
// staging buffers and kernel dispatch. Host code only; the kernels live in rvs_variant.cu.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "rvs_launch.h"
#include "rvs_robot_host.h"
#include "rvs_sort.cuh"
#include "rvs_types.h"

using namespace rvs;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;        // where the current call finds this buffer: the device allocation, or a slice of the pinned arena
    void *dev = nullptr;      // grow-only device allocation
    size_t cap = 0;
};
struct PendingCopy { void *dst; const void *src; size_t bytes; };
constexpr size_t kArenaBytes = 1 << 20;      // pinned, device-mapped staging for small calls
constexpr size_t kSmallCall = 192 << 10;     // in + out bytes up to which a host-pointer call takes the zero-copy path

}  // namespace

struct rvs_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_k2 = nullptr;   // copy streams + second compute stream of the pipelined host-pointer distance path (lazy)
    cudaEvent_t ev_in[32] = {}, ev_k[32] = {}, ev_done = nullptr;
    rvs_params prm{};
    RobotDev h_robot{};
    RobotDev *d_robot = nullptr;
    void *d_obs_pack = nullptr;               // fp16 boxes of the scan-order table, two obstacles per 32-byte record
    // coarse distance field of the table (first thresholds of the work-queue distance kernel): built lazily, rebuilt when the table changed
    void *d_df = nullptr;
    bool df_dirty = true;
    int df_on = 1;
    // permutation that makes the warps of a large distance batch geometrically coherent (rvs_sort.cuh)
    int sort_on = 1;
    int sort_per_bin = 128;                      // configurations per bin aimed at (RVS_SORT_PER_BIN overrides)
    int sort_joints = 3;                         // leading joints in the sort key (RVS_SORT_JOINTS overrides, for experiments)
    int32_t *d_perm = nullptr;
    size_t perm_cap = 0;
    int32_t *d_list = nullptr;               // FULL mode of a large batch: collision-free configurations (rvs_distance_full.cuh)
    size_t list_cap = 0;
    bool full_split_on = true;               // RVS_FULL_SPLIT=0: distance_kernel<FULL> as before (A/B)
    uint32_t *d_hist[2] = {nullptr, nullptr};    // one per compute stream of the pipelined host path
    float df_origin[3] = {0.f, 0.f, 0.f}, df_cell = 0.f;
    ObstacleRec *d_obs = nullptr;
    int M = 0, obs_cap = 0;
    std::vector<ObstacleRec> h_obs, h_obs_scan;
    std::vector<float> h_scan_dist;
    LaunchTable lt{};
    DevBuf buf[8];            // grow-only device staging for the host-pointer entry points
    int64_t launches = 0;
    int cull = 1;
    int num_ground = 0;
    unsigned long long *d_pair_count = nullptr;
    // dynamic environment (rvs_env_*): obstacle states + the timeline of the last advance, all resident
    EnvRec *d_env = nullptr;      // [env_cap] current states, then [env_tl_cap] timeline records, allocated separately
    EnvRec *d_env_tl = nullptr;
    float *d_scan_key = nullptr;
    int32_t *d_first_invalid = nullptr;
    LibcRand *d_rand = nullptr;   // [1 + rand_tl_cap]: the libc stream of random_directions, then its per-update timeline
    size_t rand_tl_cap = 0;
    int32_t env_ground_included = 0;
    int env_cap = 0;
    size_t env_tl_cap = 0, env_fi_cap = 0;
    bool env_on = false;
    rvs_env_params env_prm{};
    float *d_store = nullptr;     // device-resident State cache (rvs_store_*): store_cap slots of store_slot_floats(n, M) floats
    int store_cap = 0;
    // Small host-pointer calls (planner-shaped: one to a few hundred configurations) skip the three DMA transfers: inputs are
    // memcpy'd into a pinned, device-mapped arena, the kernel reads and writes that arena over PCIe directly, and the results are
    // memcpy'd out after one stream synchronisation (a single query drops from ~35-60 us to the launch + sync latency).
    char *arena = nullptr;
    size_t arena_off = 0;
    bool zc = false;
    std::vector<PendingCopy> pending;
    std::string err;
    // device-resident planner trees (rvs_tree_*): states of tree t at trees[t].d [cap][n]; the pointer / count tables the kernels read
    // are mirrored in d_tree_ptr / d_tree_cnt and refreshed lazily
    struct DevTree { float *d = nullptr; int count = 0, cap = 0; bool released = false; };
    std::vector<DevTree> trees;
    const float **d_tree_ptr = nullptr;
    int32_t *d_tree_cnt = nullptr;
    int tree_tab_cap = 0;
    bool tree_tab_dirty = true;
};

namespace {

enum { B_Q = 0, B_Q2, B_DC, B_PROF, B_NEAR, B_VERD, B_PLANES, B_IDX };

int fail(rvs_engine *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(rvs_engine *h, cudaError_t e, const char *what) {
    const int code = (e == cudaErrorMemoryAllocation) ? RVS_ERR_OUT_OF_MEMORY : RVS_ERR_CUDA;
    return fail(h, code, std::string(what) + ": " + cudaGetErrorString(e));
}
#define RVS_CUDA(h, call)                                                   \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return cuda_fail((h), e__, #call);          \
    } while (0)
#define RVS_TRY(expr)                       \
    do {                                    \
        const int rc__ = (expr);            \
        if (rc__ != RVS_OK) return rc__;    \
    } while (0)

int reserve(rvs_engine *h, int which, size_t bytes) {
    DevBuf &b = h->buf[which];
    if (h->zc) {
        const size_t off = (h->arena_off + 15) & ~size_t(15);
        if (off + bytes > kArenaBytes) return fail(h, RVS_ERR_OUT_OF_MEMORY, "internal: small-call arena overflow");
        b.p = h->arena + off;
        h->arena_off = off + bytes;
        return RVS_OK;
    }
    if (bytes > b.cap) {
        if (b.dev) RVS_CUDA(h, cudaFree(b.dev));
        b.dev = nullptr; b.cap = 0;
        const size_t cap = bytes + bytes / 4 + 256;
        RVS_CUDA(h, cudaMalloc(&b.dev, cap));
        b.cap = cap;
    }
    b.p = b.dev;
    return RVS_OK;
}

// every host-pointer entry point brackets its work with begin_call / finish
void begin_call(rvs_engine *h, size_t in_plus_out_bytes) {
    h->zc = h->arena != nullptr && in_plus_out_bytes <= kSmallCall;
    h->arena_off = 0;
    h->pending.clear();
}
int finish(rvs_engine *h) {
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    const bool zc = h->zc;
    h->zc = false;
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaStreamSynchronize");
    if (zc) for (const PendingCopy &c : h->pending) std::memcpy(c.dst, c.src, c.bytes);
    h->pending.clear();
    return RVS_OK;
}

int edges_dev_impl(rvs_engine *h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_valid);
int underestimation_dev_impl(rvs_engine *h, const float *d_q, int64_t N, const float *d_planes, const int32_t *d_plane_idx, int32_t R, float *d_d_c,
                             float *d_profile, uint8_t *d_status);

int build_robot(rvs_engine *h, const rvs_robot_desc *rd) {
    const char *why = fill_robot_dev(rd, &h->h_robot);
    return why ? fail(h, RVS_ERR_INVALID_ARGUMENT, why) : RVS_OK;
}

int select_table(rvs_engine *h) {
    switch (h->prm.precision) {
        case RVS_PREC_REF32: h->lt = launch_table_ref32(); return RVS_OK;
        case RVS_PREC_FAST32: h->lt = launch_table_fast32(); return RVS_OK;
        case RVS_PREC_F64: h->lt = launch_table_f64(); return RVS_OK;
        default: return fail(h, RVS_ERR_INVALID_ARGUMENT, "unknown precision");
    }
}

int check_params(rvs_engine *h, const rvs_params *p) {
    if (!p) return fail(h, RVS_ERR_INVALID_ARGUMENT, "params is NULL");
    if (!(p->max_distance > 0) || p->num_interpolation_validity_checks < 1 || !(p->equality_threshold >= 0))
        return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad RealVectorSpaceConfig values");
    return RVS_OK;
}

KArgs base_args(rvs_engine *h, const float *d_q, int64_t N) {
    KArgs a{};
    a.obstacles_pack = h->d_obs_pack;
    a.robot = h->d_robot; a.obstacles = h->d_obs; a.obstacles_scan = h->d_obs + h->obs_cap; a.scan_dist = reinterpret_cast<const float *>(h->d_obs + 2 * size_t(h->obs_cap)); a.M = h->M; a.N = N; a.q = d_q;
    a.eq_thr = h->prm.equality_threshold; a.max_dist = h->prm.max_distance;
    a.num_checks = h->prm.num_interpolation_validity_checks; a.cull = h->cull;
    a.robot_n_hint = h->h_robot.n;
    a.num_ground = h->num_ground;
    a.pair_count = h->d_pair_count;
    a.slices = 1;
    a.df = nullptr;                             // switched on by distance_dev_impl for the work-queue kernel
    return a;
}


// validity work-queue kernel: how many slices of the obstacle table per batch of `items` configurations (1 = every warp walks the
// whole table).  Worth it when the batch alone cannot fill the GPU and the table is long.
int valid_slices(const rvs_engine *h, int64_t items) {
#ifdef RVS_NO_SLICES
    return 1;
#endif
    if (!h->cull || h->M < 256 || items >= 32768) return 1;
    return std::min(8, h->M / 128);
}

int enter(rvs_engine *h) {
    if (!h) return fail(nullptr, RVS_ERR_INVALID_ARGUMENT, "handle is NULL");
    h->zc = false;                            // a call that failed half-way must not leave the zero-copy mode on
    RVS_CUDA(h, cudaSetDevice(h->device));
    return RVS_OK;
}

int launched(rvs_engine *h, cudaError_t e, const char *what) {
    if (e != cudaSuccess) return cuda_fail(h, e, what);
    h->launches++;
    return RVS_OK;
}

// Distance field for the work-queue distance kernel: 128^3 cells (64^3 / 96^3 / 192^3 measured: 1.566 / 1.534 / 1.506 ms per 1e6 against 1.521) centred on the robot base, wide enough for every point of the chain
// (sum of the link offsets + the gripper) plus a margin.  (Re)built on the handle's stream when the table changed since the last build.
#ifndef RVS_DF_DIM
#define RVS_DF_DIM 128
#endif
constexpr int kDfDim = RVS_DF_DIM;
constexpr int64_t kDfMinConfigs = 32768;        // below this the build (~0.1 ms) is not worth it
constexpr int kDfMinObstacles = 32;
int df_ready(rvs_engine *h) {
    if (!h->df_dirty && h->d_df) return RVS_OK;
    if (!h->d_df) RVS_CUDA(h, cudaMalloc(&h->d_df, size_t(kDfDim) * kDfDim * kDfDim * 2));
    double reach = double(h->h_robot.gripper_length) + 0.3;
    for (int i = 1; i <= h->h_robot.n; i++)
        reach += std::sqrt(h->h_robot.xyz[i][0] * h->h_robot.xyz[i][0] + h->h_robot.xyz[i][1] * h->h_robot.xyz[i][1] + h->h_robot.xyz[i][2] * h->h_robot.xyz[i][2]);
    h->df_cell = float(2.0 * reach / kDfDim);
    for (int c = 0; c < 3; c++) h->df_origin[c] = float(h->h_robot.xyz[0][c] - reach);
    RVS_TRY(launched(h, launch_df_build(h->d_obs, h->M, h->prm.max_distance, h->d_df, kDfDim, h->df_origin, h->df_cell, h->stream), "distance field build"));
    h->df_dirty = false;
    return RVS_OK;
}
void df_args(rvs_engine *h, KArgs &a) {
    a.df = h->d_df; a.df_dim = kDfDim;
    for (int c = 0; c < 3; c++) a.df_origin[c] = h->df_origin[c];
    a.df_inv_cell = 1.0f / h->df_cell;
    a.df_slack = h->df_cell * 0.8660255f * 1.0001f + 1e-6f;                 // half a cell diagonal
}

// Coherent warps for the work-queue distance kernel: counting sort of the batch d_q[0, N) on its leading joints, written to
// d_perm[perm_offset, perm_offset + N) (as indices relative to d_q) on stream `st`.
constexpr int64_t kSortMinConfigs = 65536;
constexpr int kSortMaxBits = 21;
int sort_ready(rvs_engine *h, size_t total) {
    if (total > h->perm_cap) {
        if (h->d_perm) RVS_CUDA(h, cudaFree(h->d_perm));
        h->d_perm = nullptr; h->perm_cap = 0;
        RVS_CUDA(h, cudaMalloc(&h->d_perm, total * sizeof(int32_t)));
        h->perm_cap = total;
    }
    for (auto &p : h->d_hist)
        if (!p) RVS_CUDA(h, cudaMalloc(&p, (size_t(1) << kSortMaxBits) * sizeof(uint32_t)));
    return RVS_OK;
}
constexpr int kFullSplitMinObstacles = 16;      // shorter tables: the PROFILE pass in front costs more than the idle lanes it removes (131072 x 6: 0.21 -> 0.27 ms)
int list_ready(rvs_engine *h, size_t N) {
    if (N + 4 <= h->list_cap) return RVS_OK;
    if (h->d_list) RVS_CUDA(h, cudaFree(h->d_list));
    h->d_list = nullptr; h->list_cap = 0;
    RVS_CUDA(h, cudaMalloc(&h->d_list, (N + 4) * sizeof(int32_t)));
    h->list_cap = N + 4;
    return RVS_OK;
}
int sort_configs(rvs_engine *h, const float *d_q, int64_t N, int64_t perm_offset, int which, cudaStream_t st) {
    SortArgs a{};
    a.q = d_q; a.N = N; a.n = h->h_robot.n; a.joints = std::min(h->sort_joints, h->h_robot.n);
    int total = 0;
    while (total < kSortMaxBits && (int64_t(h->sort_per_bin) << (total + 1)) <= N) total++;   // that many configurations per bin, or more
    for (int j = 0; j < 6; j++) a.bits[j] = 0;
    for (int b = 0; b < total; b++) a.bits[b % a.joints]++;
    for (int j = 0; j < a.joints; j++) {
        a.lo[j] = h->h_robot.limit_lo[j];
        const float span = h->h_robot.limit_hi[j] - h->h_robot.limit_lo[j];
        a.inv[j] = span > 0.f ? 1.0f / span : 0.f;
    }
    a.nbins = 1 << total; a.hist = h->d_hist[which]; a.perm = h->d_perm + perm_offset;
    RVS_CUDA(h, cudaMemsetAsync(a.hist, 0, size_t(a.nbins) * sizeof(uint32_t), st));
    RVS_TRY(launched(h, launch_sort_configs(a, st), "configuration sort"));
    h->launches += 2;                                                                     // three grids
    return RVS_OK;
}

// large validity batches take the work-queue kernel too: same permutation (edges are ordered by their first end)
int sort_for_validity(rvs_engine *h, KArgs &a, int64_t N) {
    if (!h->sort_on || !h->cull || h->M <= 0 || h->M >= (1 << 22) || N < kSortMinConfigs || N >= (int64_t(1) << 31)) return RVS_OK;
    RVS_TRY(sort_ready(h, size_t(N)));
    RVS_TRY(sort_configs(h, a.q, N, 0, 0, h->stream));
    a.perm = h->d_perm;
    return RVS_OK;
}

int h2d(rvs_engine *h, int which, const void *src, size_t bytes) {
    RVS_TRY(reserve(h, which, bytes));
    if (!bytes) return RVS_OK;
    if (h->zc) std::memcpy(h->buf[which].p, src, bytes);
    else RVS_CUDA(h, cudaMemcpyAsync(h->buf[which].p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return RVS_OK;
}
int d2h_from(rvs_engine *h, void *dst, const void *src, size_t bytes) {
    if (!bytes || !dst) return RVS_OK;
    if (h->zc) h->pending.push_back({dst, src, bytes});
    else RVS_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return RVS_OK;
}
int d2h(rvs_engine *h, void *dst, int which, size_t bytes) { return d2h_from(h, dst, h->buf[which].p, bytes); }

int distance_dev_impl(rvs_engine *h, const float *d_q, int64_t N, int32_t mode, void *d_dc, void *d_prof, void *d_near, bool want_f64,
                      cudaStream_t on = nullptr, int64_t perm_offset = 0, int which = 0) {
    if ((h->prm.precision == RVS_PREC_F64) != want_f64)
        return fail(h, RVS_ERR_INVALID_ARGUMENT, want_f64 ? "handle was not created with RVS_PREC_F64" : "RVS_PREC_F64 handles use the *_f64 entry points");
    if (N < 0 || mode < RVS_DIST_MIN || mode > RVS_DIST_FULL) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad N or mode");
    if (N == 0) return RVS_OK;
    if (!d_q || !d_dc || (mode >= RVS_DIST_PROFILE && !d_prof) || (mode == RVS_DIST_FULL && !d_near && h->M > 0))
        return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer for the requested distance mode");
    KArgs a = base_args(h, d_q, N);
    a.mode = mode; a.d_c = d_dc; a.profile = d_prof; a.nearest = d_near;
    // FULL for a large batch: PROFILE through the work-queue kernel + the nearest points of the collision-free configurations (rvs_distance_full.cuh)
    const bool split = mode == RVS_DIST_FULL && h->full_split_on && h->cull && h->M >= kFullSplitMinObstacles && N > distance_small_batch(mode, h->M) && N < (int64_t(1) << 31);
    if (split) {
        RVS_TRY(list_ready(h, size_t(N)));
        a.list = h->d_list;
        h->launches += 2;
    }
    if (h->df_on && h->cull && (mode != RVS_DIST_FULL || split) && h->M >= kDfMinObstacles && h->M < (1 << 16) && (N >= kDfMinConfigs || on)) {
        if (!on) RVS_TRY(df_ready(h));           // (the pipelined host path builds it before it forks its streams)
        if (h->d_df && !h->df_dirty) df_args(h, a);
    }
    if (h->sort_on && h->cull && (mode != RVS_DIST_FULL || split) && h->M > 0 && N >= kSortMinConfigs && N < (int64_t(1) << 31)) {      // (distance_kernel<FULL>: measured slower with the permutation, 2.10 -> 2.91 ms for 131072 x 51: its scattered result rows lose their locality)
        if (!on) RVS_TRY(sort_ready(h, size_t(N)));   // (the pipelined host path sizes the buffers before it forks its streams)
        RVS_TRY(sort_configs(h, d_q, N, perm_offset, which, on ? on : h->stream));
        a.perm = h->d_perm + perm_offset;
    }
    return launched(h, h->lt.distance(a, on ? on : h->stream), "distance kernel launch");
}

int pipeline_ready(rvs_engine *h) {
    if (h->s_in) return RVS_OK;
    RVS_CUDA(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    RVS_CUDA(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    RVS_CUDA(h, cudaStreamCreateWithFlags(&h->s_k2, cudaStreamNonBlocking));
    for (int k = 0; k < 32; k++) {
        RVS_CUDA(h, cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming));
        RVS_CUDA(h, cudaEventCreateWithFlags(&h->ev_k[k], cudaEventDisableTiming));
    }
    RVS_CUDA(h, cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    return RVS_OK;
}

// Host-pointer computeDistance. Large MIN/PROFILE batches are cut into up to 8 chunks and pipelined over three streams
// (H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c; consecutive kernels alternate between two compute
// streams so that the last partial wave of one overlaps the first wave of the next); pinned host buffers give full overlap,
// pageable ones still work. Small batches and FULL mode (whose output dwarfs everything else) take the single-shot path.
int distance_host_impl(rvs_engine *h, const float *q, int64_t N, int32_t mode, void *d_c, void *profile, void *nearest, bool want_f64) {
    RVS_TRY(enter(h));
    if (N < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "N < 0");
    if (N == 0) return RVS_OK;
    if (!q || !d_c) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL q or d_c");
    const size_t rs = want_f64 ? sizeof(double) : sizeof(float);
    const size_t n = size_t(h->h_robot.n), L = n;
    const size_t near_bytes = mode == RVS_DIST_FULL ? size_t(N) * size_t(h->M) * L * 6 * rs : 0;
    begin_call(h, size_t(N) * (n * sizeof(float) + rs * (1 + L)) + near_bytes + 64);
    RVS_TRY(reserve(h, B_Q, size_t(N) * n * sizeof(float)));
    RVS_TRY(reserve(h, B_DC, size_t(N) * rs));
    RVS_TRY(reserve(h, B_PROF, size_t(N) * L * rs));
    RVS_TRY(reserve(h, B_NEAR, near_bytes));
    float *dq = static_cast<float *>(h->buf[B_Q].p);
    char *ddc = static_cast<char *>(h->buf[B_DC].p), *dprof = static_cast<char *>(h->buf[B_PROF].p);
    constexpr int64_t kMinChunk = 1 << 16;
    static const int max_chunks = [] { const char *e = std::getenv("RVS_PIPE_CHUNKS"); return e ? std::max(1, std::min(32, std::atoi(e))) : 12; }();
    const int chunks = mode == RVS_DIST_FULL ? 1 : int(std::min<int64_t>(max_chunks, N / kMinChunk));
    if (chunks < 2) {
        if (h->zc) std::memcpy(dq, q, size_t(N) * n * sizeof(float));
        else RVS_CUDA(h, cudaMemcpyAsync(dq, q, size_t(N) * n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        RVS_TRY(distance_dev_impl(h, dq, N, mode, ddc, dprof, h->buf[B_NEAR].p, want_f64));
        RVS_TRY(d2h(h, d_c, B_DC, size_t(N) * rs));
        if (mode >= RVS_DIST_PROFILE) RVS_TRY(d2h(h, profile, B_PROF, size_t(N) * L * rs));
        if (mode == RVS_DIST_FULL) RVS_TRY(d2h(h, nearest, B_NEAR, near_bytes));
        return finish(h);
    }
    if (mode >= RVS_DIST_PROFILE && !profile) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL profile buffer");
    RVS_TRY(pipeline_ready(h));
    if (h->df_on && h->cull && h->M >= kDfMinObstacles && h->M < (1 << 16)) RVS_TRY(df_ready(h));
    if (h->sort_on && h->cull && h->M > 0) RVS_TRY(sort_ready(h, size_t(N)));
    RVS_CUDA(h, cudaEventRecord(h->ev_done, h->stream));                 // order the copy streams after earlier work on the handle
    RVS_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_done, 0));
    RVS_CUDA(h, cudaStreamWaitEvent(h->s_k2, h->ev_done, 0));
    // Chunk sizes ramp up and down (weights 1, 2, 4, 4, ..., 4, 2, 1): the pipeline's fill (first H2D) and drain (last D2H) are the part of
    // the transfers that compute cannot hide, so the first and last chunks are small.  RVS_PIPE_RAMP=0: equal chunks.
    static const bool ramp = [] { const char *e = std::getenv("RVS_PIPE_RAMP"); return !e || std::atoi(e) != 0; }();
    int64_t bounds[33];
    {
        double wsum = 0, acc = 0;
        auto weight = [&](int c) { const int edge = std::min(c, chunks - 1 - c); return (!ramp || chunks < 6) ? 4.0 : (edge == 0 ? 1.0 : (edge == 1 ? 2.0 : 4.0)); };
        for (int c = 0; c < chunks; c++) wsum += weight(c);
        bounds[0] = 0;
        for (int c = 0; c < chunks; c++) {
            acc += weight(c);
            bounds[c + 1] = c + 1 == chunks ? N : std::min<int64_t>(N, (int64_t(double(N) * acc / wsum) + 127) & ~int64_t(127));
        }
    }
    for (int c = 0; c < chunks; c++) {
        const int64_t first = bounds[c], cnt = bounds[c + 1] - first;
        if (cnt <= 0) continue;
        RVS_CUDA(h, cudaMemcpyAsync(dq + first * n, q + first * n, size_t(cnt) * n * sizeof(float), cudaMemcpyHostToDevice, h->s_in));
        RVS_CUDA(h, cudaEventRecord(h->ev_in[c], h->s_in));
        cudaStream_t sk = (c & 1) ? h->s_k2 : h->stream;
        RVS_CUDA(h, cudaStreamWaitEvent(sk, h->ev_in[c], 0));
        RVS_TRY(distance_dev_impl(h, dq + first * n, cnt, mode, ddc + first * rs, dprof + first * L * rs, nullptr, want_f64, sk, first, c & 1));
        RVS_CUDA(h, cudaEventRecord(h->ev_k[c], sk));
        RVS_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_k[c], 0));
        RVS_CUDA(h, cudaMemcpyAsync(static_cast<char *>(d_c) + first * rs, ddc + first * rs, size_t(cnt) * rs, cudaMemcpyDeviceToHost, h->s_out));
        if (mode >= RVS_DIST_PROFILE)
            RVS_CUDA(h, cudaMemcpyAsync(static_cast<char *>(profile) + first * L * rs, dprof + first * L * rs, size_t(cnt) * L * rs, cudaMemcpyDeviceToHost, h->s_out));
    }
    RVS_CUDA(h, cudaStreamSynchronize(h->s_out));
    RVS_CUDA(h, cudaStreamSynchronize(h->s_k2));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    return RVS_OK;
}

// simple device-to-host pattern for kernels with one output array living in buffer `which`
int run_simple(rvs_engine *h, const float *q, int64_t N, LaunchFn fn, void *out, size_t out_bytes, bool verdict_out, const char *what) {
    RVS_TRY(enter(h));
    if (N < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "N < 0");
    if (N == 0) return RVS_OK;
    if (!q || !out) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    begin_call(h, size_t(N) * size_t(h->h_robot.n) * sizeof(float) + out_bytes + 64);
    RVS_TRY(h2d(h, B_Q, q, size_t(N) * size_t(h->h_robot.n) * sizeof(float)));
    const int which = verdict_out ? B_VERD : B_DC;
    RVS_TRY(reserve(h, which, out_bytes));
    KArgs a = base_args(h, static_cast<const float *>(h->buf[B_Q].p), N);
    if (verdict_out) a.verdict = static_cast<uint8_t *>(h->buf[which].p); else a.d_c = h->buf[which].p;
    if (verdict_out && fn == h->lt.is_valid) {
        a.slices = valid_slices(h, N);
        if (a.slices > 1) {                                                // the slices only ever clear a verdict
            if (h->zc) std::memset(h->buf[which].p, 1, out_bytes);
            else RVS_CUDA(h, cudaMemsetAsync(h->buf[which].p, 1, out_bytes, h->stream));
        }
    }
    if (verdict_out && fn == h->lt.is_valid) RVS_TRY(sort_for_validity(h, a, N));
    RVS_TRY(launched(h, fn(a, h->stream), what));
    RVS_TRY(d2h(h, out, which, out_bytes));
    return finish(h);
}

}  // namespace

extern "C" {

int rvs_abi_version(void) { return RVS_ABI_VERSION; }

int rvs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *rvs_last_error(rvs_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rvs_create(const rvs_robot_desc *robot, const rvs_params *params, int device, rvs_handle *out) {
    if (!out) return fail(nullptr, RVS_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!robot) return fail(nullptr, RVS_ERR_INVALID_ARGUMENT, "robot is NULL");
    RVS_TRY(check_params(nullptr, params));
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, RVS_ERR_NO_DEVICE, "no CUDA device available: this engine has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, RVS_ERR_INVALID_ARGUMENT, "device index out of range");
    rvs_engine *h = new rvs_engine();
    if (const char *e_so = std::getenv("RVS_SORT_CONFIGS")) h->sort_on = std::atoi(e_so) ? 1 : 0;
    if (const char *e_fs = std::getenv("RVS_FULL_SPLIT")) h->full_split_on = std::atoi(e_fs) != 0;
    if (const char *e_sb = std::getenv("RVS_SORT_PER_BIN")) h->sort_per_bin = std::max(1, std::atoi(e_sb));
    if (const char *e_sj = std::getenv("RVS_SORT_JOINTS")) h->sort_joints = std::max(1, std::min(6, std::atoi(e_sj)));
    if (const char *e_df = std::getenv("RVS_DISTANCE_FIELD")) h->df_on = std::atoi(e_df) ? 1 : 0;      // A/B switch (results are identical either way)
    h->device = device;
    h->prm = *params;
    int rc = build_robot(h, robot);
    if (rc == RVS_OK) rc = select_table(h);
    if (rc != RVS_OK) { g_create_error = h->err; delete h; return rc; }
    auto bail = [&](cudaError_t ce, const char *what) {
        const int code = cuda_fail(nullptr, ce, what);
        if (h->d_robot) cudaFree(h->d_robot);
        if (h->d_pair_count) cudaFree(h->d_pair_count);
        if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
        delete h;
        return code;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    h->own_stream = true;
    if ((e = cudaMalloc(&h->d_robot, sizeof(RobotDev))) != cudaSuccess) return bail(e, "cudaMalloc(robot)");
    if ((e = cudaMemcpy(h->d_robot, &h->h_robot, sizeof(RobotDev), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy(robot)");
    if ((e = cudaMalloc(&h->d_pair_count, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc(pair counter)");
    if ((e = cudaMemset(h->d_pair_count, 0, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMemset(pair counter)");
    void *arena = nullptr;
    if (cudaHostAlloc(&arena, kArenaBytes, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) h->arena = static_cast<char *>(arena);
    else cudaGetLastError();                 // no pinned memory to spare: small calls simply take the DMA path
    *out = h;
    return RVS_OK;
}

int rvs_destroy(rvs_handle h) {
    if (!h) return RVS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &b : h->buf) if (b.dev) cudaFree(b.dev);
    if (h->arena) cudaFreeHost(h->arena);
    if (h->d_obs) cudaFree(h->d_obs);
    if (h->d_obs_pack) cudaFree(h->d_obs_pack);
    if (h->d_df) cudaFree(h->d_df);
    if (h->d_perm) cudaFree(h->d_perm);
    if (h->d_list) cudaFree(h->d_list);
    for (auto *p : h->d_hist) if (p) cudaFree(p);
    for (auto &t : h->trees) if (t.d) cudaFree(t.d);
    if (h->d_tree_ptr) cudaFree(h->d_tree_ptr);
    if (h->d_tree_cnt) cudaFree(h->d_tree_cnt);
    if (h->d_store) cudaFree(h->d_store);
    if (h->d_env) cudaFree(h->d_env);
    if (h->d_env_tl) cudaFree(h->d_env_tl);
    if (h->d_scan_key) cudaFree(h->d_scan_key);
    if (h->d_first_invalid) cudaFree(h->d_first_invalid);
    if (h->d_rand) cudaFree(h->d_rand);
    if (h->d_robot) cudaFree(h->d_robot);
    if (h->d_pair_count) cudaFree(h->d_pair_count);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return RVS_OK;
}

int rvs_set_params(rvs_handle h, const rvs_params *params) {
    RVS_TRY(enter(h));
    RVS_TRY(check_params(h, params));
    const int32_t keep = h->prm.precision;
    h->prm = *params;
    if (params->precision < 0) h->prm.precision = keep;      // "leave the arithmetic variant as it is"
    return select_table(h);
}

int rvs_get_params(rvs_handle h, rvs_params *out) {
    if (!h || !out) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = h->prm;
    return RVS_OK;
}

int rvs_num_dofs(rvs_handle h) { return h ? h->h_robot.n : 0; }

int rvs_get_limits(rvs_handle h, float *limit_lo, float *limit_hi) {
    if (!h || !limit_lo || !limit_hi) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL argument");
    for (int i = 0; i < h->h_robot.n; i++) { limit_lo[i] = h->h_robot.limit_lo[i]; limit_hi[i] = h->h_robot.limit_hi[i]; }
    return RVS_OK;
}

int rvs_set_stream(rvs_handle h, void *cuda_stream) {
    RVS_TRY(enter(h));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) { RVS_CUDA(h, cudaStreamDestroy(h->stream)); h->own_stream = false; }
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return RVS_OK;
}

void *rvs_get_stream(rvs_handle h) { return h ? static_cast<void *>(h->stream) : nullptr; }

int rvs_sync(rvs_handle h) {
    RVS_TRY(enter(h));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    return RVS_OK;
}

int64_t rvs_launch_count(rvs_handle h) { return h ? h->launches : 0; }

int rvs_set_culling(rvs_handle h, int32_t cull) {
    if (!h) return RVS_ERR_INVALID_ARGUMENT;
    h->cull = cull ? 1 : 0;
    return RVS_OK;
}

int rvs_num_obstacles(rvs_handle h) { return h ? h->M : 0; }

int rvs_measure_fma_peak(rvs_handle h, int32_t fp64, double *tflops) {
    RVS_TRY(enter(h));
    if (!tflops) return fail(h, RVS_ERR_INVALID_ARGUMENT, "tflops is NULL");
    cudaDeviceProp prop;
    RVS_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 8192;
    RVS_TRY(reserve(h, B_DC, 64));
    cudaEvent_t e0, e1;
    RVS_CUDA(h, cudaEventCreate(&e0));
    RVS_CUDA(h, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {        // first repetition is the warm-up
        RVS_CUDA(h, cudaEventRecord(e0, h->stream));
        RVS_TRY(launched(h, launch_fma_peak(fp64, h->buf[B_DC].p, blocks, iters, h->stream), "fma peak kernel launch"));
        RVS_CUDA(h, cudaEventRecord(e1, h->stream));
        RVS_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0;
        RVS_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = double(blocks) * 256.0 * 16.0 * double(iters) * 2.0;
        if (rep > 0 && ms > 0) best = std::fmax(best, flops / (double(ms) * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return RVS_OK;
}

int rvs_pair_count(rvs_handle h, int32_t reset, uint64_t *out) {
    RVS_TRY(enter(h));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    unsigned long long v = 0;
    RVS_CUDA(h, cudaMemcpy(&v, h->d_pair_count, sizeof(v), cudaMemcpyDeviceToHost));
    if (out) *out = uint64_t(v);
    if (reset) RVS_CUDA(h, cudaMemset(h->d_pair_count, 0, 8 * sizeof(v)));
    return RVS_OK;
}

int rvs_work_stats(rvs_handle h, int32_t reset, uint64_t *out4) {
    RVS_TRY(enter(h));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    unsigned long long v[4] = {0, 0, 0, 0};
    RVS_CUDA(h, cudaMemcpy(v, h->d_pair_count, sizeof(v), cudaMemcpyDeviceToHost));
    if (out4) for (int k = 0; k < 4; k++) out4[k] = uint64_t(v[k]);
    if (reset) RVS_CUDA(h, cudaMemset(h->d_pair_count, 0, 8 * sizeof(unsigned long long)));
    return RVS_OK;
}

int rvs_set_distance_field(rvs_handle h, int32_t on) {
    if (!h) return RVS_ERR_INVALID_ARGUMENT;
    h->df_on = on ? 1 : 0;
    return RVS_OK;
}
int rvs_set_config_sort(rvs_handle h, int32_t on) {
    if (!h) return RVS_ERR_INVALID_ARGUMENT;
    h->sort_on = on ? 1 : 0;
    return RVS_OK;
}
int rvs_second_pass_warps(rvs_handle h, int32_t reset, uint64_t *out) {
    RVS_TRY(enter(h));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    unsigned long long v = 0;
    RVS_CUDA(h, cudaMemcpy(&v, h->d_pair_count + 4, sizeof(v), cudaMemcpyDeviceToHost));
    if (out) *out = uint64_t(v);
    if (reset) RVS_CUDA(h, cudaMemset(h->d_pair_count + 4, 0, sizeof(v)));
    return RVS_OK;
}

// Uploads the table twice: [0, M) in the caller's order (what every kernel indexes results by) and [cap, cap + M) in SCAN ORDER for
// the work-queue distance kernel: nearest to the robot base first, so that the per-link minima — and with them the broad-phase
// thresholds — tighten after a few obstacles. Scan-order records carry their original index in flags bits 8..31.
namespace {
int upload_obstacles(rvs_engine *h) {
    const int M = h->M;
    if (!M) return RVS_OK;
    const double bx = h->h_robot.xyz[0][0], by = h->h_robot.xyz[0][1], bz = h->h_robot.xyz[0][2];
    std::vector<std::pair<double, int>> order;
    order.resize(size_t(M));
    for (int j = 0; j < M; j++) {
        const ObstacleRec &o = h->h_obs[size_t(j)];
        const double gx = std::max(std::max(double(o.lo[0]) - bx, bx - double(o.hi[0])), 0.0), gy = std::max(std::max(double(o.lo[1]) - by, by - double(o.hi[1])), 0.0),
                     gz = std::max(std::max(double(o.lo[2]) - bz, bz - double(o.hi[2])), 0.0);
        const double d2 = gx * gx + gy * gy + gz * gz;
        order[size_t(j)] = {d2 == d2 ? d2 : 1e300, j};
    }
    std::stable_sort(order.begin(), order.end());
    h->h_obs_scan.resize(size_t(M));
    h->h_scan_dist.resize(size_t(M));
    for (int k = 0; k < M; k++) {
        // lower bound of the obstacle's distance from the base in float (the kernel's cut-off must never overestimate it)
        const double d = std::sqrt(std::min(order[size_t(k)].first, 1e60));
        float f = float(d);
        if (double(f) > d) f = std::nextafterf(f, 0.0f);
        h->h_scan_dist[size_t(k)] = f;
    }
    for (int k = 0; k < M; k++) {
        ObstacleRec r = h->h_obs[size_t(order[size_t(k)].second)];
        r.flags = (r.flags & 1u) | (uint32_t(order[size_t(k)].second) << 8);
        h->h_obs_scan[size_t(k)] = r;
    }
    RVS_CUDA(h, cudaMemcpyAsync(h->d_obs, h->h_obs.data(), size_t(M) * sizeof(ObstacleRec), cudaMemcpyHostToDevice, h->stream));
    RVS_CUDA(h, cudaMemcpyAsync(h->d_obs + h->obs_cap, h->h_obs_scan.data(), size_t(M) * sizeof(ObstacleRec), cudaMemcpyHostToDevice, h->stream));
    RVS_CUDA(h, cudaMemcpyAsync(h->d_obs + 2 * size_t(h->obs_cap), h->h_scan_dist.data(), size_t(M) * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RVS_CUDA(h, launch_pack_scan(h->d_obs + h->obs_cap, M, h->d_obs_pack, h->stream));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    h->df_dirty = true;
    return RVS_OK;
}
}  // namespace

int rvs_set_obstacles(rvs_handle h, const float *aabb_min, const float *aabb_max, const uint8_t *is_ground, const float *min_dist_tol, int32_t M) {
    RVS_TRY(enter(h));
    if (M < 0 || M >= (1 << 24) || (M > 0 && (!aabb_min || !aabb_max))) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad obstacle arrays (0 <= M < 2^24)");
    h->h_obs.resize(size_t(M));
    for (int j = 0; j < M; j++) {
        ObstacleRec &o = h->h_obs[size_t(j)];
        for (int c = 0; c < 3; c++) { o.lo[c] = aabb_min[3 * j + c]; o.hi[c] = aabb_max[3 * j + c]; }
        o.tol = min_dist_tol ? min_dist_tol[j] : h->prm.max_distance;          // Box.cpp:13
        o.flags = (is_ground && is_ground[j]) ? 1u : 0u;
    }
    h->num_ground = 0;
    for (int j = 0; j < M; j++) h->num_ground += int(h->h_obs[size_t(j)].flags & 1u);
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (M > h->obs_cap) {
        if (h->d_obs) RVS_CUDA(h, cudaFree(h->d_obs));
        h->d_obs = nullptr; h->obs_cap = 0;
        const int cap = M + M / 2 + 16;
        RVS_CUDA(h, cudaMalloc(&h->d_obs, 2 * size_t(cap) * sizeof(ObstacleRec) + size_t(cap) * sizeof(float)));   // caller order, scan order, scan distances
        if (h->d_obs_pack) RVS_CUDA(h, cudaFree(h->d_obs_pack));
        h->d_obs_pack = nullptr;
        RVS_CUDA(h, cudaMalloc(&h->d_obs_pack, (size_t(cap) + 2) * 16));                                             // fp16 copy of the scan-order table: 32 B per obstacle pair
        h->obs_cap = cap;
    }
    if (M != h->M && h->d_store) {          // the slot stride depends on M: cached states of another table are meaningless
        RVS_CUDA(h, cudaFree(h->d_store));
        h->d_store = nullptr; h->store_cap = 0;
    }
    h->M = M;
    h->env_on = false;                       // a plain table: rvs_env_set switches the dynamic environment on again
    return upload_obstacles(h);
}

int rvs_update_obstacle_aabbs(rvs_handle h, const float *aabb_min, const float *aabb_max, int32_t M) {
    RVS_TRY(enter(h));
    if (h->env_on) return fail(h, RVS_ERR_INVALID_ARGUMENT, "the obstacle table is driven by rvs_env_*: move the objects with rvs_env_step / rvs_env_set");
    if (M != h->M) return fail(h, RVS_ERR_INVALID_ARGUMENT, "M differs from the resident obstacle table");
    if (M > 0 && (!aabb_min || !aabb_max)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL AABB arrays");
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));       // the pinned-less host mirror is reused: wait for the last copy
    for (int j = 0; j < M; j++)
        for (int c = 0; c < 3; c++) { h->h_obs[size_t(j)].lo[c] = aabb_min[3 * j + c]; h->h_obs[size_t(j)].hi[c] = aabb_max[3 * j + c]; }
    return upload_obstacles(h);
}

// ---- device-pointer entry points ---------------------------------------------------------------------------------
int rvs_distance_dev(rvs_handle h, const float *d_q, int64_t N, int32_t mode, float *d_d_c, float *d_profile, float *d_nearest) {
    RVS_TRY(enter(h));
    return distance_dev_impl(h, d_q, N, mode, d_d_c, d_profile, d_nearest, false);
}
int rvs_distance_f64_dev(rvs_handle h, const float *d_q, int64_t N, int32_t mode, double *d_d_c, double *d_profile, double *d_nearest) {
    RVS_TRY(enter(h));
    return distance_dev_impl(h, d_q, N, mode, d_d_c, d_profile, d_nearest, true);
}
int rvs_is_valid_dev(rvs_handle h, const float *d_q, int64_t N, uint8_t *d_valid) {
    RVS_TRY(enter(h));
    if (N < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "N < 0");
    if (N == 0) return RVS_OK;
    if (!d_q || !d_valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    KArgs a = base_args(h, d_q, N);
    a.verdict = d_valid;
    a.slices = valid_slices(h, N);
    if (a.slices > 1) {                                                    // the slices only ever clear a verdict
        if (h->zc) std::memset(d_valid, 1, size_t(N));
        else RVS_CUDA(h, cudaMemsetAsync(d_valid, 1, size_t(N), h->stream));
    }
    RVS_TRY(sort_for_validity(h, a, N));
    return launched(h, h->lt.is_valid(a, h->stream), "is_valid kernel launch");
}
int rvs_is_valid_edges_dev(rvs_handle h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_valid) {
    RVS_TRY(enter(h));
    return edges_dev_impl(h, d_q1, d_q2, E, d_valid);
}
}  // extern "C"
namespace {
// (no enter(): the host wrappers call these after begin_call(), and enter() would switch the zero-copy mode off again)
int edges_dev_impl(rvs_engine *h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_valid) {
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!d_q1 || !d_q2 || !d_valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (h->zc) std::memset(d_valid, 1, size_t(E));                       // small host call: the verdicts live in the pinned arena
    else RVS_CUDA(h, cudaMemsetAsync(d_valid, 1, size_t(E), h->stream));
    KArgs a = base_args(h, d_q1, E);
    a.q2 = d_q2; a.verdict = d_valid;
    a.slices = valid_slices(h, E * int64_t(h->prm.num_interpolation_validity_checks));
    RVS_TRY(sort_for_validity(h, a, E));
    return launched(h, h->lt.edge_valid(a, h->stream), "edge kernel launch");
}
int underestimation_dev_impl(rvs_engine *h, const float *d_q, int64_t N, const float *d_planes, const int32_t *d_plane_idx, int32_t R,
                             float *d_d_c, float *d_profile, uint8_t *d_status) {
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "underestimation has no FP64 entry point");
    if (N < 0 || R < 1) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad N or R");
    if (N == 0) return RVS_OK;
    if (!d_q || !d_d_c || (!d_planes && h->M > 0)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    KArgs a = base_args(h, d_q, N);
    a.planes = d_planes; a.plane_idx = d_plane_idx; a.R = R; a.d_c = d_d_c; a.profile = d_profile; a.status = d_status;
    return launched(h, h->lt.underestimation(a, h->stream), "underestimation kernel launch");
}
}  // namespace
extern "C" {
int rvs_underestimation_dev(rvs_handle h, const float *d_q, int64_t N, const float *d_planes, const int32_t *d_plane_idx, int32_t R,
                            float *d_d_c, float *d_profile, uint8_t *d_status) {
    RVS_TRY(enter(h));
    return underestimation_dev_impl(h, d_q, N, d_planes, d_plane_idx, R, d_d_c, d_profile, d_status);
}
int rvs_generate_configs_dev(rvs_handle h, uint64_t seed, int64_t first, int64_t N, float *d_q) {
    RVS_TRY(enter(h));
    if (N < 0 || first < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad range");
    if (N == 0) return RVS_OK;
    if (!d_q) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    KArgs a = base_args(h, nullptr, N);
    a.seed = seed; a.first = first;
    return launched(h, launch_generate_configs(a, d_q, h->stream), "generate kernel launch");
}

// ---- host-pointer entry points -------------------------------------------------------------------------------------
int rvs_distance(rvs_handle h, const float *q, int64_t N, int32_t mode, float *d_c, float *profile, float *nearest) {
    return distance_host_impl(h, q, N, mode, d_c, profile, nearest, false);
}
int rvs_distance_f64(rvs_handle h, const float *q, int64_t N, int32_t mode, double *d_c, double *profile, double *nearest) {
    return distance_host_impl(h, q, N, mode, d_c, profile, nearest, true);
}
int rvs_is_valid(rvs_handle h, const float *q, int64_t N, uint8_t *valid) {
    if (h && h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_INVALID_ARGUMENT, "RVS_PREC_F64 handles use rvs_is_valid_f64");
    return run_simple(h, q, N, h ? h->lt.is_valid : nullptr, valid, size_t(N > 0 ? N : 0), true, "is_valid kernel launch");
}
int rvs_is_valid_f64(rvs_handle h, const float *q, int64_t N, uint8_t *valid) {
    if (h && h->prm.precision != RVS_PREC_F64) return fail(h, RVS_ERR_INVALID_ARGUMENT, "handle was not created with RVS_PREC_F64");
    return run_simple(h, q, N, h ? h->lt.is_valid : nullptr, valid, size_t(N > 0 ? N : 0), true, "is_valid kernel launch");
}
int rvs_is_valid_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *valid) {
    RVS_TRY(enter(h));
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!q1 || !q2 || !valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    const size_t bytes = size_t(E) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, 2 * bytes + size_t(E) + 64);
    RVS_TRY(h2d(h, B_Q, q1, bytes));
    RVS_TRY(h2d(h, B_Q2, q2, bytes));
    RVS_TRY(reserve(h, B_VERD, size_t(E)));
    RVS_TRY(edges_dev_impl(h, static_cast<const float *>(h->buf[B_Q].p), static_cast<const float *>(h->buf[B_Q2].p), E,
                           static_cast<uint8_t *>(h->buf[B_VERD].p)));
    RVS_TRY(d2h(h, valid, B_VERD, size_t(E)));
    return finish(h);
}
int rvs_underestimation(rvs_handle h, const float *q, int64_t N, const float *planes, const int32_t *plane_idx, int32_t R,
                        float *d_c, float *profile, uint8_t *status) {
    RVS_TRY(enter(h));
    if (N < 0 || R < 1) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad N or R");
    if (N == 0) return RVS_OK;
    if (!q || !d_c || (!planes && h->M > 0)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (plane_idx)
        for (int64_t i = 0; i < N; i++)
            if (plane_idx[i] < 0 || plane_idx[i] >= R) return fail(h, RVS_ERR_INVALID_ARGUMENT, "plane_idx out of range");
    const size_t n = size_t(h->h_robot.n);
    begin_call(h, size_t(N) * (n * 8 + 16) + size_t(R) * size_t(h->M) * n * 6 * sizeof(float) + 128);
    RVS_TRY(h2d(h, B_Q, q, size_t(N) * n * sizeof(float)));
    RVS_TRY(h2d(h, B_PLANES, planes, size_t(R) * size_t(h->M) * n * 6 * sizeof(float)));
    if (plane_idx) RVS_TRY(h2d(h, B_IDX, plane_idx, size_t(N) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_DC, size_t(N) * sizeof(float)));
    RVS_TRY(reserve(h, B_PROF, size_t(N) * n * sizeof(float)));
    RVS_TRY(reserve(h, B_VERD, size_t(N)));
    RVS_TRY(underestimation_dev_impl(h, static_cast<const float *>(h->buf[B_Q].p), N, static_cast<const float *>(h->buf[B_PLANES].p),
                                    plane_idx ? static_cast<const int32_t *>(h->buf[B_IDX].p) : nullptr, R,
                                    static_cast<float *>(h->buf[B_DC].p), static_cast<float *>(h->buf[B_PROF].p),
                                    static_cast<uint8_t *>(h->buf[B_VERD].p)));
    RVS_TRY(d2h(h, d_c, B_DC, size_t(N) * sizeof(float)));
    RVS_TRY(d2h(h, profile, B_PROF, size_t(N) * n * sizeof(float)));
    RVS_TRY(d2h(h, status, B_VERD, size_t(N)));
    return finish(h);
}
int rvs_skeleton(rvs_handle h, const float *q, int64_t N, float *skeleton) {
    if (h && h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "skeleton output is FP32: use a REF32/FAST32 handle");
    const size_t n = h ? size_t(h->h_robot.n) : 0;
    return run_simple(h, q, N, h ? h->lt.skeleton : nullptr, skeleton, size_t(N > 0 ? N : 0) * (n + 1) * 3 * sizeof(float), false, "skeleton kernel launch");
}
int rvs_enclosing_radii(rvs_handle h, const float *q, int64_t N, float *radii) {
    if (h && h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "enclosing radii output is FP32: use a REF32/FAST32 handle");
    const size_t n = h ? size_t(h->h_robot.n) : 0;
    return run_simple(h, q, N, h ? h->lt.enclosing_radii : nullptr, radii, size_t(N > 0 ? N : 0) * n * (n + 1) * sizeof(float), false, "enclosing radii kernel launch");
}
int rvs_self_collision_capsules(rvs_handle h, const float *q, int64_t N, uint8_t *mask) {
    return run_simple(h, q, N, h ? h->lt.self_collision : nullptr, mask, size_t(N > 0 ? N : 0), true, "self collision kernel launch");
}
int rvs_set_self_collision(rvs_handle h, int32_t enabled, float eps_step) {
    RVS_TRY(enter(h));
    if (!(eps_step > 0.f)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "eps_step must be > 0");
    h->h_robot.self_collision = enabled ? 1 : 0;
    h->h_robot.eps_step = eps_step;
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));                        // no kernel of this handle may still be reading the old flags
    RVS_CUDA(h, cudaMemcpy(h->d_robot, &h->h_robot, sizeof(RobotDev), cudaMemcpyHostToDevice));
    return RVS_OK;
}
int rvs_get_self_collision(rvs_handle h, int32_t *enabled, float *eps_step) {
    if (!h) return fail(nullptr, RVS_ERR_INVALID_ARGUMENT, "handle is NULL");
    if (enabled) *enabled = h->h_robot.self_collision;
    if (eps_step) *eps_step = h->h_robot.eps_step;
    return RVS_OK;
}
// ss->isValid(q) && !ss->robot->checkSelfCollision(q)  (RRTConnect.cpp:13-15, MotionValidity.cpp:42, RGBMTStar.cpp:202)
int rvs_check_states(rvs_handle h, const float *q, int64_t N, uint8_t *valid) {
    RVS_TRY(enter(h));
    if (N < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "N < 0");
    if (N == 0) return RVS_OK;
    if (!q || !valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    const size_t bytes = size_t(N) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, bytes + size_t(N) + 64);
    RVS_TRY(h2d(h, B_Q, q, bytes));
    RVS_TRY(reserve(h, B_VERD, size_t(N)));
    KArgs a = base_args(h, static_cast<const float *>(h->buf[B_Q].p), N);
    a.verdict = static_cast<uint8_t *>(h->buf[B_VERD].p);
    a.slices = valid_slices(h, N);
    if (a.slices > 1) {
        if (h->zc) std::memset(a.verdict, 1, size_t(N));
        else RVS_CUDA(h, cudaMemsetAsync(a.verdict, 1, size_t(N), h->stream));
    }
    RVS_TRY(launched(h, h->lt.is_valid(a, h->stream), "is_valid kernel launch"));
    if (h->h_robot.self_collision && h->h_robot.kind == RVS_ROBOT_XARM6 && h->prm.precision != RVS_PREC_F64) {
        a.mode = 1; a.slices = 0;
        RVS_TRY(launched(h, h->lt.self_collision(a, h->stream), "self collision kernel launch"));
    }
    RVS_TRY(d2h(h, valid, B_VERD, size_t(N)));
    return finish(h);
}
// ss->isValid(q1, q2) && !ss->robot->checkSelfCollision(q1, q2)  (RRTConnect.cpp:100, RRTx.cpp:467-468, 507-508)
int rvs_check_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *valid) {
    RVS_TRY(enter(h));
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!q1 || !q2 || !valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    const size_t bytes = size_t(E) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, 2 * bytes + size_t(E) + 64);
    RVS_TRY(h2d(h, B_Q, q1, bytes));
    RVS_TRY(h2d(h, B_Q2, q2, bytes));
    RVS_TRY(reserve(h, B_VERD, size_t(E)));
    const float *d_q1 = static_cast<const float *>(h->buf[B_Q].p), *d_q2 = static_cast<const float *>(h->buf[B_Q2].p);
    uint8_t *d_valid = static_cast<uint8_t *>(h->buf[B_VERD].p);
    RVS_TRY(edges_dev_impl(h, d_q1, d_q2, E, d_valid));
    if (h->h_robot.self_collision && h->h_robot.kind == RVS_ROBOT_XARM6 && h->prm.precision != RVS_PREC_F64) {
        KArgs a = base_args(h, d_q1, E);
        a.q2 = d_q2; a.verdict = d_valid; a.d_c = nullptr;
        RVS_TRY(launched(h, h->lt.self_collision_edges(a, h->stream), "self collision edge kernel launch"));
    }
    RVS_TRY(d2h(h, valid, B_VERD, size_t(E)));
    return finish(h);
}
// Both halves of the edge test in ONE call, reported apart (RRTx::chooseParent needs to know WHICH one failed, RRTx.cpp:466-468):
// obstacle_ok[e] = isValid(q1_e, q2_e); self_hit[e] / q2_out[e] as rvs_self_collision_edges.
int rvs_check_edges_detailed(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *obstacle_ok, uint8_t *self_hit, float *q2_out) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "self-collision runs on REF32 / FAST32 handles");
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!q1 || !q2 || !obstacle_ok || !self_hit || !q2_out) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    const size_t bytes = size_t(E) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, 3 * bytes + 2 * size_t(E) + 64);
    RVS_TRY(h2d(h, B_Q, q1, bytes));
    RVS_TRY(h2d(h, B_Q2, q2, bytes));
    RVS_TRY(reserve(h, B_VERD, 2 * size_t(E)));
    RVS_TRY(reserve(h, B_DC, bytes));
    const float *d_q1 = static_cast<const float *>(h->buf[B_Q].p), *d_q2 = static_cast<const float *>(h->buf[B_Q2].p);
    uint8_t *d_ok = static_cast<uint8_t *>(h->buf[B_VERD].p), *d_hit = d_ok + E;
    RVS_TRY(edges_dev_impl(h, d_q1, d_q2, E, d_ok));
    KArgs a = base_args(h, d_q1, E);
    a.q2 = d_q2; a.verdict = d_hit; a.d_c = h->buf[B_DC].p;
    RVS_TRY(launched(h, h->lt.self_collision_edges(a, h->stream), "self collision edge kernel launch"));
    RVS_TRY(d2h(h, obstacle_ok, B_VERD, size_t(E)));
    RVS_TRY(d2h_from(h, self_hit, d_hit, size_t(E)));
    RVS_TRY(d2h(h, q2_out, B_DC, bytes));
    return finish(h);
}
int rvs_self_collision_edges_dev(rvs_handle h, const float *d_q1, const float *d_q2, int64_t E, uint8_t *d_hit, float *d_q2_out) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "self-collision runs on REF32 / FAST32 handles");
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!d_q1 || !d_q2 || !d_hit || !d_q2_out) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    KArgs a = base_args(h, d_q1, E);
    a.q2 = d_q2; a.verdict = d_hit; a.d_c = d_q2_out;
    return launched(h, h->lt.self_collision_edges(a, h->stream), "self collision edge kernel launch");
}
int rvs_self_collision_edges(rvs_handle h, const float *q1, const float *q2, int64_t E, uint8_t *hit, float *q2_out) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "self-collision runs on REF32 / FAST32 handles");
    if (E < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "E < 0");
    if (E == 0) return RVS_OK;
    if (!q1 || !q2 || !hit || !q2_out) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    const size_t bytes = size_t(E) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, 3 * bytes + size_t(E) + 64);
    RVS_TRY(h2d(h, B_Q, q1, bytes));
    RVS_TRY(h2d(h, B_Q2, q2, bytes));
    RVS_TRY(reserve(h, B_VERD, size_t(E)));
    RVS_TRY(reserve(h, B_DC, bytes));
    KArgs a = base_args(h, static_cast<const float *>(h->buf[B_Q].p), E);
    a.q2 = static_cast<const float *>(h->buf[B_Q2].p); a.verdict = static_cast<uint8_t *>(h->buf[B_VERD].p); a.d_c = h->buf[B_DC].p;
    RVS_TRY(launched(h, h->lt.self_collision_edges(a, h->stream), "self collision edge kernel launch"));
    RVS_TRY(d2h(h, hit, B_VERD, size_t(E)));
    RVS_TRY(d2h(h, q2_out, B_DC, bytes));
    return finish(h);
}


static int primitive_impl(rvs_handle h, const float *AB, const float *radius, const float *obs, int64_t K, float *out, uint8_t *hit) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "primitives run on REF32 / FAST32 handles");
    if (K < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "K < 0");
    if (K == 0) return RVS_OK;
    if (!AB || !radius || !obs || (!out && !hit)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    begin_call(h, size_t(K) * (13 * sizeof(float) + 8 * sizeof(float) + 1) + 128);
    RVS_TRY(h2d(h, B_Q, AB, size_t(K) * 6 * sizeof(float)));
    RVS_TRY(h2d(h, B_Q2, radius, size_t(K) * sizeof(float)));
    RVS_TRY(h2d(h, B_PLANES, obs, size_t(K) * 6 * sizeof(float)));
    RVS_TRY(reserve(h, B_DC, size_t(K) * 8 * sizeof(float)));
    RVS_TRY(reserve(h, B_VERD, size_t(K)));
    KArgs a = base_args(h, static_cast<const float *>(h->buf[B_Q].p), K);
    a.q2 = static_cast<const float *>(h->buf[B_Q2].p);
    a.planes = static_cast<const float *>(h->buf[B_PLANES].p);
    if (hit) a.verdict = static_cast<uint8_t *>(h->buf[B_VERD].p); else a.d_c = h->buf[B_DC].p;
    RVS_TRY(launched(h, h->lt.primitive(a, h->stream), "primitive kernel launch"));
    if (hit) RVS_TRY(d2h(h, hit, B_VERD, size_t(K))); else RVS_TRY(d2h(h, out, B_DC, size_t(K) * 8 * sizeof(float)));
    return finish(h);
}
int rvs_capsule_box_distance(rvs_handle h, const float *AB, const float *radius, const float *obs, int64_t K, float *out) {
    return primitive_impl(h, AB, radius, obs, K, out, nullptr);
}
int rvs_capsule_box_collision(rvs_handle h, const float *AB, const float *radius, const float *obs, int64_t K, uint8_t *hit) {
    return primitive_impl(h, AB, radius, obs, K, nullptr, hit);
}

int rvs_extend_spines(rvs_handle h, const float *q_root, const float *root_profile, float root_d_c, const float *root_planes,
                      const float *q_targets, int32_t S, const rvs_spine_params *sp, float *q_reached, float *d_c_reached, uint8_t *status) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "spines run in FP32: use a REF32/FAST32 handle");
    if (S < 0 || !sp || sp->num_layers < 1 || sp->num_iter_spine < 1) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad S or spine parameters");
    if (S == 0) return RVS_OK;
    if (!q_root || !root_profile || !q_targets || !q_reached || !d_c_reached || !status || (!root_planes && h->M > 0))
        return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    // one packed upload: [q_root n][root_profile n][q_targets S*n], planes separately; one packed download: [q_reached S*n][d_c S]
    const size_t n = size_t(h->h_robot.n), in_floats = 2 * n + size_t(S) * n, planes_floats = size_t(h->M) * n * 6;
    begin_call(h, (in_floats + planes_floats + size_t(S) * (n + 2)) * sizeof(float) + 128);
    std::vector<float> pack(in_floats);
    std::copy(q_root, q_root + n, pack.begin());
    std::copy(root_profile, root_profile + n, pack.begin() + n);
    std::copy(q_targets, q_targets + size_t(S) * n, pack.begin() + 2 * n);
    RVS_TRY(h2d(h, B_Q, pack.data(), in_floats * sizeof(float)));
    if (!h->zc) RVS_CUDA(h, cudaStreamSynchronize(h->stream));      // `pack` is pageable and dies with this frame
    RVS_TRY(h2d(h, B_PLANES, root_planes, planes_floats * sizeof(float)));
    RVS_TRY(reserve(h, B_DC, (size_t(S) * n + size_t(S)) * sizeof(float)));
    RVS_TRY(reserve(h, B_VERD, size_t(S)));
    const float *d_in = static_cast<const float *>(h->buf[B_Q].p);
    float *d_out = static_cast<float *>(h->buf[B_DC].p);
    SpineArgs a{};
    a.robot = h->d_robot; a.M = h->M; a.S = S;
    a.q_root = d_in; a.root_profile = d_in + n; a.q_targets = d_in + 2 * n; a.root_d_c = root_d_c;
    a.root_planes = static_cast<const float *>(h->buf[B_PLANES].p);
    a.num_layers = sp->num_layers; a.num_iter_spine = sp->num_iter_spine; a.use_expanded_bubble = sp->use_expanded_bubble;
    a.d_crit = sp->d_crit; a.max_dist = h->prm.max_distance; a.eq_thr = h->prm.equality_threshold;
    a.q_reached = d_out; a.d_c_reached = d_out + size_t(S) * n; a.status = static_cast<uint8_t *>(h->buf[B_VERD].p);
    RVS_TRY(launched(h, h->lt.spines(a, h->stream), "spine kernel launch"));
    RVS_TRY(d2h_from(h, q_reached, d_out, size_t(S) * n * sizeof(float)));
    RVS_TRY(d2h_from(h, d_c_reached, d_out + size_t(S) * n, size_t(S) * sizeof(float)));
    RVS_TRY(d2h(h, status, B_VERD, size_t(S)));
    return finish(h);
}

// ---- device-resident State cache -------------------------------------------------------------------------------------------
int rvs_store_reserve(rvs_handle h, int32_t capacity) {
    RVS_TRY(enter(h));
    if (capacity < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "capacity < 0");
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "the state cache is FP32: use a REF32/FAST32 handle");
    if (capacity <= h->store_cap) return RVS_OK;
    const size_t slot = store_slot_floats(h->h_robot.n, h->M) * sizeof(float);
    float *grown = nullptr;
    RVS_CUDA(h, cudaMalloc(&grown, slot * size_t(capacity)));
    if (h->d_store) {
        RVS_CUDA(h, cudaMemcpyAsync(grown, h->d_store, slot * size_t(h->store_cap), cudaMemcpyDeviceToDevice, h->stream));
        RVS_CUDA(h, cudaStreamSynchronize(h->stream));
        RVS_CUDA(h, cudaFree(h->d_store));
    }
    h->d_store = grown;
    h->store_cap = capacity;
    return RVS_OK;
}

int rvs_store_capacity(rvs_handle h) { return h ? h->store_cap : 0; }

namespace {
int check_slots(rvs_engine *h, const int32_t *slots, int32_t count) {
    for (int32_t i = 0; i < count; i++)
        if (slots[i] < 0 || slots[i] >= h->store_cap) return fail(h, RVS_ERR_INVALID_ARGUMENT, "slot index outside the reserved state cache");
    return RVS_OK;
}
StoreArgs store_args(rvs_engine *h) {
    StoreArgs a{};
    a.robot = h->d_robot; a.obstacles = h->d_obs; a.M = h->M; a.n = h->h_robot.n;
    a.eq_thr = h->prm.equality_threshold; a.max_dist = h->prm.max_distance; a.store = h->d_store;
    return a;
}
}  // namespace

int rvs_store_distance(rvs_handle h, const float *q, int32_t N, const int32_t *slots, float *d_c, float *profile) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "the state cache is FP32: use a REF32/FAST32 handle");
    if (N < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "N < 0");
    if (N == 0) return RVS_OK;
    if (!q || !slots || !d_c) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    RVS_TRY(check_slots(h, slots, N));
    const size_t n = size_t(h->h_robot.n);
    begin_call(h, size_t(N) * (n * 8 + 16) + 128);
    RVS_TRY(h2d(h, B_Q, q, size_t(N) * n * sizeof(float)));
    RVS_TRY(h2d(h, B_IDX, slots, size_t(N) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_DC, size_t(N) * sizeof(float)));
    RVS_TRY(reserve(h, B_PROF, size_t(N) * n * sizeof(float)));
    StoreArgs a = store_args(h);
    a.N = N; a.q = static_cast<const float *>(h->buf[B_Q].p); a.slots = static_cast<const int32_t *>(h->buf[B_IDX].p);
    a.d_c = static_cast<float *>(h->buf[B_DC].p); a.profile = profile ? static_cast<float *>(h->buf[B_PROF].p) : nullptr;
    RVS_TRY(launched(h, h->lt.store_distance(a, h->stream), "store distance kernel launch"));
    RVS_TRY(d2h(h, d_c, B_DC, size_t(N) * sizeof(float)));
    if (profile) RVS_TRY(d2h(h, profile, B_PROF, size_t(N) * n * sizeof(float)));
    return finish(h);
}

int rvs_store_extend_spines(rvs_handle h, const int32_t *root_slots, const float *q_targets, int32_t S, const rvs_spine_params *sp, int32_t gen2,
                            float *q_layers, int32_t *n_layers, float *d_c_last, uint8_t *status) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "spines run in FP32: use a REF32/FAST32 handle");
    if (S < 0 || !sp || sp->num_layers < 1 || sp->num_iter_spine < 1) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad S or spine parameters");
    if (S == 0) return RVS_OK;
    if (!root_slots || !q_targets || !q_layers || !n_layers || !d_c_last || !status) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    RVS_TRY(check_slots(h, root_slots, S));
    const size_t n = size_t(h->h_robot.n), layers = size_t(sp->num_layers);
    begin_call(h, size_t(S) * (n * 4 + 4 + layers * n * 4 + 4 + 4 + 1) + 256);
    RVS_TRY(h2d(h, B_Q, q_targets, size_t(S) * n * sizeof(float)));
    RVS_TRY(h2d(h, B_IDX, root_slots, size_t(S) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_NEAR, size_t(S) * layers * n * sizeof(float)));
    RVS_TRY(reserve(h, B_DC, size_t(S) * sizeof(float)));
    RVS_TRY(reserve(h, B_PROF, size_t(S) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_VERD, size_t(S)));
    StoreArgs a = store_args(h);
    a.S = S; a.root_slot = static_cast<const int32_t *>(h->buf[B_IDX].p); a.q_targets = static_cast<const float *>(h->buf[B_Q].p);
    a.num_layers = sp->num_layers; a.num_iter_spine = sp->num_iter_spine; a.use_expanded_bubble = sp->use_expanded_bubble; a.gen2 = gen2 ? 1 : 0;
    a.d_crit = sp->d_crit;
    a.q_layers = static_cast<float *>(h->buf[B_NEAR].p); a.n_layers = static_cast<int32_t *>(h->buf[B_PROF].p);
    a.d_c_last = static_cast<float *>(h->buf[B_DC].p); a.status = static_cast<uint8_t *>(h->buf[B_VERD].p);
    RVS_TRY(launched(h, h->lt.store_spines(a, h->stream), "store spine kernel launch"));
    RVS_TRY(d2h(h, q_layers, B_NEAR, size_t(S) * layers * n * sizeof(float)));
    RVS_TRY(d2h(h, n_layers, B_PROF, size_t(S) * sizeof(int32_t)));
    RVS_TRY(d2h(h, d_c_last, B_DC, size_t(S) * sizeof(float)));
    RVS_TRY(d2h(h, status, B_VERD, size_t(S)));
    return finish(h);
}

// connectGenSpine's loop (RGBTConnect.cpp:120-150) for R requests at once: up to `steps` times { generalized spine from the current
// state's slot towards the target; computeDistance of the state it reached into the next slot }, the 2 * steps kernels launched back to
// back and gating themselves on device state — one synchronisation per loop instead of two per step.
int rvs_store_connect(rvs_handle h, const int32_t *root_slots, const float *q_targets, const int32_t *new_slots, int32_t R, int32_t steps,
                      const rvs_spine_params *sp, float *q_layers, int32_t *n_layers, float *d_c_new, uint8_t *status) {
    RVS_TRY(enter(h));
    if (h->prm.precision == RVS_PREC_F64) return fail(h, RVS_ERR_UNSUPPORTED, "spines run in FP32: use a REF32/FAST32 handle");
    if (R < 0 || steps < 1 || steps > 64 || !sp || sp->num_layers < 1 || sp->num_iter_spine < 1) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad R, steps (1..64) or spine parameters");
    if (R == 0) return RVS_OK;
    if (!root_slots || !q_targets || !new_slots || !q_layers || !n_layers || !d_c_new || !status) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    RVS_TRY(check_slots(h, root_slots, R));
    const size_t n = size_t(h->h_robot.n), layers = size_t(sp->num_layers), Rs = size_t(R), K = size_t(steps);
    // per-step tables in [step][request] order: root slot of the step, slot of its result (-1: this request makes fewer steps)
    std::vector<int32_t> tab(2 * K * Rs + Rs, 1);                            // ... followed by the chain states (all alive)
    for (size_t r = 0; r < Rs; r++)
        for (size_t k = 0; k < K; k++) {
            const int32_t fresh = new_slots[r * K + k];
            if (fresh >= 0) RVS_TRY(check_slots(h, &fresh, 1));
            tab[K * Rs + k * Rs + r] = fresh;
            tab[k * Rs + r] = k == 0 ? root_slots[r] : new_slots[r * K + k - 1];
            if (tab[k * Rs + r] < 0) { tab[k * Rs + r] = root_slots[r]; tab[K * Rs + k * Rs + r] = -1; }     // (a gap ends the chain)
        }
    begin_call(h, kSmallCall + 1);                                          // device staging: the kernels of a chain read each other's results
    RVS_TRY(h2d(h, B_Q, q_targets, Rs * n * sizeof(float)));
    RVS_TRY(h2d(h, B_IDX, tab.data(), tab.size() * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_NEAR, K * Rs * layers * n * sizeof(float)));
    RVS_TRY(reserve(h, B_DC, 2 * K * Rs * sizeof(float)));
    RVS_TRY(reserve(h, B_PROF, K * Rs * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_VERD, K * Rs));
    const int32_t *d_tab = static_cast<const int32_t *>(h->buf[B_IDX].p);
    float *d_layers = static_cast<float *>(h->buf[B_NEAR].p), *d_dc = static_cast<float *>(h->buf[B_DC].p);
    int32_t *d_nl = static_cast<int32_t *>(h->buf[B_PROF].p);
    uint8_t *d_st = static_cast<uint8_t *>(h->buf[B_VERD].p);
    for (size_t k = 0; k < K; k++) {
        StoreArgs a = store_args(h);
        a.S = R; a.N = R;
        a.root_slot = d_tab + k * Rs; a.slots = d_tab + K * Rs + k * Rs; a.chain_state = const_cast<int32_t *>(d_tab) + 2 * K * Rs;
        a.q_targets = static_cast<const float *>(h->buf[B_Q].p);
        a.num_layers = sp->num_layers; a.num_iter_spine = sp->num_iter_spine; a.use_expanded_bubble = sp->use_expanded_bubble; a.gen2 = 1;
        a.d_crit = sp->d_crit;
        a.q_layers = d_layers + k * Rs * layers * n; a.n_layers = d_nl + k * Rs; a.d_c_last = d_dc + K * Rs + k * Rs; a.status = d_st + k * Rs;
        RVS_TRY(launched(h, h->lt.store_spines(a, h->stream), "connect chain: spine kernel launch"));
        a.q = a.q_layers; a.layers_n = a.n_layers; a.d_c = d_dc + k * Rs; a.profile = nullptr;
        RVS_TRY(launched(h, h->lt.store_distance(a, h->stream), "connect chain: distance kernel launch"));
    }
    std::vector<float> layers_t(K * Rs * layers * n), dc_t(K * Rs);
    std::vector<int32_t> nl_t(K * Rs);
    std::vector<uint8_t> st_t(K * Rs);
    RVS_TRY(d2h(h, layers_t.data(), B_NEAR, layers_t.size() * sizeof(float)));
    RVS_TRY(d2h(h, dc_t.data(), B_DC, dc_t.size() * sizeof(float)));
    RVS_TRY(d2h(h, nl_t.data(), B_PROF, nl_t.size() * sizeof(int32_t)));
    RVS_TRY(d2h(h, st_t.data(), B_VERD, st_t.size()));
    RVS_TRY(finish(h));
    for (size_t r = 0; r < Rs; r++)                                         // [step][request] -> the caller's [request][step]
        for (size_t k = 0; k < K; k++) {
            std::copy_n(layers_t.begin() + (k * Rs + r) * layers * n, layers * n, q_layers + (r * K + k) * layers * n);
            n_layers[r * K + k] = nl_t[k * Rs + r]; d_c_new[r * K + k] = dc_t[k * Rs + r]; status[r * K + k] = st_t[k * Rs + r];
        }
    return RVS_OK;
}

int rvs_store_read(rvs_handle h, int32_t slot, float *q, float *d_c, float *profile, float *planes) {
    RVS_TRY(enter(h));
    RVS_TRY(check_slots(h, &slot, 1));
    const size_t n = size_t(h->h_robot.n);
    const size_t stride = store_slot_floats(int(n), h->M), head = store_header_floats(int(n));
    std::vector<float> tmp(stride);
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    RVS_CUDA(h, cudaMemcpy(tmp.data(), h->d_store + size_t(slot) * stride, stride * sizeof(float), cudaMemcpyDeviceToHost));
    if (q) std::copy(tmp.begin(), tmp.begin() + n, q);
    if (d_c) *d_c = tmp[n];
    if (profile) std::copy(tmp.begin() + n + 1, tmp.begin() + 2 * n + 1, profile);
    if (planes) std::copy(tmp.begin() + head, tmp.end(), planes);
    return RVS_OK;
}


// ---- dynamic environment (rvs_env.cuh) ---------------------------------------------------------------------------------
namespace {
EnvArgs env_args(rvs_engine *h, int32_t K, float dt) {
    EnvArgs a{};
    a.robot = h->d_robot; a.state = h->d_env; a.timeline = h->d_env_tl; a.M = h->M; a.K = K; a.dt = dt;
    a.motion_type = h->env_prm.motion_type;
    for (int c = 0; c < 3; c++) a.ws_center[c] = h->env_prm.ws_center[c];
    a.ws_radius = h->env_prm.ws_radius; a.base_radius = h->env_prm.base_radius; a.robot_max_vel = h->env_prm.robot_max_vel;
    a.first_invalid = h->d_first_invalid; a.commit_edge = -1;
    a.eq_thr = h->prm.equality_threshold; a.max_dist = h->prm.max_distance;
    a.obstacles = h->d_obs; a.obstacles_scan = h->d_obs + h->obs_cap; a.scan_dist = reinterpret_cast<float *>(h->d_obs + 2 * size_t(h->obs_cap));
    a.obstacles_pack = h->d_obs_pack;
    a.scan_key = h->d_scan_key;
    a.env_ground_included = h->env_ground_included;
    a.rand_state = h->d_rand; a.rand_timeline = h->d_rand ? h->d_rand + 1 : nullptr;
    return a;
}
// (re)allocate the stream + its timeline of K states, keeping the current stream
int env_rand_ready(rvs_engine *h, size_t K) {
    if (h->d_rand && K <= h->rand_tl_cap) return RVS_OK;
    const size_t cap = std::max<size_t>(K, 16);
    LibcRand *grown = nullptr;
    RVS_CUDA(h, cudaMalloc(&grown, (1 + cap) * sizeof(LibcRand)));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_rand) {
        RVS_CUDA(h, cudaMemcpy(grown, h->d_rand, sizeof(LibcRand), cudaMemcpyDeviceToDevice));
        RVS_CUDA(h, cudaFree(h->d_rand));
    } else {
        LibcRand g;
        g.seed(1);                                                         // rand() without srand() behaves as after srand(1)
        RVS_CUDA(h, cudaMemcpy(grown, &g, sizeof(LibcRand), cudaMemcpyHostToDevice));
    }
    h->d_rand = grown; h->rand_tl_cap = cap;
    return RVS_OK;
}
int env_ready(rvs_engine *h, int32_t K, float dt) {
    if (!h->env_on) return fail(h, RVS_ERR_INVALID_ARGUMENT, "no dynamic environment: call rvs_env_set first");
    if (K < 1 || !(dt > 0.f)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "steps / K >= 1 and delta_time > 0 required");
    const size_t need = size_t(K) * size_t(h->M);
    if (need > h->env_tl_cap) {
        if (h->d_env_tl) RVS_CUDA(h, cudaFree(h->d_env_tl));
        h->d_env_tl = nullptr; h->env_tl_cap = 0;
        RVS_CUDA(h, cudaMalloc(&h->d_env_tl, std::max<size_t>(need, 1) * sizeof(EnvRec)));
        h->env_tl_cap = need;
    }
    if (h->env_prm.motion_type == RVS_MOTION_RANDOM_DIRECTIONS) RVS_TRY(env_rand_ready(h, size_t(K)));
    return RVS_OK;
}
}  // namespace

int rvs_env_set(rvs_handle h, const rvs_env_params *ep, const float *dim, const float *pos, const uint8_t *rotated, const float *vel, const float *max_vel,
                const uint8_t *is_dynamic, const uint8_t *is_ground, const float *min_dist_tol, int32_t M) {
    RVS_TRY(enter(h));
    if (!ep || M < 0 || (M > 0 && (!dim || !pos))) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad environment arrays");
    if (ep->motion_type < RVS_MOTION_STRAIGHT || ep->motion_type > RVS_MOTION_LIGHT_DIRECTIONS) return fail(h, RVS_ERR_INVALID_ARGUMENT, "unknown motion_type");
    std::vector<EnvRec> recs(size_t(M), EnvRec{});
    std::vector<float> mn(size_t(M) * 3), mx(size_t(M) * 3);
    for (int j = 0; j < M; j++) {
        EnvRec &r = recs[size_t(j)];
        float half[3];
        for (int c = 0; c < 3; c++) half[c] = 0.5f * dim[3 * j + c];
        if (rotated && rotated[j]) {          // FCL: AABB of the bounding sphere for a non-identity rotation
            const float rad = std::sqrt(half[0] * half[0] + half[1] * half[1] + half[2] * half[2]);
            half[0] = half[1] = half[2] = rad;
        }
        for (int c = 0; c < 3; c++) {
            r.half[c] = half[c]; r.pos[c] = pos[3 * j + c]; r.vel[c] = vel ? vel[3 * j + c] : 0.f;
            mn[size_t(3 * j + c)] = r.pos[c] + (-half[c]); mx[size_t(3 * j + c)] = r.pos[c] + half[c];
        }
        r.max_vel = max_vel ? max_vel[j] : 0.f; r.path_len = 0.f; r.sign = 1;             // Environment.cpp:76-77
        r.dynamic = (is_dynamic && is_dynamic[j]) ? 1u : 0u;
        r.tol = min_dist_tol ? min_dist_tol[j] : h->prm.max_distance;
        r.flags = (is_ground && is_ground[j]) ? 1u : 0u;
    }
    h->env_ground_included = h->h_robot.ground_included;                    // Environment.cpp:64-68: the robot node's ground_included
    RVS_TRY(rvs_set_obstacles(h, mn.data(), mx.data(), is_ground, min_dist_tol, M));
    if (M > h->env_cap) {
        if (h->d_env) RVS_CUDA(h, cudaFree(h->d_env));
        if (h->d_scan_key) RVS_CUDA(h, cudaFree(h->d_scan_key));
        h->d_env = nullptr; h->d_scan_key = nullptr; h->env_cap = 0;
        const int cap = M + M / 2 + 16;
        RVS_CUDA(h, cudaMalloc(&h->d_env, size_t(cap) * sizeof(EnvRec)));
        RVS_CUDA(h, cudaMalloc(&h->d_scan_key, size_t(cap) * sizeof(float)));
        h->env_cap = cap;
    }
    if (M > 0) RVS_CUDA(h, cudaMemcpy(h->d_env, recs.data(), size_t(M) * sizeof(EnvRec), cudaMemcpyHostToDevice));
    h->env_prm = *ep;
    h->env_on = true;
    return RVS_OK;
}

// random_directions' extra state: accelerations, the ground rule of Environment::isValid, and the libc stream (srand(seed))
int rvs_env_set_random_directions(rvs_handle h, const float *acc, int32_t ground_included, int64_t seed) {
    RVS_TRY(enter(h));
    if (!h->env_on) return fail(h, RVS_ERR_INVALID_ARGUMENT, "no dynamic environment: call rvs_env_set first");
    if (seed > 0xFFFFFFFFll) return fail(h, RVS_ERR_INVALID_ARGUMENT, "seed does not fit srand()'s unsigned int");
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (acc && h->M > 0) {
        std::vector<EnvRec> recs(static_cast<size_t>(h->M), EnvRec{});
        RVS_CUDA(h, cudaMemcpy(recs.data(), h->d_env, size_t(h->M) * sizeof(EnvRec), cudaMemcpyDeviceToHost));
        for (int j = 0; j < h->M; j++)
            for (int c = 0; c < 3; c++) recs[size_t(j)].acc[c] = acc[3 * j + c];
        RVS_CUDA(h, cudaMemcpy(h->d_env, recs.data(), size_t(h->M) * sizeof(EnvRec), cudaMemcpyHostToDevice));
    }
    if (ground_included >= 0) h->env_ground_included = ground_included;
    RVS_TRY(env_rand_ready(h, 1));
    if (seed >= 0) {
        LibcRand g;
        g.seed(uint32_t(seed));
        RVS_CUDA(h, cudaMemcpy(h->d_rand, &g, sizeof(LibcRand), cudaMemcpyHostToDevice));
    }
    return RVS_OK;
}
int rvs_env_get_random_directions(rvs_handle h, float *acc, uint32_t *draws) {
    RVS_TRY(enter(h));
    if (!h->env_on) return fail(h, RVS_ERR_INVALID_ARGUMENT, "no dynamic environment: call rvs_env_set first");
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (acc && h->M > 0) {
        std::vector<EnvRec> recs(static_cast<size_t>(h->M), EnvRec{});
        RVS_CUDA(h, cudaMemcpy(recs.data(), h->d_env, size_t(h->M) * sizeof(EnvRec), cudaMemcpyDeviceToHost));
        for (int j = 0; j < h->M; j++)
            for (int c = 0; c < 3; c++) acc[3 * j + c] = recs[size_t(j)].acc[c];
    }
    if (draws) {
        *draws = 0;
        if (h->d_rand) {
            LibcRand g;
            RVS_CUDA(h, cudaMemcpy(&g, h->d_rand, sizeof(LibcRand), cudaMemcpyDeviceToHost));
            *draws = g.draws;
        }
    }
    return RVS_OK;
}

int rvs_env_step(rvs_handle h, float delta_time, int32_t steps) {
    RVS_TRY(enter(h));
    if (steps == 0) return RVS_OK;
    RVS_TRY(env_ready(h, steps, delta_time));
    if (h->M == 0) return RVS_OK;
    EnvArgs a = env_args(h, steps, delta_time);
    RVS_TRY(launched(h, launch_env_advance(a, h->stream), "environment advance launch"));
    RVS_TRY(launched(h, launch_env_commit(a, h->stream), "environment commit launch"));
    h->df_dirty = true;
    h->launches++;                                                          // commit = two grids
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    return RVS_OK;
}

int rvs_env_get(rvs_handle h, float *pos, float *vel) {
    RVS_TRY(enter(h));
    if (!h->env_on) return fail(h, RVS_ERR_INVALID_ARGUMENT, "no dynamic environment: call rvs_env_set first");
    std::vector<EnvRec> recs(static_cast<size_t>(h->M), EnvRec{});
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->M > 0) RVS_CUDA(h, cudaMemcpy(recs.data(), h->d_env, size_t(h->M) * sizeof(EnvRec), cudaMemcpyDeviceToHost));
    for (int j = 0; j < h->M; j++)
        for (int c = 0; c < 3; c++) {
            if (pos) pos[3 * j + c] = recs[size_t(j)].pos[c];
            if (vel) vel[3 * j + c] = recs[size_t(j)].vel[c];
        }
    return RVS_OK;
}

// q_prev / q_cur [E][n] (points == nullptr) or points [E][K][n]
static int motion_validity_impl(rvs_handle h, const float *q_prev, const float *q_cur, const float *points, int64_t E, int32_t K, float delta_time,
                                int32_t commit_edge, uint8_t *valid, int32_t *first_invalid) {
    RVS_TRY(enter(h));
    if (E < 0 || commit_edge >= E) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad E or commit_edge");
    if (E == 0) return RVS_OK;
    if ((points ? false : (!q_prev || !q_cur)) || !valid) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    RVS_TRY(env_ready(h, K, delta_time));
    if (size_t(E) > h->env_fi_cap) {
        if (h->d_first_invalid) RVS_CUDA(h, cudaFree(h->d_first_invalid));
        h->d_first_invalid = nullptr; h->env_fi_cap = 0;
        const size_t cap = size_t(E) + size_t(E) / 2 + 64;
        RVS_CUDA(h, cudaMalloc(&h->d_first_invalid, cap * sizeof(int32_t)));
        h->env_fi_cap = cap;
    }
    const size_t bytes = size_t(E) * size_t(h->h_robot.n) * sizeof(float);
    begin_call(h, size_t(-1));                                              // device staging (first_invalid lives in device memory anyway)
    if (points) RVS_TRY(h2d(h, B_Q, points, bytes * size_t(K)));
    else {
        RVS_TRY(h2d(h, B_Q, q_prev, bytes));
        RVS_TRY(h2d(h, B_Q2, q_cur, bytes));
    }
    RVS_CUDA(h, cudaMemsetAsync(h->d_first_invalid, 0x7f, size_t(E) * sizeof(int32_t), h->stream));      // 0x7f7f7f7f: "no invalid point"
    EnvArgs a = env_args(h, K, delta_time);
    if (points) { a.points = static_cast<const float *>(h->buf[B_Q].p); a.q_prev = a.q_cur = nullptr; }
    else { a.points = nullptr; a.q_prev = static_cast<const float *>(h->buf[B_Q].p); a.q_cur = static_cast<const float *>(h->buf[B_Q2].p); }
    a.E = E; a.commit_edge = commit_edge;
    if (h->M > 0) RVS_TRY(launched(h, launch_env_advance(a, h->stream), "environment advance launch"));
    RVS_TRY(launched(h, h->lt.motion_valid(a, h->stream), "motion validity launch"));
    if (commit_edge >= 0 && h->M > 0) {
        RVS_TRY(launched(h, launch_env_commit(a, h->stream), "environment commit launch"));
    h->df_dirty = true;
        h->launches++;
    }
    std::vector<int32_t> fi(static_cast<size_t>(E), 0);
    RVS_CUDA(h, cudaMemcpyAsync(fi.data(), h->d_first_invalid, size_t(E) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    RVS_TRY(finish(h));
    for (int64_t e = 0; e < E; e++) {
        const bool ok = fi[size_t(e)] > K;
        valid[e] = ok ? 1 : 0;
        if (first_invalid) first_invalid[e] = ok ? 0 : fi[size_t(e)];
    }
    return RVS_OK;
}

int rvs_motion_validity(rvs_handle h, const float *q_prev, const float *q_cur, int64_t E, int32_t K, float delta_time, int32_t commit_edge, uint8_t *valid,
                        int32_t *first_invalid) {
    return motion_validity_impl(h, q_prev, q_cur, nullptr, E, K, delta_time, commit_edge, valid, first_invalid);
}

int rvs_motion_validity_points(rvs_handle h, const float *points, int64_t E, int32_t K, float delta_time, int32_t commit_edge, uint8_t *valid,
                               int32_t *first_invalid) {
    if (!points && E > 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    return motion_validity_impl(h, nullptr, nullptr, points, E, K, delta_time, commit_edge, valid, first_invalid);
}

}  // extern "C"

// ---- device-resident planner trees: nearest neighbour and radius queries (rvs_nn.cuh) --------------------------------------------
namespace {
int tree_tables(rvs_engine *h) {
    if (!h->tree_tab_dirty) return RVS_OK;
    const int T = int(h->trees.size());
    if (T > h->tree_tab_cap) {
        if (h->d_tree_ptr) RVS_CUDA(h, cudaFree(h->d_tree_ptr));
        if (h->d_tree_cnt) RVS_CUDA(h, cudaFree(h->d_tree_cnt));
        h->d_tree_ptr = nullptr; h->d_tree_cnt = nullptr;
        const int cap = T + T / 2 + 16;
        RVS_CUDA(h, cudaMalloc(&h->d_tree_ptr, size_t(cap) * sizeof(float *)));
        RVS_CUDA(h, cudaMalloc(&h->d_tree_cnt, size_t(cap) * sizeof(int32_t)));
        h->tree_tab_cap = cap;
    }
    std::vector<const float *> ptr(static_cast<size_t>(T), nullptr);
    std::vector<int32_t> cnt(static_cast<size_t>(T), 0);
    for (int t = 0; t < T; t++) { ptr[size_t(t)] = h->trees[size_t(t)].d; cnt[size_t(t)] = h->trees[size_t(t)].count; }
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (T > 0) {
        RVS_CUDA(h, cudaMemcpy(h->d_tree_ptr, ptr.data(), size_t(T) * sizeof(float *), cudaMemcpyHostToDevice));
        RVS_CUDA(h, cudaMemcpy(h->d_tree_cnt, cnt.data(), size_t(T) * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    h->tree_tab_dirty = false;
    return RVS_OK;
}
}  // namespace

extern "C" {
int rvs_tree_create(rvs_handle h, int32_t *tree) {
    RVS_TRY(enter(h));
    if (!tree) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL tree id");
    for (size_t t = 0; t < h->trees.size(); t++)
        if (h->trees[t].released) { h->trees[t].released = false; h->trees[t].count = 0; *tree = int32_t(t); return RVS_OK; }     // reuse a released id
    h->trees.emplace_back();
    h->tree_tab_dirty = true;
    *tree = int32_t(h->trees.size() - 1);
    return RVS_OK;
}
int rvs_tree_release(rvs_handle h, int32_t tree) {
    RVS_TRY(enter(h));
    if (tree < 0 || size_t(tree) >= h->trees.size() || h->trees[size_t(tree)].released) return fail(h, RVS_ERR_INVALID_ARGUMENT, "unknown tree");
    auto &t = h->trees[size_t(tree)];
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (t.d) RVS_CUDA(h, cudaFree(t.d));
    t.d = nullptr; t.count = 0; t.cap = 0; t.released = true;
    h->tree_tab_dirty = true;
    return RVS_OK;
}
int rvs_tree_clear(rvs_handle h, int32_t tree) {
    RVS_TRY(enter(h));
    if (tree < 0 || size_t(tree) >= h->trees.size() || h->trees[size_t(tree)].released) return fail(h, RVS_ERR_INVALID_ARGUMENT, "unknown tree");
    h->trees[size_t(tree)].count = 0;
    h->tree_tab_dirty = true;
    return RVS_OK;
}
int rvs_tree_size(rvs_handle h, int32_t tree) {
    if (!h || tree < 0 || size_t(tree) >= h->trees.size()) return -1;
    return h->trees[size_t(tree)].count;
}
int rvs_tree_append(rvs_handle h, int32_t tree, const float *q, int32_t count) {
    RVS_TRY(enter(h));
    if (tree < 0 || size_t(tree) >= h->trees.size() || h->trees[size_t(tree)].released || count < 0 || (count > 0 && !q)) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad tree append");
    if (count == 0) return RVS_OK;
    auto &t = h->trees[size_t(tree)];
    const size_t row = size_t(h->h_robot.n) * sizeof(float);
    if (t.count + count > t.cap) {
        const int cap = std::max(2 * t.cap, std::max(t.count + count, 1024));
        float *grown = nullptr;
        RVS_CUDA(h, cudaMalloc(&grown, size_t(cap) * row));
        RVS_CUDA(h, cudaStreamSynchronize(h->stream));
        if (t.d && t.count > 0) RVS_CUDA(h, cudaMemcpy(grown, t.d, size_t(t.count) * row, cudaMemcpyDeviceToDevice));
        if (t.d) RVS_CUDA(h, cudaFree(t.d));
        t.d = grown; t.cap = cap;
    }
    RVS_CUDA(h, cudaMemcpyAsync(t.d + size_t(t.count) * size_t(h->h_robot.n), q, size_t(count) * row, cudaMemcpyHostToDevice, h->stream));
    RVS_CUDA(h, cudaStreamSynchronize(h->stream));                    // q is the caller's pageable memory
    t.count += count;
    h->tree_tab_dirty = true;
    return RVS_OK;
}
int rvs_tree_nearest(rvs_handle h, const int32_t *trees, const float *queries, int32_t Q, int32_t *idx, float *dist2) {
    RVS_TRY(enter(h));
    if (Q < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "Q < 0");
    if (Q == 0) return RVS_OK;
    if (!queries || !idx) return fail(h, RVS_ERR_INVALID_ARGUMENT, "NULL buffer");
    if (h->trees.empty()) return fail(h, RVS_ERR_INVALID_ARGUMENT, "no tree");
    if (trees)
        for (int k = 0; k < Q; k++)
            if (trees[k] < 0 || size_t(trees[k]) >= h->trees.size() || h->trees[size_t(trees[k])].released) return fail(h, RVS_ERR_INVALID_ARGUMENT, "unknown tree");
    RVS_TRY(tree_tables(h));
    const size_t n = size_t(h->h_robot.n);
    begin_call(h, size_t(Q) * (n * 4 + 12) + 64);
    RVS_TRY(h2d(h, B_Q, queries, size_t(Q) * n * sizeof(float)));
    if (trees) RVS_TRY(h2d(h, B_IDX, trees, size_t(Q) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_PROF, size_t(Q) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_DC, size_t(Q) * sizeof(float)));
    NnArgs a{};
    a.points = h->d_tree_ptr; a.counts = h->d_tree_cnt; a.n = int32_t(n);
    a.tree_of = trees ? static_cast<const int32_t *>(h->buf[B_IDX].p) : nullptr;
    a.queries = static_cast<const float *>(h->buf[B_Q].p); a.Q = Q;
    a.idx = static_cast<int32_t *>(h->buf[B_PROF].p); a.dist2 = static_cast<float *>(h->buf[B_DC].p);
    RVS_TRY(launched(h, launch_nn_nearest(a, h->stream), "nearest-neighbour kernel launch"));
    RVS_TRY(d2h(h, idx, B_PROF, size_t(Q) * sizeof(int32_t)));
    if (dist2) RVS_TRY(d2h(h, dist2, B_DC, size_t(Q) * sizeof(float)));
    return finish(h);
}
int rvs_tree_radius(rvs_handle h, int32_t tree, const float *query, float radius, int32_t exclude, int32_t capacity, int32_t *idx, float *dist,
                    int32_t *count) {
    RVS_TRY(enter(h));
    if (tree < 0 || size_t(tree) >= h->trees.size() || h->trees[size_t(tree)].released || !query || !idx || !dist || !count || capacity < 0) return fail(h, RVS_ERR_INVALID_ARGUMENT, "bad radius query");
    RVS_TRY(tree_tables(h));
    const size_t n = size_t(h->h_robot.n);
    begin_call(h, n * 4 + size_t(capacity) * 8 + 64);
    RVS_TRY(h2d(h, B_Q, query, n * sizeof(float)));
    RVS_TRY(reserve(h, B_PROF, (size_t(capacity) + 1) * sizeof(int32_t)));
    RVS_TRY(reserve(h, B_DC, size_t(capacity) * sizeof(float)));
    NnArgs a{};
    a.points = h->d_tree_ptr; a.counts = h->d_tree_cnt; a.n = int32_t(n);
    a.queries = static_cast<const float *>(h->buf[B_Q].p);
    a.tree = tree; a.radius = radius; a.exclude = exclude; a.capacity = capacity;
    a.out_idx = static_cast<int32_t *>(h->buf[B_PROF].p) + 1; a.out_dist = static_cast<float *>(h->buf[B_DC].p);
    a.out_count = static_cast<int32_t *>(h->buf[B_PROF].p);
    RVS_TRY(launched(h, launch_nn_radius(a, h->stream), "radius kernel launch"));
    RVS_TRY(d2h(h, count, B_PROF, sizeof(int32_t)));
    // the hit lists are copied in full (the count is only known after the synchronisation); capacity bounds them
    RVS_TRY(d2h_from(h, idx, static_cast<int32_t *>(h->buf[B_PROF].p) + 1, size_t(capacity) * sizeof(int32_t)));
    RVS_TRY(d2h(h, dist, B_DC, size_t(capacity) * sizeof(float)));
    return finish(h);
}
}  // extern "C"
