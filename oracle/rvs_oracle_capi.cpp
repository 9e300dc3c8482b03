// rvs_oracle_capi.cpp — C entry points of the CPU oracle for ctypes (TEST INFRASTRUCTURE ONLY; see rvs_oracle.hpp).
// Mirrors the argument layout of include/rvs_cuda.h so the parity tests can feed both sides the same arrays.

#include "rvs_oracle_engine.hpp"
#include "rvs_oracle_env.hpp"

using namespace orc_capi;

namespace {

template <class T>
int distance_impl(Engine &e, const float *q, int64_t N, int mode, T *d_c, T *profile, T *nearest, int threads) {
    const auto cfg = config_of<T>(e.prm);
    const auto &sc = Pick<T>::scene(e);
    const int n = e.robot.n;
    const size_t M = sc.boxes.size();
    parallel_for(N, threads, [&](int64_t b, int64_t en) {
        std::vector<orc::V3<T>> sk(n + 1);
        std::vector<T> prof(n);
        uint64_t visited = 0;
        for (int64_t i = b; i < en; i++) {
            orc::skeleton(e.robot, q + i * n, sk.data());
            bool hit;
            T *np = (mode == RVS_DIST_FULL && nearest) ? nearest + size_t(i) * M * n * 6 : nullptr;
            d_c[i] = orc::distance(e.robot, sc, sk.data(), cfg, prof.data(), np, &hit, &visited);
            if (mode >= RVS_DIST_PROFILE && profile)
                for (int k = 0; k < n; k++) profile[i * n + k] = prof[k];
        }
        e.pair_count += visited;
    });
    return 0;
}

template <class T>
int is_valid_impl(const Engine &e, const float *q, int64_t N, uint8_t *out, int threads) {
    const auto cfg = config_of<T>(e.prm);
    const auto &sc = Pick<T>::scene(e);
    const int n = e.robot.n;
    parallel_for(N, threads, [&](int64_t b, int64_t en) {
        std::vector<orc::V3<T>> sk(n + 1);
        for (int64_t i = b; i < en; i++) {
            orc::skeleton(e.robot, q + i * n, sk.data());
            out[i] = orc::is_valid(e.robot, sc, sk.data(), cfg) ? 1 : 0;
        }
    });
    return 0;
}

template <class T> orc::V3<T> v3(const T *p) { return {p[0], p[1], p[2]}; }
template <class T> orc::Box6<T> box6(const T *p) { orc::Box6<T> b; for (int i = 0; i < 6; i++) b.o[i] = p[i]; return b; }
template <class T> void put_near(const orc::Near<T> &r, T *out) {
    out[0] = r.d; out[1] = r.has ? T(1) : T(0);
    out[2] = r.P1.x; out[3] = r.P1.y; out[4] = r.P1.z; out[5] = r.P2.x; out[6] = r.P2.y; out[7] = r.P2.z;
}

}  // namespace

extern "C" {

void *orc_create(const rvs_robot_desc *rd, const rvs_params *prm) {
    if (!rd || !prm || rd->num_dofs < 1 || rd->num_dofs > RVS_MAX_DOFS) return nullptr;
    auto *e = new Engine();
    e->prm = *prm;
    orc::Robot &rb = e->robot;
    rb.kind = rd->kind;
    rb.n = rd->num_dofs;
    rb.ground_included = rd->ground_included;
    rb.gripper_length = rd->gripper_length;
    for (int k = 0; k <= rb.n; k++) {
        rb.fixed_rot.push_back(orc::rpy_matrix(rd->joint_rpy[k][0], rd->joint_rpy[k][1], rd->joint_rpy[k][2]));
        rb.xyz.push_back({rd->joint_xyz[k][0], rd->joint_xyz[k][1], rd->joint_xyz[k][2]});
    }
    for (int k = 0; k < rb.n; k++) {
        const double *a = rd->joint_axis[k];
        const double len = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);   // KDL::Joint normalises the axis
        rb.axis.push_back({a[0] / len, a[1] / len, a[2] / len});
        rb.radius.push_back(rd->capsule_radius[k]);
        e->lo.push_back(rd->limit_lo[k]);
        e->hi.push_back(rd->limit_hi[k]);
    }
    return e;
}

void orc_destroy(void *h) { delete static_cast<Engine *>(h); }

uint64_t orc_pair_count(void *h, int reset) {
    auto *e = static_cast<Engine *>(h);
    const uint64_t v = e->pair_count.load();
    if (reset) e->pair_count = 0;
    return v;
}

int orc_set_params(void *h, const rvs_params *prm) { static_cast<Engine *>(h)->prm = *prm; return 0; }

int orc_set_obstacles(void *h, const float *mn, const float *mx, const uint8_t *ground, const float *tol, int32_t M) {
    auto *e = static_cast<Engine *>(h);
    e->sc32 = {};
    e->sc64 = {};
    for (int j = 0; j < M; j++) {
        orc::Box6<float> b{{mn[3 * j], mn[3 * j + 1], mn[3 * j + 2], mx[3 * j], mx[3 * j + 1], mx[3 * j + 2]}};
        orc::Box6<double> bd;
        for (int k = 0; k < 6; k++) bd.o[k] = double(b.o[k]);
        const float t = tol ? tol[j] : e->prm.max_distance;
        e->sc32.boxes.push_back(b); e->sc32.is_ground.push_back(ground ? ground[j] : 0); e->sc32.min_dist_tol.push_back(t);
        e->sc64.boxes.push_back(bd); e->sc64.is_ground.push_back(ground ? ground[j] : 0); e->sc64.min_dist_tol.push_back(double(t));
    }
    return 0;
}

int orc_skeleton_f32(void *h, const float *q, int64_t N, float *sk) {
    auto *e = static_cast<Engine *>(h);
    const int n = e->robot.n;
    std::vector<orc::V3<float>> s(n + 1);
    for (int64_t i = 0; i < N; i++) {
        orc::skeleton(e->robot, q + i * n, s.data());
        for (int k = 0; k <= n; k++) { float *d = sk + (i * (n + 1) + k) * 3; d[0] = s[k].x; d[1] = s[k].y; d[2] = s[k].z; }
    }
    return 0;
}
int orc_skeleton_f64(void *h, const float *q, int64_t N, double *sk) {
    auto *e = static_cast<Engine *>(h);
    const int n = e->robot.n;
    std::vector<orc::V3<double>> s(n + 1);
    for (int64_t i = 0; i < N; i++) {
        orc::skeleton(e->robot, q + i * n, s.data());
        for (int k = 0; k <= n; k++) { double *d = sk + (i * (n + 1) + k) * 3; d[0] = s[k].x; d[1] = s[k].y; d[2] = s[k].z; }
    }
    return 0;
}

int orc_is_valid_f32(void *h, const float *q, int64_t N, uint8_t *out, int threads) { return is_valid_impl<float>(*static_cast<Engine *>(h), q, N, out, threads); }
int orc_is_valid_f64(void *h, const float *q, int64_t N, uint8_t *out, int threads) { return is_valid_impl<double>(*static_cast<Engine *>(h), q, N, out, threads); }

int orc_is_valid_edges_f32(void *h, const float *q1, const float *q2, int64_t E, uint8_t *out, int threads) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    parallel_for(E, threads, [&](int64_t b, int64_t en) {
        for (int64_t i = b; i < en; i++) out[i] = orc::is_valid_edge(e->robot, e->sc32, q1 + i * n, q2 + i * n, cfg) ? 1 : 0;
    });
    return 0;
}

int orc_distance_f32(void *h, const float *q, int64_t N, int32_t mode, float *d_c, float *profile, float *nearest, int threads) {
    return distance_impl<float>(*static_cast<Engine *>(h), q, N, mode, d_c, profile, nearest, threads);
}
int orc_distance_f64(void *h, const float *q, int64_t N, int32_t mode, double *d_c, double *profile, double *nearest, int threads) {
    return distance_impl<double>(*static_cast<Engine *>(h), q, N, mode, d_c, profile, nearest, threads);
}

int orc_underestimation_f32(void *h, const float *q, int64_t N, const float *planes, const int32_t *plane_idx, int32_t R,
                            float *d_c, float *profile, uint8_t *status) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    const size_t M = e->sc32.boxes.size();
    std::vector<orc::V3<float>> sk(n + 1);
    std::vector<float> prof(n);
    for (int64_t i = 0; i < N; i++) {
        const int r = plane_idx ? plane_idx[i] : 0;
        if (r < 0 || r >= R) return 1;
        orc::skeleton(e->robot, q + i * n, sk.data());
        bool rej;
        d_c[i] = orc::underestimation(e->robot, M, sk.data(), planes + size_t(r) * M * n * 6, cfg, prof.data(), &rej);
        if (status) status[i] = rej ? 1 : 0;
        if (profile) for (int k = 0; k < n; k++) profile[i * n + k] = rej ? 0.f : prof[k];
    }
    return 0;
}

int orc_extend_spines_f32(void *h, const float *q_root, const float *root_profile, float root_d_c, const float *root_planes, const float *q_targets,
                          int32_t S, const rvs_spine_params *sp, float *q_reached, float *d_c_reached, uint8_t *status) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    orc::SpineConfig sc;
    sc.num_layers = sp->num_layers; sc.num_iter_spine = sp->num_iter_spine; sc.use_expanded_bubble = sp->use_expanded_bubble != 0; sc.d_crit = sp->d_crit;
    for (int32_t s = 0; s < S; s++)
        status[s] = uint8_t(orc::extend_gen_spine(e->robot, e->sc32.boxes.size(), cfg, sc, q_root, root_profile, root_d_c, root_planes,
                                                  q_targets + size_t(s) * n, q_reached + size_t(s) * n, d_c_reached + s));
    return 0;
}

// spines with per-spine roots: root r = (q_roots[r], root_profiles[r], root_d_c[r], root_planes[r]); root_idx[S]
int orc_extend_spines_layers_f32(void *h, const float *q_roots, const float *root_profiles, const float *root_d_c, const float *root_planes,
                                 const int32_t *root_idx, const float *q_targets, int32_t S, const rvs_spine_params *sp, int32_t gen2,
                                 float *q_layers, int32_t *n_layers, float *d_c_last, uint8_t *status) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const size_t n = size_t(e->robot.n), M = e->sc32.boxes.size();
    orc::SpineConfig sc;
    sc.num_layers = sp->num_layers; sc.num_iter_spine = sp->num_iter_spine; sc.use_expanded_bubble = sp->use_expanded_bubble != 0; sc.d_crit = sp->d_crit;
    for (int32_t s = 0; s < S; s++) {
        const size_t r = size_t(root_idx[s]);
        int layers = 0;
        status[s] = uint8_t(orc::extend_gen_spine_layers(e->robot, M, cfg, sc, gen2 != 0, q_roots + r * n, root_profiles + r * n, root_d_c[r],
                                                         root_planes + r * M * n * 6, q_targets + size_t(s) * n,
                                                         q_layers + size_t(s) * size_t(sp->num_layers) * n, &layers, d_c_last + s));
        n_layers[s] = layers;
    }
    return 0;
}

int orc_enclosing_radii_f32(void *h, const float *q, int64_t N, float *radii) {
    auto *e = static_cast<Engine *>(h);
    const int n = e->robot.n;
    std::vector<orc::V3<float>> sk(n + 1);
    for (int64_t i = 0; i < N; i++) {
        orc::skeleton(e->robot, q + i * n, sk.data());
        orc::enclosing_radii(e->robot, sk.data(), radii + i * n * (n + 1));
    }
    return 0;
}

int orc_self_collision_capsules_f32(void *h, const float *q, int64_t N, uint8_t *mask) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    std::vector<orc::V3<float>> sk(n + 1);
    for (int64_t i = 0; i < N; i++) {
        orc::skeleton(e->robot, q + i * n, sk.data());
        mask[i] = orc::self_collision_mask(e->robot, sk.data(), cfg);
    }
    return 0;
}

// AbstractRobot::setSelfCollisionChecking (Scenario.cpp:101-103) + RRTConnectConfig::EPS_STEP (the marching step of xArm6.cpp:340-341)
int orc_set_self_collision(void *h, int32_t enabled, float eps_step) {
    auto *e = static_cast<Engine *>(h);
    e->robot.self_collision_checking = enabled != 0;
    e->robot.eps_step = eps_step;
    return 0;
}
// xArm6::checkSelfCollision(q1, q2&) (xArm6.cpp:284-368), capsule stage: hit[E], q2_out[E x n] (= q2, truncated on a hit)
int orc_self_collision_edges_f32(void *h, const float *q1, const float *q2, int64_t E, uint8_t *hit, float *q2_out) {
    auto *e = static_cast<Engine *>(h);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    for (int64_t i = 0; i < E; i++) {
        for (int k = 0; k < n; k++) q2_out[i * n + k] = q2[i * n + k];
        hit[i] = orc::self_collision_edge(e->robot, cfg, q1 + i * n, q2_out + i * n) ? 1 : 0;
    }
    return 0;
}

int orc_generate_configs(void *h, uint64_t seed, int64_t first, int64_t N, float *q) {
    auto *e = static_cast<Engine *>(h);
    const int n = e->robot.n;
    for (int64_t i = 0; i < N; i++) orc::generate_config(seed, first + i, n, e->lo.data(), e->hi.data(), q + i * n);
    return 0;
}

// ---- geometric primitives, batched over K cases. out rows: (d, has_points, P1xyz, P2xyz) ------------------
int orc_seg_seg_f32(int64_t K, const float *ABCD /*K x 12*/, float thr, float maxd, float *out /*K x 8*/) {
    const orc::Config<float> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) { const float *p = ABCD + 12 * i; put_near(orc::seg_seg(v3(p), v3(p + 3), v3(p + 6), v3(p + 9), cfg), out + 8 * i); }
    return 0;
}
int orc_seg_seg_f64(int64_t K, const double *ABCD, double thr, double maxd, double *out) {
    const orc::Config<double> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) { const double *p = ABCD + 12 * i; put_near(orc::seg_seg(v3(p), v3(p + 3), v3(p + 6), v3(p + 9), cfg), out + 8 * i); }
    return 0;
}
int orc_seg_point_f32(int64_t K, const float *ABC /*K x 9*/, float thr, float maxd, float *out) {
    const orc::Config<float> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) { const float *p = ABC + 9 * i; put_near(orc::seg_point(v3(p), v3(p + 3), v3(p + 6), cfg), out + 8 * i); }
    return 0;
}
int orc_capsule_box_distance_f32(int64_t K, const float *AB /*K x 6*/, const float *radius, const float *obs /*K x 6*/, float thr, float maxd, float *out) {
    const orc::Config<float> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) put_near(orc::capsule_box_distance(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], box6(obs + 6 * i), cfg), out + 8 * i);
    return 0;
}
int orc_capsule_box_distance_f64(int64_t K, const double *AB, const double *radius, const double *obs, double thr, double maxd, double *out) {
    const orc::Config<double> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) put_near(orc::capsule_box_distance(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], box6(obs + 6 * i), cfg), out + 8 * i);
    return 0;
}
int orc_capsule_box_collision_f32(int64_t K, const float *AB, const float *radius, const float *obs, float thr, float maxd, uint8_t *out) {
    const orc::Config<float> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) out[i] = orc::capsule_hits_box(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], box6(obs + 6 * i), cfg) ? 1 : 0;
    return 0;
}
int orc_capsule_rect_collision_f32(int64_t K, const float *AB, const float *radius, const float *obs, const int32_t *face, float thr, float maxd, uint8_t *out) {
    const orc::Config<float> cfg{thr, maxd, 0};
    for (int64_t i = 0; i < K; i++) out[i] = orc::capsule_hits_rectangle(v3(AB + 6 * i), v3(AB + 6 * i + 3), radius[i], box6(obs + 6 * i), face[i], cfg) ? 1 : 0;
    return 0;
}


// ---- dynamic environment (row f-2): Environment::updateEnvironment + MotionValidity::check ------------------------------------
// env_params[6] = (WS_center xyz, WS_radius, base_radius, robot_max_vel); state arrays are updated in place.
static std::vector<orc::EnvObject> env_objects(int M, const float *half, const float *pos, const float *vel, const float *max_vel, const int32_t *sign,
                                               const float *path_len, const uint8_t *dynamic) {
    std::vector<orc::EnvObject> objs(size_t(M > 0 ? M : 0));
    for (int j = 0; j < M; j++) {
        orc::EnvObject &o = objs[size_t(j)];
        o.half = {half[3 * j], half[3 * j + 1], half[3 * j + 2]};
        o.pos = {pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]};
        o.vel = {vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]};
        o.max_vel = max_vel[j]; o.sign = sign[j]; o.path_len = path_len[j]; o.dynamic = dynamic[j] != 0;
    }
    return objs;
}
static void env_store(const std::vector<orc::EnvObject> &objs, float *pos, float *vel, int32_t *sign, float *path_len) {
    for (size_t j = 0; j < objs.size(); j++) {
        const orc::EnvObject &o = objs[j];
        pos[3 * j] = o.pos.x; pos[3 * j + 1] = o.pos.y; pos[3 * j + 2] = o.pos.z;
        vel[3 * j] = o.vel.x; vel[3 * j + 1] = o.vel.y; vel[3 * j + 2] = o.vel.z;
        sign[j] = o.sign; path_len[j] = o.path_len;
    }
}
static orc::EnvParams env_params_of(const float *p, int motion_type) { return {motion_type, {p[0], p[1], p[2]}, p[3], p[4], p[5]}; }

int orc_env_update(const float *env_params, int32_t motion_type, int32_t M, const float *half, float *pos, float *vel, const float *max_vel, int32_t *sign,
                   float *path_len, const uint8_t *dynamic, float dt, int32_t steps, float *aabb_min, float *aabb_max) {
    if (motion_type == orc::MOTION_RANDOM_DIRECTIONS) return 5;
    const orc::EnvParams ep = env_params_of(env_params, motion_type);
    auto objs = env_objects(M, half, pos, vel, max_vel, sign, path_len, dynamic);
    for (int k = 0; k < steps; k++) orc::env_update(ep, objs, dt);
    env_store(objs, pos, vel, sign, path_len);
    if (aabb_min && aabb_max) for (int j = 0; j < M; j++) orc::env_aabb(objs[size_t(j)], aabb_min + 3 * j, aabb_max + 3 * j);
    return 0;
}

// Environment::moveRandomDirections (Environment.cpp:289-333): `steps` updates with the libc stream `rand_state` (36 words: LibcRand's table, position,
// draws; orc_libc_srand fills it).  acc is updated in place like pos and vel.
float orc_coord_norm(const float *a, const float *b, int32_t n) { return orc::coord_norm(a, b, n); }
void orc_libc_srand(uint32_t seed, uint32_t *rand_state) {
    orc::LibcRand g;
    g.seed(seed);
    std::copy_n(g.r, 34, rand_state); rand_state[34] = g.pos; rand_state[35] = g.draws;
}
uint32_t orc_libc_rand(uint32_t *rand_state) {
    orc::LibcRand g;
    std::copy_n(rand_state, 34, g.r); g.pos = rand_state[34]; g.draws = rand_state[35];
    const uint32_t v = g.next();
    std::copy_n(g.r, 34, rand_state); rand_state[34] = g.pos; rand_state[35] = g.draws;
    return v;
}
int orc_env_update_random(const float *env_params, int32_t ground_included, int32_t M, const float *half, float *pos, float *vel, float *acc, const float *max_vel,
                          const uint8_t *dynamic, float dt, int32_t steps, uint32_t *rand_state, float *aabb_min, float *aabb_max) {
    orc::EnvParams ep = env_params_of(env_params, orc::MOTION_RANDOM_DIRECTIONS);
    orc::LibcRand g;
    std::copy_n(rand_state, 34, g.r); g.pos = rand_state[34]; g.draws = rand_state[35];
    ep.ground_included = ground_included; ep.rng = &g;
    std::vector<int32_t> sign(size_t(M > 0 ? M : 0), 1);
    std::vector<float> path_len(size_t(M > 0 ? M : 0), 0.f);
    auto objs = env_objects(M, half, pos, vel, max_vel, sign.data(), path_len.data(), dynamic);
    for (int j = 0; j < M; j++) objs[size_t(j)].acc = {acc[3 * j], acc[3 * j + 1], acc[3 * j + 2]};
    for (int k = 0; k < steps; k++) orc::env_update(ep, objs, dt);
    env_store(objs, pos, vel, sign.data(), path_len.data());
    for (int j = 0; j < M; j++) { acc[3 * j] = objs[size_t(j)].acc.x; acc[3 * j + 1] = objs[size_t(j)].acc.y; acc[3 * j + 2] = objs[size_t(j)].acc.z; }
    if (aabb_min && aabb_max) for (int j = 0; j < M; j++) orc::env_aabb(objs[size_t(j)], aabb_min + 3 * j, aabb_max + 3 * j);
    std::copy_n(g.r, 34, rand_state); rand_state[34] = g.pos; rand_state[35] = g.draws;
    return 0;
}

// MotionValidity::check(q_previous, q_current) for E motions that share one obstacle timeline: validity point k (1..K) of every
// motion is tested against the environment after k updates.  first_invalid[e] = k of the first invalid point, 0 if the motion is
// valid.  commit_edge >= 0: the state arrays and the engine's obstacle table are left as the reference leaves them after checking
// THAT motion (updated first_invalid times, or K times); commit_edge < 0: nothing is changed.
// points != nullptr: MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-82) — validity point k of motion i is points[i][k-1][:]
static int motion_validity_impl(void *h, const float *env_params, int32_t motion_type, int32_t M, const float *half, float *pos, float *vel, const float *max_vel,
                                int32_t *sign, float *path_len, const uint8_t *dynamic, const float *q_prev, const float *q_cur, const float *points, int64_t E,
                                int32_t K, float dt, int32_t commit_edge, uint8_t *valid, int32_t *first_invalid) {
    auto *e = static_cast<Engine *>(h);
    if (motion_type == orc::MOTION_RANDOM_DIRECTIONS) return 5;
    if (int(e->sc32.boxes.size()) != M || commit_edge >= E) return 1;
    const orc::EnvParams ep = env_params_of(env_params, motion_type);
    auto objs = env_objects(M, half, pos, vel, max_vel, sign, path_len, dynamic);
    const auto cfg = config_of<float>(e->prm);
    const int n = e->robot.n;
    orc::Scene<float> sc = e->sc32;
    std::vector<orc::V3<float>> sk(size_t(n + 1));
    std::vector<float> qk(size_t(n), 0.f);
    std::vector<std::vector<orc::EnvObject>> timeline;
    for (int64_t i = 0; i < E; i++) { valid[i] = 1; first_invalid[i] = 0; }
    for (int k = 1; k <= K; k++) {
        orc::env_update(ep, objs, dt);
        timeline.push_back(objs);
        for (int j = 0; j < M; j++) { float mn[3], mx[3]; orc::env_aabb(objs[size_t(j)], mn, mx); sc.boxes[size_t(j)] = {{mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]}}; }
        for (int64_t i = 0; i < E; i++) {
            if (!valid[i]) continue;
            if (points) std::copy_n(points + (i * K + (k - 1)) * n, n, qk.begin());
            else orc::motion_point(q_prev + i * n, q_cur + i * n, n, k, K, qk.data());
            orc::skeleton(e->robot, qk.data(), sk.data());
            if (!orc::is_valid(e->robot, sc, sk.data(), cfg)) { valid[i] = 0; first_invalid[i] = k; }
        }
    }
    if (commit_edge >= 0 && K > 0) {
        const int kc = first_invalid[commit_edge] ? first_invalid[commit_edge] : K;
        const auto &st = timeline[size_t(kc - 1)];
        env_store(st, pos, vel, sign, path_len);
        for (int j = 0; j < M; j++) {
            float mn[3], mx[3];
            orc::env_aabb(st[size_t(j)], mn, mx);
            e->sc32.boxes[size_t(j)] = {{mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]}};
            for (int c = 0; c < 3; c++) { e->sc64.boxes[size_t(j)].o[c] = double(mn[c]); e->sc64.boxes[size_t(j)].o[3 + c] = double(mx[c]); }
        }
    }
    return 0;
}
int orc_motion_validity(void *h, const float *env_params, int32_t motion_type, int32_t M, const float *half, float *pos, float *vel, const float *max_vel,
                        int32_t *sign, float *path_len, const uint8_t *dynamic, const float *q_prev, const float *q_cur, int64_t E, int32_t K, float dt,
                        int32_t commit_edge, uint8_t *valid, int32_t *first_invalid) {
    return motion_validity_impl(h, env_params, motion_type, M, half, pos, vel, max_vel, sign, path_len, dynamic, q_prev, q_cur, nullptr, E, K, dt, commit_edge, valid, first_invalid);
}
int orc_motion_validity_points(void *h, const float *env_params, int32_t motion_type, int32_t M, const float *half, float *pos, float *vel, const float *max_vel,
                               int32_t *sign, float *path_len, const uint8_t *dynamic, const float *points, int64_t E, int32_t K, float dt,
                               int32_t commit_edge, uint8_t *valid, int32_t *first_invalid) {
    return motion_validity_impl(h, env_params, motion_type, M, half, pos, vel, max_vel, sign, path_len, dynamic, nullptr, nullptr, points, E, K, dt, commit_edge, valid, first_invalid);
}
}  // extern "C"
