// rvs_ref_capi.cpp — C entry points around the REFERENCE's own classes, compiled from /root/reference unmodified and where they lie,
// against oracle/eigen_shim and oracle/env_shim.  TEST INFRASTRUCTURE ONLY (oracle/_ref/librvs_ref.so, git-ignored):
//   * base::RealVectorSpace (isValid(q), isValid(q1,q2), computeDistance, computeDistanceUnderestimation) from RealVectorSpace.cpp
//     (+ State.cpp, StateSpace.cpp, RealVectorSpaceState.cpp, CollisionAndDistance.cpp, AbstractRobot.cpp, Environment.cpp, Object.cpp,
//     Box.cpp): pins the LOOPS of the hot path — link / obstacle order, the ground rule, min_dist_tol, the first-collision exit and
//     its profile fill, nearest-point storage, the edge discretisation (k = N..1, float step), the underestimation's reject rule and
//     state side effects — on top of the already pinned capsule / box primitives;
//   * planning::rbt::RBTConnect::extendSpine and RGBTConnect::extendGenSpine / extendGenSpine2 from RBTConnect.cpp / RGBTConnect.cpp
//     (+ RRTConnect, AbstractPlanner, PlannerInfo, PlanningTypes, Tree, the *ConnectConfig sources; nanoflann is an exhaustive
//     exact-nearest-neighbour stand-in): pins the bur spines (row f-1);
//   * RRTConnect / RBTConnect / RGBTConnect::solve() themselves, one run per srand(seed): pins the planners of
//     rpmplv2_b200/cpp/planning run by run (rows f-3 / f-4);
//   * planning::trajectory::MotionValidity::check(q_previous, q_current) from MotionValidity.cpp on the reference's own
//     Environment::updateEnvironment: pins the dynamic-obstacle motion check (row f-2).
// Built a second time with -DRVS_DROPIN (oracle/_ref/libdropin_ref.so, linked against rpmplv2_b200/librvs_cuda.so): the harness then
// constructs base::RealVectorSpaceCUDA (rpmplv2_b200/cpp, compiled here against the reference's REAL headers) instead of
// base::RealVectorSpace, and the reference's own planners / loops run over the CUDA engine through base::StateSpace's virtual
// interface — the drop-in itself, checked on the GPU box against the golden vectors of the CPU build (tests/test_dropin.py).
// What it cannot pin: the robot's forward kinematics and enclosing radii (KDL) — the harness robot answers computeSkeleton and
// computeEnclosingRadii with the oracle's own restatements, so both sides see identical skeletons and radii.
#include <cstdint>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../include/rvs_cuda.h"
#include "rvs_oracle.hpp"

#define private public                                         // env::Environment keeps WS_center / WS_radius / motion_type / sign / path_len private
#include "Environment.h"
#undef private
#include "RealVectorSpace.h"
#ifdef RVS_DROPIN                                              // the same harness with the CUDA state space under the reference's planners
#include "RealVectorSpaceCUDA.h"
#endif
#include "MotionValidity.h"
#define protected public                                       // extendSpine / extendGenSpine / extendGenSpine2 are protected members
#include "RGBTConnect.h"
#include "PatternTree.h"
#undef protected

namespace {

class HarnessRobot : public robots::AbstractRobot {
public:
    orc::Robot chain;
    explicit HarnessRobot(const rvs_robot_desc *rd) {
        chain.kind = rd->kind; chain.n = rd->num_dofs; chain.ground_included = rd->ground_included; chain.gripper_length = rd->gripper_length;
        for (int k = 0; k <= chain.n; k++) {
            chain.fixed_rot.push_back(orc::rpy_matrix(rd->joint_rpy[k][0], rd->joint_rpy[k][1], rd->joint_rpy[k][2]));
            chain.xyz.push_back({rd->joint_xyz[k][0], rd->joint_xyz[k][1], rd->joint_xyz[k][2]});
        }
        for (int k = 0; k < chain.n; k++) {
            const double *a = rd->joint_axis[k];
            const double len = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            chain.axis.push_back({a[0] / len, a[1] / len, a[2] / len});
            chain.radius.push_back(rd->capsule_radius[k]);
            capsules_radius.push_back(rd->capsule_radius[k]);
            limits.emplace_back(rd->limit_lo[k], rd->limit_hi[k]);
        }
        type = "harness";
        num_DOFs = size_t(chain.n);
        links.resize(size_t(chain.n));                       // getNumLinks() == links.size()
        ground_included = size_t(rd->ground_included);
        gripper_length = rd->gripper_length;
        self_collision_checking = false;
    }
    void setState(const std::shared_ptr<base::State>) override {}
    std::shared_ptr<std::vector<KDL::Frame>> computeForwardKinematics(const std::shared_ptr<base::State>) override { return nullptr; }
    std::shared_ptr<base::State> computeInverseKinematics(const KDL::Rotation &, const KDL::Vector &, const std::shared_ptr<base::State>) override { return nullptr; }
    std::shared_ptr<Eigen::MatrixXf> computeSkeleton(const std::shared_ptr<base::State> &q) override {
        std::vector<orc::V3<float>> sk(size_t(chain.n + 1));
        std::vector<float> qq(size_t(chain.n));
        for (int i = 0; i < chain.n; i++) qq[size_t(i)] = q->getCoord(i);
        orc::skeleton(chain, qq.data(), sk.data());
        auto m = std::make_shared<Eigen::MatrixXf>(3, chain.n + 1);
        for (int k = 0; k <= chain.n; k++) { (*m)(0, k) = sk[size_t(k)].x; (*m)(1, k) = sk[size_t(k)].y; (*m)(2, k) = sk[size_t(k)].z; }
        return m;
    }
    // the robot classes need KDL: the harness answers with the oracle's restatement of computeEnclosingRadii on the same skeleton
    std::shared_ptr<Eigen::MatrixXf> computeEnclosingRadii(const std::shared_ptr<base::State> &q) override {
        std::vector<orc::V3<float>> sk(size_t(chain.n + 1));
        std::vector<float> qq(size_t(chain.n)), R(size_t(chain.n) * size_t(chain.n + 1));
        for (int i = 0; i < chain.n; i++) qq[size_t(i)] = q->getCoord(i);
        orc::skeleton(chain, qq.data(), sk.data());
        orc::enclosing_radii(chain, sk.data(), R.data());
        auto m = std::make_shared<Eigen::MatrixXf>(chain.n, chain.n + 1);
        for (int i = 0; i < chain.n; i++)
            for (int j = 0; j <= chain.n; j++) (*m)(i, j) = R[size_t(i) * size_t(chain.n + 1) + size_t(j)];
        return m;
    }
    bool checkSelfCollision(const std::shared_ptr<base::State> &, std::shared_ptr<base::State> &) override { return false; }
    bool checkSelfCollision(const std::shared_ptr<base::State> &) override { return false; }
};

struct Harness {
    std::shared_ptr<HarnessRobot> robot;
    std::shared_ptr<env::Environment> env;
    std::shared_ptr<base::RealVectorSpace> ss;
    int n = 0;
};

std::shared_ptr<base::State> state_of(Harness &h, const float *q) {
    Eigen::VectorXf c(h.n);
    for (int i = 0; i < h.n; i++) c(i) = q[i];
    return h.ss->getNewState(c);
}

}  // namespace

extern "C" {

void *rvsref_create(const rvs_robot_desc *rd, float eq_thr, float max_dist, int32_t num_checks) {
    RealVectorSpaceConfig::EQUALITY_THRESHOLD = eq_thr;
    RealVectorSpaceConfig::MAX_DISTANCE = max_dist;
    RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS = size_t(num_checks);
    auto *h = new Harness();
    h->n = rd->num_dofs;
    h->robot = std::make_shared<HarnessRobot>(rd);
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());                              // the YAML constructor reports the (stubbed, empty) file
    h->env = std::make_shared<env::Environment>(std::string("unused.yaml"), std::string(""));
    std::cout.rdbuf(keep);
#ifdef RVS_DROPIN
    try {
        h->ss = std::make_shared<base::RealVectorSpaceCUDA>(size_t(h->n), h->robot, h->env, *rd);
    } catch (const std::exception &e) {
        std::cerr << "rvsref_create (drop-in): " << e.what() << std::endl;
        delete h;
        return nullptr;
    }
#else
    h->ss = std::make_shared<base::RealVectorSpace>(size_t(h->n), h->robot, h->env);
#endif
    return h;
}
void rvsref_destroy(void *p) { delete static_cast<Harness *>(p); }

int rvsref_set_obstacles(void *p, const float *mn, const float *mx, const uint8_t *ground, const float *tol, int32_t M) {
    auto *h = static_cast<Harness *>(p);
    h->env->removeAllObjects();
    for (int j = 0; j < M; j++) {
        const fcl::Vector3f lo(mn[3 * j], mn[3 * j + 1], mn[3 * j + 2]), hi(mx[3 * j], mx[3 * j + 1], mx[3 * j + 2]);
        auto obj = std::make_shared<env::Box>(fcl::Vector3f(hi - lo), fcl::Vector3f(0.5f * (lo + hi)), fcl::Quaternionf::Identity(), (ground && ground[j]) ? "ground" : "");
        obj->getCollObject()->stubSetAABB(lo, hi);              // the table's exact values (what FCL's computeAABB would report is not pinned here)
        if (tol) obj->setMinDistTol(tol[j]);
        h->env->addObject(obj);
    }
#ifdef RVS_DROPIN
    static_cast<base::RealVectorSpaceCUDA *>(h->ss.get())->refreshObstacles();
#endif
    return 0;
}

int rvsref_is_valid(void *p, const float *q, int64_t N, uint8_t *out) {
    auto *h = static_cast<Harness *>(p);
    for (int64_t i = 0; i < N; i++) out[i] = h->ss->isValid(state_of(*h, q + i * h->n)) ? 1 : 0;
    return 0;
}
int rvsref_is_valid_edges(void *p, const float *q1, const float *q2, int64_t E, uint8_t *out) {
    auto *h = static_cast<Harness *>(p);
    for (int64_t i = 0; i < E; i++) out[i] = h->ss->isValid(state_of(*h, q1 + i * h->n), state_of(*h, q2 + i * h->n)) ? 1 : 0;
    return 0;
}
// d_c[N], profile[N x L], nearest[N x M x L x 6] (zeros where the State holds nullptr), has_nearest[N], is_real[N]
int rvsref_distance(void *p, const float *q, int64_t N, float *d_c, float *profile, float *nearest, uint8_t *has_nearest, uint8_t *is_real) {
    auto *h = static_cast<Harness *>(p);
    const size_t L = size_t(h->n), M = h->env->getNumObjects();
    for (int64_t i = 0; i < N; i++) {
        auto s = state_of(*h, q + i * h->n);
        d_c[i] = h->ss->computeDistance(s, false);
        if (s->getDistance() != d_c[i]) return 2;
        for (size_t l = 0; l < L; l++) profile[size_t(i) * L + l] = s->getDistanceProfile()[l];
        const auto np = s->getNearestPoints();
        has_nearest[i] = np ? 1 : 0;
        is_real[i] = s->getIsRealDistance() ? 1 : 0;
        if (nearest)
            for (size_t j = 0; j < M; j++)
                for (size_t l = 0; l < L; l++)
                    for (int k = 0; k < 6; k++) nearest[((size_t(i) * M + j) * L + l) * 6 + k] = np ? (*np)[j](k, Eigen::Index(l)) : 0.f;
    }
    return 0;
}
// planes: M x L x 6 of ONE root.  status[i] = 1 when the reference returned 0 and left the state untouched (d_c stays -1).
int rvsref_underestimation(void *p, const float *q, int64_t N, const float *planes, float *d_c, float *profile, uint8_t *status, uint8_t *is_real) {
    auto *h = static_cast<Harness *>(p);
    const size_t L = size_t(h->n), M = h->env->getNumObjects();
    auto np = std::make_shared<std::vector<Eigen::MatrixXf>>(M, Eigen::MatrixXf(6, Eigen::Index(L)));
    for (size_t j = 0; j < M; j++)
        for (size_t l = 0; l < L; l++)
            for (int k = 0; k < 6; k++) (*np)[j](k, Eigen::Index(l)) = planes[(j * L + l) * 6 + size_t(k)];
    for (int64_t i = 0; i < N; i++) {
        auto s = state_of(*h, q + i * h->n);
        d_c[i] = h->ss->computeDistanceUnderestimation(s, np);
        const bool untouched = s->getDistance() == -1;
        status[i] = untouched ? 1 : 0;
        is_real[i] = s->getIsRealDistance() ? 1 : 0;
        for (size_t l = 0; l < L; l++) profile[size_t(i) * L + l] = untouched ? 0.f : s->getDistanceProfile()[l];
    }
    return 0;
}


// Bur spines (row f-1) through the reference's own planning::rbt::RBTConnect::extendSpine (RBTConnect.cpp:85-150) and
// RGBTConnect::extendGenSpine / extendGenSpine2 (RGBTConnect.cpp:73-118, use_real_dist = false), compiled unmodified.
// mode 0: extendSpine (one layer), 1: extendGenSpine (only the final state, written as layer 0), 2: extendGenSpine2 (every layer).
// Every spine starts from a FRESH root state at q_roots[s] (computeDistance fills its profile and planes, as the planners do).
// q_layers[S x num_layers x n], n_layers[S], d_last[S] (distance cached in the last state: underestimation or -1), status[S].
int rvsref_extend_spines(void *p, const float *q_roots, const float *q_targets, int32_t S, int32_t num_layers, int32_t num_iter_spine,
                         int32_t use_expanded_bubble, float d_crit, int32_t mode, float *q_layers, int32_t *n_layers, float *d_last, uint8_t *status) {
    auto *h = static_cast<Harness *>(p);
    RGBTConnectConfig::NUM_LAYERS = size_t(num_layers);
    RBTConnectConfig::NUM_ITER_SPINE = size_t(num_iter_spine);
    RBTConnectConfig::USE_EXPANDED_BUBBLE = use_expanded_bubble != 0;
    RBTConnectConfig::D_CRIT = d_crit;
    planning::rbt::RGBTConnect planner(h->ss);
    const size_t n = size_t(h->n);
    for (int32_t s = 0; s < S; s++) {
        auto q = state_of(*h, q_roots + size_t(s) * n), q_e = state_of(*h, q_targets + size_t(s) * n);
        h->ss->computeDistance(q, false);
        std::vector<std::shared_ptr<base::State>> made;
        base::State::Status st;
        if (mode == 2) {
            if (!(q->getDistance() > 0)) { n_layers[s] = 0; d_last[s] = q->getDistance(); status[s] = uint8_t(base::State::Status::Trapped); continue; }
            std::shared_ptr<std::vector<std::shared_ptr<base::State>>> list;
            std::tie(st, list) = planner.extendGenSpine2(q, q_e, false);
            made = *list;
        } else {
            std::shared_ptr<base::State> q_new;
            if (mode == 1) std::tie(st, q_new) = planner.extendGenSpine(q, q_e, false);
            else std::tie(st, q_new) = planner.extendSpine(q, q_e);
            made.push_back(q_new);
        }
        n_layers[s] = int32_t(made.size());
        status[s] = uint8_t(st);
        d_last[s] = made.back()->getDistance();
        for (size_t l = 0; l < made.size(); l++)
            for (size_t i = 0; i < n; i++) q_layers[(size_t(s) * size_t(num_layers) + l) * n + i] = made[l]->getCoord(int(i));
    }
    return 0;
}


// Row f-2 through the reference's own planning::trajectory::MotionValidity::check(q_previous, q_current) (MotionValidity.cpp:30-51),
// compiled unmodified: T consecutive motions on ONE environment, as a DRGBT run makes them (every check advances the obstacles by the
// updates it made before returning).  The environment's private parameters are set directly (the YAML reader is a stand-in).
// Outputs: valid[T], passed[T] (validity points appended to `path` by that check), pos_after / vel_after [T x M x 3].
static void motion_env_setup(Harness *h, const float *env_params, int32_t motion_type, int32_t M, const float *dim, const float *pos, const float *vel,
                             const float *max_vel, const uint8_t *dynamic, const uint8_t *ground, const float *tol) {
    auto &e = *h->env;
    e.removeAllObjects();
    e.WS_center = fcl::Vector3f(env_params[0], env_params[1], env_params[2]);
    e.WS_radius = env_params[3];
    e.setBaseRadius(env_params[4]);
    e.setRobotMaxVel(env_params[5]);
    e.motion_type = static_cast<decltype(e.motion_type)>(motion_type);
    e.sign = Eigen::VectorXi::Ones(M);
    e.path_len = Eigen::VectorXf::Zero(M);
    for (int j = 0; j < M; j++) {
        auto obj = std::make_shared<env::Box>(fcl::Vector3f(dim[3 * j], dim[3 * j + 1], dim[3 * j + 2]), fcl::Vector3f(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]),
                                              fcl::Quaternionf::Identity(), ground[j] ? "ground" : (dynamic[j] ? "dynamic_obstacle" : ""));
        obj->setMaxVel(max_vel[j]);
        obj->setMinDistTol(tol[j]);
        e.addObject(obj, fcl::Vector3f(vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]));
    }
}
static void motion_env_read(Harness *h, int32_t M, float *pos_after, float *vel_after) {
    for (int j = 0; j < M; j++)
        for (int c = 0; c < 3; c++) {
            pos_after[size_t(j) * 3 + size_t(c)] = h->env->getObject(size_t(j))->getPosition()(c);
            vel_after[size_t(j) * 3 + size_t(c)] = h->env->getObject(size_t(j))->getVelocity()(c);
        }
}

int rvsref_motion_run(void *p, const float *env_params /*WS_center xyz, WS_radius, base_radius, robot_max_vel*/, int32_t motion_type, int32_t M,
                      const float *dim, const float *pos, const float *vel, const float *max_vel, const uint8_t *dynamic, const uint8_t *ground,
                      const float *tol, const float *q_prev, const float *q_cur, int32_t T, int32_t K, float max_iter_time, uint8_t *valid,
                      int32_t *passed, float *pos_after, float *vel_after) {
    auto *h = static_cast<Harness *>(p);
    motion_env_setup(h, env_params, motion_type, M, dim, pos, vel, max_vel, dynamic, ground, tol);
    std::vector<std::shared_ptr<base::State>> path;
    planning::trajectory::MotionValidity mv(h->ss, 0.01f, max_iter_time, &path);
    TrajectoryConfig::NUM_VALIDITY_POINTS_CHECK = size_t(K);       // the constructor derives it from the obstacles' max_vel; the test fixes it
    for (int32_t t = 0; t < T; t++) {
        const size_t before = path.size();
        valid[t] = mv.check(state_of(*h, q_prev + size_t(t) * size_t(h->n)), state_of(*h, q_cur + size_t(t) * size_t(h->n))) ? 1 : 0;
        passed[t] = int32_t(path.size() - before);
        motion_env_read(h, M, pos_after + size_t(t) * size_t(M) * 3, vel_after + size_t(t) * size_t(M) * 3);
    }
    return 0;
}

// The other overload, MotionValidity::check(traj_points_current_iter) (MotionValidity.cpp:60-82): T consecutive checks of K explicit
// trajectory points each (points: T x K x n), delta_time = max_iter_time / NUM_VALIDITY_POINTS_CHECK with that static set to K.
int rvsref_motion_points_run(void *p, const float *env_params, int32_t motion_type, int32_t M, const float *dim, const float *pos, const float *vel,
                             const float *max_vel, const uint8_t *dynamic, const uint8_t *ground, const float *tol, const float *points, int32_t T,
                             int32_t K, float max_iter_time, uint8_t *valid, int32_t *passed, float *pos_after, float *vel_after) {
    auto *h = static_cast<Harness *>(p);
    motion_env_setup(h, env_params, motion_type, M, dim, pos, vel, max_vel, dynamic, ground, tol);
    std::vector<std::shared_ptr<base::State>> path;
    planning::trajectory::MotionValidity mv(h->ss, 0.01f, max_iter_time, &path);
    TrajectoryConfig::NUM_VALIDITY_POINTS_CHECK = size_t(K);
    const size_t n = size_t(h->n);
    for (int32_t t = 0; t < T; t++) {
        std::vector<Eigen::VectorXf> traj;
        for (int32_t k = 0; k < K; k++) {
            Eigen::VectorXf v(h->n);
            for (size_t i = 0; i < n; i++) v(int(i)) = points[(size_t(t) * size_t(K) + size_t(k)) * n + i];
            traj.push_back(v);
        }
        const size_t before = path.size();
        valid[t] = mv.check(traj) ? 1 : 0;
        passed[t] = int32_t(path.size() - before);
        motion_env_read(h, M, pos_after + size_t(t) * size_t(M) * 3, vel_after + size_t(t) * size_t(M) * 3);
    }
    return 0;
}


// Row f-3 through the reference's own planners: ONE run of RRTConnect / RBTConnect / RGBTConnect::solve() (kind 0 / 1 / 2) on the
// harness scene after srand(seed) — Eigen::VectorXf::Random (here: the work-alike's, one std::rand() per coefficient) is the planners'
// only source of randomness, so a planner that draws from the same stream can be compared state by state.  base::Tree's kd-tree is the
// exhaustive stand-in of env_shim/nanoflann.hpp (exact nearest neighbour, like nanoflann).
// Returns 0, or 2 when the constructor throws (start / goal invalid).  path: path_len x n (truncated to max_path rows).
int rvsref_plan(void *p, int32_t kind, const float *q_start, const float *q_goal, uint32_t seed, int64_t max_num_iter, int64_t max_num_states,
                float max_planning_time, float eps_step, int32_t max_extension_steps, float delta, int32_t num_spines, int32_t num_layers,
                int32_t num_iter_spine, int32_t use_expanded_bubble, float d_crit, int32_t *success, int64_t *iterations, int64_t *states,
                int32_t *path_len, float *path, int32_t max_path, int64_t *tree_sizes) {
    auto *h = static_cast<Harness *>(p);
    RRTConnectConfig::MAX_NUM_ITER = RBTConnectConfig::MAX_NUM_ITER = RGBTConnectConfig::MAX_NUM_ITER = size_t(max_num_iter);
    RRTConnectConfig::MAX_NUM_STATES = RBTConnectConfig::MAX_NUM_STATES = RGBTConnectConfig::MAX_NUM_STATES = size_t(max_num_states);
    RRTConnectConfig::MAX_PLANNING_TIME = RBTConnectConfig::MAX_PLANNING_TIME = RGBTConnectConfig::MAX_PLANNING_TIME = max_planning_time;
    RRTConnectConfig::EPS_STEP = eps_step;
    RRTConnectConfig::MAX_EXTENSION_STEPS = size_t(max_extension_steps);
    RBTConnectConfig::DELTA = delta;
    RBTConnectConfig::NUM_SPINES = size_t(num_spines);
    RBTConnectConfig::NUM_ITER_SPINE = size_t(num_iter_spine);
    RBTConnectConfig::USE_EXPANDED_BUBBLE = use_expanded_bubble != 0;
    RBTConnectConfig::D_CRIT = d_crit;
    RGBTConnectConfig::NUM_LAYERS = size_t(num_layers);
    std::srand(seed);
    std::unique_ptr<planning::rrt::RRTConnect> planner;
    try {
        auto qs = state_of(*h, q_start), qg = state_of(*h, q_goal);
        if (kind == 0) planner = std::make_unique<planning::rrt::RRTConnect>(h->ss, qs, qg);
        else if (kind == 1) planner = std::make_unique<planning::rbt::RBTConnect>(h->ss, qs, qg);
        else planner = std::make_unique<planning::rbt::RGBTConnect>(h->ss, qs, qg);
    } catch (const std::domain_error &) {
        return 2;
    }
    *success = planner->solve() ? 1 : 0;
    *iterations = int64_t(planner->getPlannerInfo()->getNumIterations());
    *states = int64_t(planner->getPlannerInfo()->getNumStates());
    const auto &pth = planner->getPath();
    *path_len = *success ? int32_t(pth.size()) : 0;
    for (int32_t k = 0; k < *path_len && k < max_path; k++)
        for (int i = 0; i < h->n; i++) path[size_t(k) * size_t(h->n) + size_t(i)] = pth[size_t(k)]->getCoord(i);
    tree_sizes[0] = int64_t(planner->trees[0]->getNumStates());      // (getTree() returns a copy whose destructor clears the shared state list)
    tree_sizes[1] = int64_t(planner->trees[1]->getNumStates());
    return 0;
}


// planning::rbt::PatternTree::generateLocalTree (PatternTree.cpp:62-91), compiled unmodified: the local tree of generalized burs around
// q_root.  Outputs (node order = the reference's base::Tree order): coord[count x n], parent[count] (-1 for the root), d_c[count],
// profile[count x n] (NaN where the State holds no profile), is_real[count].  Returns the number of nodes (<= max_nodes), or -1.
int64_t rvsref_pattern_tree(void *p, const float *q_root, int32_t tree_layers, int32_t num_layers, int32_t num_iter_spine, int32_t use_expanded_bubble,
                            float d_crit, float delta, int64_t max_nodes, float *coord, int32_t *parent, float *d_c, float *profile, uint8_t *is_real) {
    auto *h = static_cast<Harness *>(p);
    RGBTConnectConfig::NUM_LAYERS = size_t(num_layers);
    RBTConnectConfig::NUM_ITER_SPINE = size_t(num_iter_spine);
    RBTConnectConfig::USE_EXPANDED_BUBBLE = use_expanded_bubble != 0;
    RBTConnectConfig::D_CRIT = d_crit;
    RBTConnectConfig::DELTA = delta;
    planning::rbt::PatternTree pattern(h->ss, size_t(tree_layers));
    const auto tree = pattern.generateLocalTree(state_of(*h, q_root));
    const int64_t count = int64_t(tree->getNumStates());
    if (count > max_nodes) return -1;
    const size_t n = size_t(h->n);
    for (int64_t k = 0; k < count; k++) {
        const auto q = tree->getState(size_t(k));
        for (size_t i = 0; i < n; i++) coord[size_t(k) * n + i] = q->getCoord(int(i));
        parent[k] = q->getParent() ? int32_t(q->getParent()->getIdx()) : -1;
        d_c[k] = q->getDistance();
        is_real[k] = q->getIsRealDistance() ? 1 : 0;
        const auto &prof = q->getDistanceProfile();
        for (size_t i = 0; i < n; i++) profile[size_t(k) * n + i] = prof.size() == n ? prof[i] : NAN;
    }
    return count;
}

}  // extern "C"
