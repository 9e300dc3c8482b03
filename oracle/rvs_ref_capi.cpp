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
// Round 2: the ROBOT classes are the reference's own too.  oracle/env_shim now carries work-alikes of orocos-KDL, kdl_parser, urdfdom
// and glog (third-party, not in /root/reference), so src/robots/{xArm6,Planar2DOF,Planar10DOF,Spatial10DOF,Spatial14DOF,Spatial18DOF,
// Spatial22DOF}.cpp compile unmodified and rvsref_create_robot() builds the very class scenario::Scenario would (Scenario.cpp:13-51)
// from the reference's own URDF.  computeSkeleton (xArm6.cpp:209-235, Planar2DOF.cpp:117-130), computeEnclosingRadii
// (xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30), computeCapsulesDistance and both checkSelfCollision overloads
// (xArm6.cpp:284-424; the FCL mesh stage answers "collide", see env_shim/fcl/fcl.h) are therefore the reference's code on every golden
// path; only KDL's frame arithmetic underneath is a restatement of a third-party library (env_shim/kdl/kdl_shim.hpp).
// rvsref_create() keeps the descriptor-only HarnessRobot (oracle FK + oracle radii) for chains that have no reference class.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../include/rvs_cuda.h"
#include "rvs_oracle.hpp"

#define private public                                         // env::Environment keeps WS_center / WS_radius / motion_type / sign / path_len private,
#include "Environment.h"                                       // xArm6 (pulled in by RealVectorSpace.h) its computeCapsulesDistance
#include "RealVectorSpace.h"
#include "xArm6.h"
#undef private
#ifdef RVS_DROPIN                                              // the same harness with the CUDA state space under the reference's planners
#include "RealVectorSpaceCUDA.h"
#endif
#include "MotionValidity.h"
#include "Planar10DOF.h"
#include "Spatial14DOF.h"
#include "Spatial18DOF.h"
#include "Spatial22DOF.h"
#include <queue>                                               // RRTx.h relies on these arriving through its third-party includes
#include <random>
#include <unordered_set>
#define protected public                                       // extendSpine / extendGenSpine / extendGenSpine2 are protected members,
#include "RGBTConnect.h"                                       // and so are RRTx's findNeighbors / chooseParent / rewireNeighbors / updatePath
#include "PatternTree.h"
#include "RRTx.h"
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
    void setType(const std::string &t) { type = t; }
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
    // Where the reference's own robot class cannot be built (no /root/reference on the GPU box, generic test chains) the harness robot
    // answers self-collision with the oracle's restatement of xArm6.cpp:284-424 — checked bit for bit against the compiled xArm6.cpp in
    // this container (tests/golden/make_robot_golden.py); every other kind answers false like the reference's stubs.
    bool checkSelfCollision(const std::shared_ptr<base::State> &q1, std::shared_ptr<base::State> &q2) override {
        if (!self_collision_checking || chain.kind != orc::XARM6) return false;
        chain.self_collision_checking = true;
        chain.eps_step = RRTConnectConfig::EPS_STEP;
        const orc::Config<float> cfg{RealVectorSpaceConfig::EQUALITY_THRESHOLD, RealVectorSpaceConfig::MAX_DISTANCE, int(RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS)};
        std::vector<float> a(size_t(chain.n)), b(size_t(chain.n));
        for (int i = 0; i < chain.n; i++) { a[size_t(i)] = q1->getCoord(i); b[size_t(i)] = q2->getCoord(i); }
        if (!orc::self_collision_edge(chain, cfg, a.data(), b.data())) return false;
        Eigen::VectorXf c(chain.n);
        for (int i = 0; i < chain.n; i++) c(i) = b[size_t(i)];
        q2 = std::make_shared<base::RealVectorSpaceState>(c);
        return true;
    }
    bool checkSelfCollision(const std::shared_ptr<base::State> &q) override {
        if (!self_collision_checking || chain.kind != orc::XARM6) return false;
        const orc::Config<float> cfg{RealVectorSpaceConfig::EQUALITY_THRESHOLD, RealVectorSpaceConfig::MAX_DISTANCE, int(RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS)};
        std::vector<orc::V3<float>> sk(size_t(chain.n + 1));
        std::vector<float> qq(size_t(chain.n));
        for (int i = 0; i < chain.n; i++) qq[size_t(i)] = q->getCoord(i);
        orc::skeleton(chain, qq.data(), sk.data());
        return orc::self_collision_mask(chain, sk.data(), cfg) != 0;
    }
};

struct Harness {
    std::shared_ptr<robots::AbstractRobot> robot;
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

static void *harness_finish(Harness *h, const rvs_robot_desc *rd, float eq_thr, float max_dist, int32_t num_checks) {
    RealVectorSpaceConfig::EQUALITY_THRESHOLD = eq_thr;
    RealVectorSpaceConfig::MAX_DISTANCE = max_dist;
    RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS = size_t(num_checks);
    h->n = rd->num_dofs;
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

void *rvsref_create(const rvs_robot_desc *rd, float eq_thr, float max_dist, int32_t num_checks) {
    auto *h = new Harness();
    h->robot = std::make_shared<HarnessRobot>(rd);
    return harness_finish(h, rd, eq_thr, max_dist, num_checks);
}

// The reference's OWN robot class, constructed the way scenario::Scenario does (Scenario.cpp:13-51) from the reference's own URDF:
// type = "xarm6" | "planar_2dof" | "planar_10dof" | "spatial_10dof" | "spatial_14dof" | "spatial_18dof" | "spatial_22dof".
// gripper_length / ground_included / capsule radii come from `rd` (the scenario YAML's robot node); self_collision overrides the
// constructor's default (Scenario.cpp:101-103).  Returns nullptr when the class cannot be built (URDF missing, chain mismatch).
void *rvsref_create_robot(const rvs_robot_desc *rd, const char *type_c, const char *urdf, int32_t self_collision, float eq_thr, float max_dist,
                          int32_t num_checks) {
    const std::string type(type_c);
    std::shared_ptr<robots::AbstractRobot> robot;
    try {
        if (type == "xarm6") robot = std::make_shared<robots::xArm6>(urdf, rd->gripper_length, size_t(rd->ground_included));
        else if (type == "planar_2dof") robot = std::make_shared<robots::Planar2DOF>(urdf);
        else if (type == "planar_10dof") robot = std::make_shared<robots::Planar10DOF>(urdf);
        else if (type == "spatial_10dof") robot = std::make_shared<robots::Spatial10DOF>(urdf, size_t(rd->ground_included));
        else if (type == "spatial_14dof") robot = std::make_shared<robots::Spatial14DOF>(urdf, size_t(rd->ground_included));
        else if (type == "spatial_18dof") robot = std::make_shared<robots::Spatial18DOF>(urdf, size_t(rd->ground_included));
        else if (type == "spatial_22dof") robot = std::make_shared<robots::Spatial22DOF>(urdf, size_t(rd->ground_included));
        else return nullptr;
    } catch (const std::exception &e) {
        std::cerr << "rvsref_create_robot: " << e.what() << std::endl;
        return nullptr;
    }
    if (int(robot->getNumDOFs()) != rd->num_dofs) return nullptr;
    robot->setCapsulesRadius(std::vector<float>(rd->capsule_radius, rd->capsule_radius + rd->num_dofs));
    robot->setSelfCollisionChecking(self_collision != 0);
    auto *h = new Harness();
    h->robot = robot;
    return harness_finish(h, rd, eq_thr, max_dist, num_checks);
}
void rvsref_destroy(void *p) { delete static_cast<Harness *>(p); }
// AbstractRobot::setSelfCollisionChecking (Scenario.cpp:101-103) on whichever robot class the harness holds
void rvsref_set_self_collision(void *p, int32_t enabled) { static_cast<Harness *>(p)->robot->setSelfCollisionChecking(enabled != 0); }

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


#ifndef RVS_DROPIN
// CPU arm of bench.py (`cpu_baseline.kind = "reference"`, `--impl reference`): base::RealVectorSpace::computeDistance — the reference's
// own loops and CollisionAndDistance primitives, compiled from /root/reference at -O3 — over N configurations, split statically over
// `threads` std::threads.  The reference is single-threaded and not re-entrant (robot->setConfiguration, per-State caches), so every
// thread owns a complete harness (robot + environment + state space), all built before the threads start; the static
// RealVectorSpaceConfig values are only read.  Every configuration gets a fresh State, as a sampler would make it (no cache hits).
// FK comes from the harness robot (the reference's URDF / mesh files do not travel to the GPU box).  d_c[N], profile[N x L].
}  // extern "C"
#include <thread>
extern "C" {
int rvsref_distance_mt(const rvs_robot_desc *rd, float eq_thr, float max_dist, const float *mn, const float *mx, const uint8_t *ground, const float *tol,
                       int32_t M, const float *q, int64_t N, int32_t threads, float *d_c, float *profile) {
    if (threads < 1) threads = 1;
    std::vector<void *> hs;
    for (int t = 0; t < threads; t++) {
        void *h = rvsref_create(rd, eq_thr, max_dist, 10);
        if (!h) return 1;
        rvsref_set_obstacles(h, mn, mx, ground, tol, M);
        hs.push_back(h);
    }
    const size_t L = size_t(rd->num_dofs);
    std::vector<std::thread> pool;
    const int64_t chunk = (N + threads - 1) / threads;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            auto *h = static_cast<Harness *>(hs[size_t(t)]);
            const int64_t b = int64_t(t) * chunk, e = std::min<int64_t>(N, b + chunk);
            for (int64_t i = b; i < e; i++) {
                auto s = state_of(*h, q + size_t(i) * L);
                d_c[i] = h->ss->computeDistance(s, false);
                for (size_t l = 0; l < L; l++) profile[size_t(i) * L + l] = s->getDistanceProfile()[l];
            }
        });
    for (auto &th : pool) th.join();
    for (void *h : hs) rvsref_destroy(h);
    return 0;
}
#endif

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
// rvsref_set_plan_log(path): the next rvsref_plan calls also write the planner's own log (outputPlannerData: header, both trees, the path;
// RBTConnect.cpp:203-253 / RRTConnect.cpp) to that file; NULL or "" switches it off again.
static std::string g_plan_log;
void rvsref_set_plan_log(const char *path) { g_plan_log = path ? path : ""; }

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
    if (!g_plan_log.empty()) planner->outputPlannerData(g_plan_log);
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

// ---- the robot classes themselves (rows A8, A10, A11) ------------------------------------------------------------------------------
// frames[N x n x 12] (row-major 3x3 rotation then translation, double): computeForwardKinematics (xArm6.cpp:107-134, Planar2DOF.cpp:82-108)
int rvsref_frames(void *p, const float *q, int64_t N, double *frames) {
    auto *h = static_cast<Harness *>(p);
    const size_t n = size_t(h->n);
    for (int64_t i = 0; i < N; i++) {
        const auto fr = h->robot->computeForwardKinematics(state_of(*h, q + size_t(i) * n));
        if (!fr) return 2;
        for (size_t k = 0; k < n; k++) {
            double *o = frames + (size_t(i) * n + k) * 12;
            for (int c = 0; c < 9; c++) o[c] = (*fr)[k].M.data[c];
            for (int c = 0; c < 3; c++) o[9 + c] = (*fr)[k].p(c);
        }
    }
    return 0;
}
// sk[N x (n+1) x 3]: computeSkeleton (xArm6.cpp:209-235, Planar2DOF.cpp:117-130)
int rvsref_skeleton(void *p, const float *q, int64_t N, float *sk) {
    auto *h = static_cast<Harness *>(p);
    const size_t n = size_t(h->n);
    for (int64_t i = 0; i < N; i++) {
        const auto m = h->robot->computeSkeleton(state_of(*h, q + size_t(i) * n));
        for (size_t k = 0; k <= n; k++)
            for (int c = 0; c < 3; c++) sk[(size_t(i) * (n + 1) + k) * 3 + size_t(c)] = (*m)(c, Eigen::Index(k));
    }
    return 0;
}
// R[N x n x (n+1)] row-major: computeEnclosingRadii (xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30)
int rvsref_enclosing_radii(void *p, const float *q, int64_t N, float *R) {
    auto *h = static_cast<Harness *>(p);
    const size_t n = size_t(h->n);
    for (int64_t i = 0; i < N; i++) {
        const auto m = h->robot->computeEnclosingRadii(state_of(*h, q + size_t(i) * n));
        for (size_t a = 0; a < n; a++)
            for (size_t b = 0; b <= n; b++) R[(size_t(i) * n + a) * (n + 1) + b] = (*m)(Eigen::Index(a), Eigen::Index(b));
    }
    return 0;
}
// xArm6 only: d[N x 4] = computeCapsulesDistance for the pairs {0-5, 1-5, 0-4, 1-4} in checkSelfCollision's order (xArm6.cpp:374-424);
// hit[N] = checkSelfCollision(q) (capsule stage; the mesh stage answers "collide").  Other robots: hit = their constant answer, d = 0.
int rvsref_self_collision(void *p, const float *q, int64_t N, float *d, uint8_t *hit) {
    auto *h = static_cast<Harness *>(p);
    const size_t n = size_t(h->n);
    auto *xarm = dynamic_cast<robots::xArm6 *>(h->robot.get());
    static const size_t pairs[4][2] = {{0, 5}, {1, 5}, {0, 4}, {1, 4}};
    for (int64_t i = 0; i < N; i++) {
        auto s = state_of(*h, q + size_t(i) * n);
        for (int k = 0; k < 4; k++) d[size_t(i) * 4 + size_t(k)] = xarm ? xarm->computeCapsulesDistance(s, pairs[k][0], pairs[k][1]) : 0.f;
        const std::shared_ptr<base::State> cs = s;
        hit[i] = h->robot->checkSelfCollision(cs) ? 1 : 0;
    }
    return 0;
}
// checkSelfCollision(q1, q2&) (xArm6.cpp:284-368): hit[E], q2_out[E x n] = q2 after the call (truncated on a hit).  eps_step = RRTConnectConfig::EPS_STEP.
int rvsref_self_collision_edges(void *p, const float *q1, const float *q2, int64_t E, float eps_step, uint8_t *hit, float *q2_out) {
    auto *h = static_cast<Harness *>(p);
    const size_t n = size_t(h->n);
    RRTConnectConfig::EPS_STEP = eps_step;
    for (int64_t e = 0; e < E; e++) {
        const std::shared_ptr<base::State> a = state_of(*h, q1 + size_t(e) * n);
        std::shared_ptr<base::State> b = state_of(*h, q2 + size_t(e) * n);
        hit[e] = h->robot->checkSelfCollision(a, b) ? 1 : 0;
        for (size_t i = 0; i < n; i++) q2_out[size_t(e) * n + i] = b->getCoord(int(i));
    }
    return 0;
}
// what the class reports about itself (limits[n x 2], radii[n]); returns num_DOFs
int rvsref_robot_info(void *p, float *limits, float *radii, int32_t *ground_included, float *gripper_length, int32_t *self_collision, int32_t *num_links) {
    auto *h = static_cast<Harness *>(p);
    const auto &rb = *h->robot;
    for (size_t i = 0; i < rb.getNumDOFs(); i++) {
        limits[2 * i] = rb.getLimits()[i].first; limits[2 * i + 1] = rb.getLimits()[i].second;
        radii[i] = rb.getCapsuleRadius(i);
    }
    *ground_included = int32_t(rb.getGroundIncluded());
    *gripper_length = rb.getGripperLength();
    *self_collision = rb.getSelfCollisionChecking() ? 1 : 0;
    *num_links = int32_t(rb.getNumLinks());
    return int(rb.getNumDOFs());
}

}  // extern "C"

#ifdef RVS_DROPIN
// ---- the drop-in class driven like a maintainer's code would drive it (tests/test_cpp_class.py) --------------------------------------
// base::RealVectorSpaceCUDA constructed through its URDF constructor (rvs::chain_from_urdf_file), single-query overrides against the
// batched overloads, the State cache rules (Q19 / Q20 of SURVEY.md §8.1), the staleness guard, the dynamic environment.
#include "rvs_urdf.h"
namespace {
std::string g_drv_error;
struct DrvSpace {
    std::shared_ptr<HarnessRobot> robot;
    std::shared_ptr<env::Environment> env;
    std::shared_ptr<base::RealVectorSpaceCUDA> ss;
};
void drv_add_boxes(env::Environment &e, const float *mn, const float *mx, const uint8_t *ground, const float *tol, int M) {
    for (int j = 0; j < M; j++) {
        const fcl::Vector3f lo(mn[3 * j], mn[3 * j + 1], mn[3 * j + 2]), hi(mx[3 * j], mx[3 * j + 1], mx[3 * j + 2]);
        auto obj = std::make_shared<env::Box>(fcl::Vector3f(hi - lo), fcl::Vector3f(0.5f * (lo + hi)), fcl::Quaternionf::Identity(), ground[j] ? "ground" : "random_obstacle");
        obj->getCollObject()->stubSetAABB(lo, hi);
        obj->setMinDistTol(tol[j]);
        e.addObject(obj);
    }
}
DrvSpace drv_make_space(const char *urdf_path, const rvs_robot_desc *rd, const float *mn, const float *mx, const uint8_t *ground, const float *tol, int M,
                        float eq_thr, int num_checks) {
    RealVectorSpaceConfig::EQUALITY_THRESHOLD = eq_thr;
    RealVectorSpaceConfig::MAX_DISTANCE = 1000.f;
    RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS = size_t(num_checks);
    DrvSpace d;
    d.robot = std::make_shared<HarnessRobot>(rd);
    d.robot->setType(rd->kind == RVS_ROBOT_XARM6 ? "xarm6" : rd->kind == RVS_ROBOT_PLANAR ? "planar_" + std::to_string(rd->num_dofs) + "dof"
                                                                                            : "spatial_" + std::to_string(rd->num_dofs) + "dof");
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    d.env = std::make_shared<env::Environment>(std::string("unused.yaml"), std::string(""));
    std::cout.rdbuf(keep);
    drv_add_boxes(*d.env, mn, mx, ground, tol, M);
    d.ss = std::make_shared<base::RealVectorSpaceCUDA>(size_t(rd->num_dofs), d.robot, d.env, std::string(urdf_path));
    return d;
}
std::shared_ptr<base::State> drv_state(base::StateSpace &ss, const float *q, int n) {
    Eigen::VectorXf v(n);
    for (int i = 0; i < n; i++) v(i) = q[i];
    return ss.getNewState(v);
}
}  // namespace

extern "C" {
const char *drv_last_error() { return g_drv_error.c_str(); }

// URDF reader only (no GPU needed)
int drv_parse_urdf(const char *path, const char *robot_type, rvs_robot_desc *out) {
    return rvs::chain_from_urdf_file(path, robot_type, out, &g_drv_error) ? 0 : 1;
}

// Runs N configurations through (a) the single-query overrides, state by state, and (b) the batched overloads.
// Outputs: valid[N], d_c[N], profile[N*L], nearest[N*M*L*6] (zeros where the State holds nullptr), has_nearest[N],
//          d_c_batch[N], valid_batch[N], edge_valid[N-1] (q_i -> q_{i+1}), edge_valid_batch[N-1],
//          under[N] / under_batch[N]: underestimation w.r.t. the planes of configuration `root` (must be collision-free).
int drv_run(const char *urdf_path, const rvs_robot_desc *rd, const float *aabb_min, const float *aabb_max, const uint8_t *is_ground, const float *tol, int M,
            float eq_thr, int num_checks, const float *q, int N, int root,
            uint8_t *valid, float *d_c, float *profile, float *nearest, uint8_t *has_nearest, float *d_c_batch, uint8_t *valid_batch,
            uint8_t *edge_valid, uint8_t *edge_valid_batch, float *under, float *under_batch, uint8_t *under_is_real) {
    try {
        const int n = rd->num_dofs;
        DrvSpace sp = drv_make_space(urdf_path, rd, aabb_min, aabb_max, is_ground, tol, M, eq_thr, num_checks);
        auto &ss = sp.ss;
        const size_t L = size_t(n);
        std::vector<std::shared_ptr<base::State>> single, batch;
        for (int i = 0; i < N; i++) { single.push_back(drv_state(*ss, q + size_t(i) * n, n)); batch.push_back(drv_state(*ss, q + size_t(i) * n, n)); }
        for (int i = 0; i < N; i++) {
            valid[i] = ss->isValid(single[i]) ? 1 : 0;
            d_c[i] = ss->computeDistance(single[i]);
            if (single[i]->getDistance() != d_c[i] || !single[i]->getIsRealDistance()) { g_drv_error = "State cache not written"; return 3; }
            std::memcpy(profile + i * L, single[i]->getDistanceProfile().data(), L * sizeof(float));
            const auto np = single[i]->getNearestPoints();
            has_nearest[i] = np ? 1 : 0;
            if (np)
                for (int j = 0; j < M; j++)
                    for (size_t l = 0; l < L; l++)
                        for (int k = 0; k < 6; k++) nearest[((size_t(i) * M + j) * L + l) * 6 + k] = (*np)[j](k, Eigen::Index(l));
            // cached value is returned without recomputation (RealVectorSpace.cpp:320-321)
            if (d_c[i] > 0 && ss->computeDistance(single[i]) != d_c[i]) { g_drv_error = "cache rule broken"; return 4; }
        }
        const std::vector<uint8_t> vb = ss->isValidBatch(batch);
        const std::vector<float> db = ss->computeDistanceBatch(batch);
        std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> edges;
        for (int i = 0; i + 1 < N; i++) {
            edge_valid[i] = ss->isValid(single[i], single[i + 1]) ? 1 : 0;
            edges.emplace_back(batch[i], batch[i + 1]);
        }
        const std::vector<uint8_t> eb = ss->isValidEdges(edges);
        for (int i = 0; i < N; i++) {
            valid_batch[i] = vb[i]; d_c_batch[i] = db[i];
            if (batch[i]->getDistance() != db[i]) { g_drv_error = "batched State cache not written"; return 5; }
            if (std::memcmp(batch[i]->getDistanceProfile().data(), profile + i * L, L * sizeof(float)) != 0) { g_drv_error = "batched profile differs from single-query profile"; return 6; }
        }
        for (int i = 0; i + 1 < N; i++) edge_valid_batch[i] = eb[i];
        // underestimation against the root's planes on fresh states (a state with a real distance returns it unchanged, Q20)
        const auto planes = single[root]->getNearestPoints();
        if (!planes) { g_drv_error = "root configuration collides"; return 7; }
        if (ss->computeDistanceUnderestimation(single[0], planes) != d_c[0] && d_c[0] > 0) { g_drv_error = "real distance must be returned unchanged"; return 8; }
        std::vector<std::shared_ptr<base::State>> fresh_single, fresh_batch;
        for (int i = 0; i < N; i++) { fresh_single.push_back(drv_state(*ss, q + size_t(i) * n, n)); fresh_batch.push_back(drv_state(*ss, q + size_t(i) * n, n)); }
        for (int i = 0; i < N; i++) {
            under[i] = ss->computeDistanceUnderestimation(fresh_single[i], planes);
            under_is_real[i] = fresh_single[i]->getIsRealDistance() ? 1 : 0;
            // rejected (a plane distance < 0): returns 0 and leaves the state untouched (d_c stays -1)
            if (under[i] == 0 && fresh_single[i]->getDistance() != -1) { g_drv_error = "rejected underestimation touched the state"; return 9; }
        }
        const std::vector<float> ub = ss->computeDistanceUnderestimationBatch(fresh_batch, planes);
        for (int i = 0; i < N; i++) under_batch[i] = ub[i];
        // profile-only batch: distance + profile land on the State, planes do not; the next single call supplies them
        {
            std::vector<std::shared_ptr<base::State>> lean;
            for (int i = 0; i < N; i++) lean.push_back(drv_state(*ss, q + size_t(i) * n, n));
            const std::vector<float> dl = ss->computeDistanceBatch(lean, false, false);
            for (int i = 0; i < N; i++) {
                if (dl[i] != d_c[i]) { g_drv_error = "profile-only batch distance differs"; return 13; }
                if (lean[i]->getNearestPoints() != nullptr) { g_drv_error = "profile-only batch left nearest points behind"; return 14; }
                if (d_c[i] > 0 && (ss->computeDistance(lean[i]) != d_c[i] || lean[i]->getNearestPoints() == nullptr)) { g_drv_error = "single call after a profile-only batch did not supply the planes"; return 15; }
            }
        }
        // staleness guard: the environment changes behind the state space's back (as MotionValidity::check / DRGBT / RRTx do) and the
        // very next query must see it.  Object 0..M-1 are pushed 10 m away: everything becomes valid.
        {
            for (int j = 0; j < M; j++) {
                auto obj = sp.env->getObject(size_t(j));
                const fcl::Vector3f far_lo(10.f + j, 10.f, 10.f), far_hi(10.2f + j, 10.2f, 10.2f);
                obj->getCollObject()->stubSetAABB(far_lo, far_hi);
            }
            for (int i = 0; i < N; i++)
                if (!ss->isValid(drv_state(*ss, q + size_t(i) * n, n))) { g_drv_error = "stale obstacle table: a query after the environment moved still collides"; return 16; }
        }
        return 0;
    } catch (const std::exception &e) {
        g_drv_error = e.what();
        return 1;
    }
}

// Spine sets and self-collision through the class (xArm6): extendSpinesBatch from ONE root towards S targets, checkEdgesBatch and
// checkSelfCollisionEdges on E edges.  Outputs: q_reached[S x n], d_reached[S], status[S]; edge_ok[E]; sc_hit[E], q2_after[E x n].
int drv_spines(const char *urdf_path, const rvs_robot_desc *rd, int self_collision, const float *aabb_min, const float *aabb_max, const uint8_t *is_ground, const float *tol,
               int M, const float *q_root, const float *q_targets, int S, int num_layers, int num_iter_spine, int use_expanded_bubble, float d_crit,
               float *q_reached, float *d_reached, uint8_t *status, const float *e1, const float *e2, int E, uint8_t *edge_ok, uint8_t *sc_hit, float *q2_after) {
    try {
        const int n = rd->num_dofs;
        DrvSpace sp = drv_make_space(urdf_path, rd, aabb_min, aabb_max, is_ground, tol, M, 1e-4f, 10);
        sp.robot->setSelfCollisionChecking(self_collision != 0);
        RRTConnectConfig::EPS_STEP = 0.1f;
        auto &ss = sp.ss;
        auto root = drv_state(*ss, q_root, n);
        std::vector<std::shared_ptr<base::State>> targets;
        for (int s = 0; s < S; s++) targets.push_back(drv_state(*ss, q_targets + size_t(s) * n, n));
        const auto res = ss->extendSpinesBatch(root, targets, size_t(num_layers), size_t(num_iter_spine), use_expanded_bubble != 0, d_crit);
        for (int s = 0; s < S; s++) {
            status[s] = uint8_t(res[size_t(s)].first);
            d_reached[s] = res[size_t(s)].second->getDistance();
            for (int i = 0; i < n; i++) q_reached[size_t(s) * n + i] = res[size_t(s)].second->getCoord(i);
        }
        std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> edges;
        for (int e = 0; e < E; e++) edges.emplace_back(drv_state(*ss, e1 + size_t(e) * n, n), drv_state(*ss, e2 + size_t(e) * n, n));
        const auto ok = ss->checkEdgesBatch(edges);
        const auto hit = ss->checkSelfCollisionEdges(edges);
        for (int e = 0; e < E; e++) {
            edge_ok[e] = ok[size_t(e)]; sc_hit[e] = hit[size_t(e)];
            for (int i = 0; i < n; i++) q2_after[size_t(e) * n + i] = edges[size_t(e)].second->getCoord(i);
        }
        return 0;
    } catch (const std::exception &e) {
        g_drv_error = e.what();
        return 1;
    }
}

// Dynamic environment (row f-2): T consecutive MotionValidity-style checks q_t -> q_{t+1} through checkMotion (the environment
// advances as the reference's loop would), then one what-if batch of all T motions against the final environment.
// Outputs: valid[T], passed[T] (validity points appended to `path`), pos_after[T*M*3] (object positions mirrored back into
// env::Object after every check), batch_valid[T].
int drv_motion(const char *urdf_path, const rvs_robot_desc *rd, const float *dim, const float *pos, const float *vel, const float *max_vel,
               const uint8_t *is_dynamic, const uint8_t *is_ground, const float *tol, int M, int motion_type,
               const float *ws /*center xyz, radius, base_radius, robot_max_vel*/, const float *q, int T, int num_checks, float delta_time, uint8_t *valid,
               int32_t *passed, float *pos_after, uint8_t *batch_valid) {
    try {
        const int n = rd->num_dofs;
        DrvSpace sp = drv_make_space(urdf_path, rd, nullptr, nullptr, nullptr, nullptr, 0, 1e-4f, 10);
        Harness hh;
        hh.env = sp.env; hh.n = n;
        motion_env_setup(&hh, ws, motion_type, M, dim, pos, vel, max_vel, is_dynamic, is_ground, tol);
        auto &ss = sp.ss;
        ss->attachDynamicEnvironment(rvs_motion_type(motion_type));
        std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> motions;
        for (int t = 0; t < T; t++) {
            auto a = drv_state(*ss, q + size_t(t) * n, n), b = drv_state(*ss, q + size_t(t + 1) * n, n);
            motions.emplace_back(a, b);
            std::vector<std::shared_ptr<base::State>> path;
            valid[t] = ss->checkMotion(a, b, size_t(num_checks), delta_time, &path) ? 1 : 0;
            passed[t] = int32_t(path.size());
            for (int j = 0; j < M; j++)
                for (int k = 0; k < 3; k++) pos_after[(size_t(t) * M + j) * 3 + k] = ss->env->getObject(size_t(j))->getPosition()(k);
        }
        const std::vector<uint8_t> bv = ss->checkMotionBatch(motions, size_t(num_checks), delta_time);
        for (int t = 0; t < T; t++) batch_valid[t] = bv[size_t(t)];
        // the other overload (MotionValidity.cpp:60-82): the same motion handed over as explicit trajectory points must get the what-if
        // verdict of the interpolated edge, append the points that passed, and move the environment
        {
            const auto &m = motions[size_t(T - 1)];
            const float dist = ss->getNorm(m.first, m.second);
            std::vector<Eigen::VectorXf> traj;
            for (size_t k = 1; k <= size_t(num_checks); k++) traj.push_back(ss->interpolateEdge(m.first, m.second, dist * k / num_checks, dist)->getCoord());
            const float before = ss->env->getObject(size_t(M - 1))->getPosition()(0);
            std::vector<std::shared_ptr<base::State>> path;
            const bool ok = ss->checkMotion(traj, delta_time, &path);
            if (ok != (bv[size_t(T - 1)] != 0)) { g_drv_error = "checkMotion(traj_points) disagrees with the what-if check of the same motion"; return 10; }
            if (ok && path.size() != size_t(num_checks)) { g_drv_error = "checkMotion(traj_points) did not append every point"; return 11; }
            if (is_dynamic[M - 1] && max_vel[M - 1] > 0 && ss->env->getObject(size_t(M - 1))->getPosition()(0) == before) { g_drv_error = "checkMotion(traj_points) did not move the environment"; return 12; }
        }
        return 0;
    } catch (const std::exception &e) {
        g_drv_error = e.what();
        return 1;
    }
}
}  // extern "C"
#endif  // RVS_DROPIN

// ---- RRTx's edge-batch consumers (row f-4): RRTx.cpp compiled unmodified -----------------------------------------------------------
// RRTx.cpp references five symbols of the trajectory classes (spline / Ruckig interpolation of the executed motion: outside the path,
// and their sources need more of Eigen than the work-alike has).  Four are never reached with TRAJECTORY_INTERPOLATION = None; the
// UpdatingState constructor is, from RRTx's own constructor: it leaves a zeroed object (null shared_ptrs), which RRTx only ever hands a
// null trajectory to.  The definitions carry the symbols' mangled names so that the unmodified object file links.
extern "C" {
[[noreturn]] static void rvsref_unlinked_abort() { std::cerr << "rvsref: a trajectory symbol outside the harness was called" << std::endl; std::abort(); }
void rvsref_unlinked_a() __asm__("_ZN8planning10trajectory10TrajectoryC1ERKSt10shared_ptrIN4base10StateSpaceEENS0_5StateEf");
void rvsref_unlinked_a() { rvsref_unlinked_abort(); }
void rvsref_unlinked_b() __asm__("_ZN8planning10trajectory16TrajectoryRuckigC1ERKSt10shared_ptrIN4base10StateSpaceEENS0_5StateEf");
void rvsref_unlinked_b() { rvsref_unlinked_abort(); }
void rvsref_unlinked_c() __asm__("_ZN8planning10trajectory5StateC1ERKN5Eigen6MatrixIfLin1ELi1EEE");
void rvsref_unlinked_c() { rvsref_unlinked_abort(); }
void rvsref_unlinked_d() __asm__("_ZN8planning10trajectory13UpdatingState6updateERSt10shared_ptrIN4base5StateEES6_S5_RNS4_6StatusE");
void rvsref_unlinked_d() { rvsref_unlinked_abort(); }
void rvsref_unlinked_e(void *self) __asm__("_ZN8planning10trajectory13UpdatingStateC1ERKSt10shared_ptrIN4base10StateSpaceEENS_23TrajectoryInterpolationEf");
void rvsref_unlinked_e(void *self) { std::memset(self, 0, sizeof(planning::trajectory::UpdatingState)); }

// The tree RRTx grows from q_goal when its random states are `samples`: the tree-growth block of RRTx::solve (RRTx.cpp:157-190) —
// getNearestState, extend, setCost, findNeighbors, chooseParent, upgradeTree, rewireNeighbors, all the reference's own members, called
// in the reference's order on the planner object RRTx's own constructor set up (tree, kd-tree, eta / mi / zeta / gamma_rrt).
// r_rewire > 0: constant radius (RRTxConfig::R_REWIRE), else shrinkingBallRadius(tree size) per iteration; radii_out [S] records it.
// path_start >= 0: updatePath() / computePath() with start_state = that tree state.  Returns the number of states, -1 when the tree
// outgrows max_states, -2 when the constructor throws (q_start invalid).
int64_t rvsref_rrtx_grow(void *p, float eps_step, float r_rewire, int32_t max_neighbors, const float *q_start, const float *q_goal, const float *samples,
                         int32_t S, int64_t max_states, float *coord, int32_t *parent, float *cost, int32_t *new_index, float *radii_out, int32_t path_start,
                         int32_t *path, int32_t *path_len, int32_t *path_invalid) {
    auto *h = static_cast<Harness *>(p);
    RRTxConfig::EPS_STEP = eps_step;
    RRTxConfig::R_REWIRE = r_rewire;
    RRTxConfig::MAX_NEIGHBORS = size_t(max_neighbors);
    RRTxConfig::TRAJECTORY_INTERPOLATION = planning::TrajectoryInterpolation::None;
    std::unique_ptr<planning::rrtx::RRTx> planner;
    try {
        planner = std::make_unique<planning::rrtx::RRTx>(h->ss, state_of(*h, q_start), state_of(*h, q_goal));
    } catch (const std::domain_error &) {
        return -2;
    }
    const size_t n = size_t(h->n);
    for (int32_t s = 0; s < S; s++) {
        if (int64_t(planner->tree->getNumStates()) >= max_states) return -1;
        planner->r_rewire = RRTxConfig::R_REWIRE > 0 ? RRTxConfig::R_REWIRE : planner->shrinkingBallRadius(planner->tree->getNumStates());
        if (radii_out) radii_out[s] = planner->r_rewire;
        const auto q_rand = state_of(*h, samples + size_t(s) * n);
        const auto q_near = planner->tree->getNearestState(q_rand);
        base::State::Status status;
        std::shared_ptr<base::State> q_new;
        std::tie(status, q_new) = planner->extend(q_near, q_rand);
        new_index[s] = -1;
        if (status == base::State::Status::Trapped) continue;
        q_new->setCost(q_near->getCost() + planner->distance(q_near, q_new));
        const auto neighbors = planner->findNeighbors(q_new, planner->r_rewire);
        planner->chooseParent(q_new, neighbors);
        planner->tree->upgradeTree(q_new, q_new->getParent());
        planner->rewireNeighbors(q_new, neighbors);
        new_index[s] = int32_t(q_new->getIdx());
    }
    const int64_t count = int64_t(planner->tree->getNumStates());
    for (int64_t k = 0; k < count; k++) {
        const auto q = planner->tree->getState(size_t(k));
        for (size_t i = 0; i < n; i++) coord[size_t(k) * n + i] = q->getCoord(int(i));
        parent[k] = q->getParent() ? int32_t(q->getParent()->getIdx()) : -1;
        cost[k] = q->getCost();
    }
    if (path_start >= 0 && path_start < count) {
        planner->start_state = planner->tree->getState(size_t(path_start));
        planner->computePath();
        *path_len = int32_t(planner->path_current.size());
        for (int32_t k = 0; k < *path_len; k++) path[k] = int32_t(planner->path_current[size_t(k)]->getIdx());
        *path_invalid = planner->path_current.empty() ? 0 : (planner->updatePath() ? 1 : 0);
    }
    return count;
}
}
