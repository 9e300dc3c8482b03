// tests/cpp/class_driver.cpp — exercises base::RealVectorSpaceCUDA (rpmplv2_b200/cpp) through the reference's own call
// pattern: build a robot + environment, construct the state space, call the single-query methods and the batched
// overloads, return what landed in the base::State caches. Built against rpmplv2_b200/cpp/stubs (Eigen/FCL/KDL are not
// in this image); called from tests/test_cpp_class.py through ctypes.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "RealVectorSpaceCUDA.h"
#include "rvs_urdf.h"

namespace {
std::string g_error;

std::shared_ptr<base::RealVectorSpaceCUDA> make_space(const char *urdf_path, const char *robot_type, int n, const float *radii,
                                                       float gripper_length, int ground_included, const float *limits,
                                                       const float *aabb_min, const float *aabb_max, const uint8_t *is_ground, const float *tol, int M,
                                                       float eq_thr, int num_checks) {
    RealVectorSpaceConfig::EQUALITY_THRESHOLD = eq_thr;
    RealVectorSpaceConfig::NUM_INTERPOLATION_VALIDITY_CHECKS = size_t(num_checks);
    auto robot = std::make_shared<robots::AbstractRobot>();
    robot->type = robot_type;
    robot->num_DOFs = robot->num_links = size_t(n);
    robot->capsules_radius.assign(radii, radii + n);
    robot->gripper_length = gripper_length;
    robot->ground_included = size_t(ground_included);
    for (int i = 0; i < n; i++) robot->limits.emplace_back(limits[2 * i], limits[2 * i + 1]);
    auto env = std::make_shared<env::Environment>();
    for (int j = 0; j < M; j++) {
        fcl::AABB box;
        for (int k = 0; k < 3; k++) { box.min_[k] = aabb_min[3 * j + k]; box.max_[k] = aabb_max[3 * j + k]; }
        auto obj = std::make_shared<env::Box>();
        obj->setCollObject(std::make_shared<fcl::CollisionObjectf>(box));
        obj->setLabel(is_ground[j] ? "ground" : "random_obstacle");
        obj->setMinDistTol(tol[j]);
        env->addObject(obj);
    }
    return std::make_shared<base::RealVectorSpaceCUDA>(size_t(n), robot, env, std::string(urdf_path));
}

std::shared_ptr<base::State> state_of(base::StateSpace &ss, const float *q, int n) {
    Eigen::VectorXf v(n);
    for (int i = 0; i < n; i++) v(i) = q[i];
    return ss.getNewState(v);
}
}  // namespace

extern "C" {

const char *drv_last_error() { return g_error.c_str(); }

// URDF reader only (no GPU needed)
int drv_parse_urdf(const char *path, const char *robot_type, rvs_robot_desc *out) {
    return rvs::chain_from_urdf_file(path, robot_type, out, &g_error) ? 0 : 1;
}

// Runs N configurations through (a) the single-query overrides, state by state, and (b) the batched overloads.
// Outputs: valid[N], d_c[N], profile[N*L], nearest[N*M*L*6] (zeros where the State holds nullptr), has_nearest[N],
//          d_c_batch[N], valid_batch[N], edge_valid[N-1] (q_i -> q_{i+1}), edge_valid_batch[N-1],
//          under[N] / under_batch[N]: underestimation w.r.t. the planes of configuration `root` (must be collision-free).
int drv_run(const char *urdf_path, const char *robot_type, int n, const float *radii, float gripper_length, int ground_included,
            const float *limits, const float *aabb_min, const float *aabb_max, const uint8_t *is_ground, const float *tol, int M,
            float eq_thr, int num_checks, const float *q, int N, int root,
            uint8_t *valid, float *d_c, float *profile, float *nearest, uint8_t *has_nearest, float *d_c_batch, uint8_t *valid_batch,
            uint8_t *edge_valid, uint8_t *edge_valid_batch, float *under, float *under_batch, uint8_t *under_is_real) {
    try {
        auto ss = make_space(urdf_path, robot_type, n, radii, gripper_length, ground_included, limits, aabb_min, aabb_max, is_ground, tol, M, eq_thr, num_checks);
        if (ss->state_space_type != base::StateSpaceType::RealVectorSpaceCUDA) { g_error = "wrong state space type"; return 2; }
        const size_t L = size_t(n);
        std::vector<std::shared_ptr<base::State>> single, batch;
        for (int i = 0; i < N; i++) { single.push_back(state_of(*ss, q + size_t(i) * n, n)); batch.push_back(state_of(*ss, q + size_t(i) * n, n)); }
        for (int i = 0; i < N; i++) {
            valid[i] = ss->isValid(single[i]) ? 1 : 0;
            d_c[i] = ss->computeDistance(single[i]);
            if (single[i]->getDistance() != d_c[i] || !single[i]->getIsRealDistance()) { g_error = "State cache not written"; return 3; }
            std::memcpy(profile + i * L, single[i]->getDistanceProfile().data(), L * sizeof(float));
            const auto np = single[i]->getNearestPoints();
            has_nearest[i] = np ? 1 : 0;
            if (np)
                for (int j = 0; j < M; j++)
                    for (size_t l = 0; l < L; l++)
                        for (int k = 0; k < 6; k++) nearest[((size_t(i) * M + j) * L + l) * 6 + k] = (*np)[j](k, Eigen::Index(l));
            // cached value is returned without recomputation (RealVectorSpace.cpp:320-321)
            if (d_c[i] > 0 && ss->computeDistance(single[i]) != d_c[i]) { g_error = "cache rule broken"; return 4; }
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
            if (batch[i]->getDistance() != db[i]) { g_error = "batched State cache not written"; return 5; }
            if (std::memcmp(batch[i]->getDistanceProfile().data(), profile + i * L, L * sizeof(float)) != 0) { g_error = "batched profile differs from single-query profile"; return 6; }
        }
        for (int i = 0; i + 1 < N; i++) edge_valid_batch[i] = eb[i];
        // underestimation against the root's planes on fresh states (a state with a real distance returns it unchanged, Q20)
        const auto planes = single[root]->getNearestPoints();
        if (!planes) { g_error = "root configuration collides"; return 7; }
        if (ss->computeDistanceUnderestimation(single[0], planes) != d_c[0] && d_c[0] > 0) { g_error = "real distance must be returned unchanged"; return 8; }
        std::vector<std::shared_ptr<base::State>> fresh_single, fresh_batch;
        for (int i = 0; i < N; i++) { fresh_single.push_back(state_of(*ss, q + size_t(i) * n, n)); fresh_batch.push_back(state_of(*ss, q + size_t(i) * n, n)); }
        for (int i = 0; i < N; i++) {
            under[i] = ss->computeDistanceUnderestimation(fresh_single[i], planes);
            under_is_real[i] = fresh_single[i]->getIsRealDistance() ? 1 : 0;
            // rejected (a plane distance < 0): returns 0 and leaves the state untouched (d_c stays -1)
            if (under[i] == 0 && fresh_single[i]->getDistance() != -1) { g_error = "rejected underestimation touched the state"; return 9; }
        }
        const std::vector<float> ub = ss->computeDistanceUnderestimationBatch(fresh_batch, planes);
        for (int i = 0; i < N; i++) under_batch[i] = ub[i];
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    }
}


// Dynamic environment (row f-2): T consecutive MotionValidity-style checks q_t -> q_{t+1} through checkMotion (the environment
// advances as the reference's loop would), then one what-if batch of all T motions against the final environment.
// Outputs: valid[T], passed[T] (validity points appended to `path`), pos_after[T*M*3] (object positions mirrored back into
// env::Object after every check), batch_valid[T].
int drv_motion(const char *urdf_path, const char *robot_type, int n, const float *radii, float gripper_length, int ground_included, const float *limits,
               const float *dim, const float *pos, const float *vel, const float *max_vel, const uint8_t *is_dynamic, const uint8_t *is_ground,
               const float *tol, int M, int motion_type, const float *ws /*center xyz, radius, base_radius, robot_max_vel*/,
               const float *q, int T, int num_checks, float delta_time, uint8_t *valid, int32_t *passed, float *pos_after, uint8_t *batch_valid) {
    try {
        std::vector<float> mn(size_t(3 * M)), mx(size_t(3 * M));
        for (int j = 0; j < M; j++)
            for (int k = 0; k < 3; k++) { const float half = 0.5f * dim[3 * j + k]; mn[3 * j + k] = pos[3 * j + k] + (-half); mx[3 * j + k] = pos[3 * j + k] + half; }
        auto ss = make_space(urdf_path, robot_type, n, radii, gripper_length, ground_included, limits, mn.data(), mx.data(), is_ground, tol, M, 1e-4f, 10);
        for (int j = 0; j < M; j++) {
            auto obj = ss->env->getObject(size_t(j));
            if (is_dynamic[j]) obj->setLabel("dynamic_obstacle");
            obj->setPosition(fcl::Vector3f(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]));
            obj->setVelocity(fcl::Vector3f(vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]));
            obj->setMaxVel(max_vel[j]);
        }
        ss->env->stubSetWorkspace(fcl::Vector3f(ws[0], ws[1], ws[2]), ws[3]);
        ss->env->setBaseRadius(ws[4]);
        ss->env->setRobotMaxVel(ws[5]);
        ss->attachDynamicEnvironment(rvs_motion_type(motion_type));
        std::vector<std::pair<std::shared_ptr<base::State>, std::shared_ptr<base::State>>> motions;
        for (int t = 0; t < T; t++) {
            auto a = state_of(*ss, q + size_t(t) * n, n), b = state_of(*ss, q + size_t(t + 1) * n, n);
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
            if (ok != (bv[size_t(T - 1)] != 0)) { g_error = "checkMotion(traj_points) disagrees with the what-if check of the same motion"; return 10; }
            if (ok && path.size() != size_t(num_checks)) { g_error = "checkMotion(traj_points) did not append every point"; return 11; }
            if (is_dynamic[M - 1] && max_vel[M - 1] > 0 && ss->env->getObject(size_t(M - 1))->getPosition()(0) == before) { g_error = "checkMotion(traj_points) did not move the environment"; return 12; }
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    }
}

}  // extern "C"
