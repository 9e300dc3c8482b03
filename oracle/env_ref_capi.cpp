// env_ref_capi.cpp — C entry point around the REFERENCE's own env::Environment::updateEnvironment, compiled from
// /root/reference/src/environments/{Environment,Object}.cpp and objects/Box.cpp (unmodified, where they lie) against oracle/eigen_shim and
// oracle/env_shim.  TEST INFRASTRUCTURE ONLY: pins the motion-law logic of oracle/rvs_oracle_env.hpp.  Output: oracle/_ref/ (git-ignored).
//
// env::Environment keeps its workspace parameters, motion type and the two_tunnels bookkeeping private and only reads them from YAML;
// the harness sets them directly (`private` is opened for this translation unit only — the reference sources are compiled as they are).
#include <cstdint>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define private public
#include "Environment.h"
#undef private

extern "C" int envref_update(const float *env_params /*WS_center xyz, WS_radius, base_radius, robot_max_vel*/, int32_t motion_type, int32_t M, const float *dim,
                             float *pos, float *vel, const float *max_vel, int32_t *sign, float *path_len, const uint8_t *dynamic, float dt, int32_t steps,
                             float *aabb_min, float *aabb_max) {
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());                                  // the constructor and addObject print
    env::Environment e(std::string("unused.yaml"), std::string(""));
    std::cout.rdbuf(keep);
    e.WS_center = fcl::Vector3f(env_params[0], env_params[1], env_params[2]);
    e.WS_radius = env_params[3];
    e.setBaseRadius(env_params[4]);
    e.setRobotMaxVel(env_params[5]);
    e.setGroundIncluded(2);
    e.motion_type = static_cast<decltype(e.motion_type)>(motion_type);       // enum order: straight, circular, two_tunnels, random_directions, light_directions
    e.sign = Eigen::VectorXi::Ones(M);
    e.path_len = Eigen::VectorXf::Zero(M);
    for (int j = 0; j < M; j++) {
        auto obj = std::make_shared<env::Box>(fcl::Vector3f(dim[3 * j], dim[3 * j + 1], dim[3 * j + 2]), fcl::Vector3f(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]),
                                              fcl::Quaternionf::Identity(), dynamic[j] ? "dynamic_obstacle" : "");
        obj->setMaxVel(max_vel[j]);
        e.addObject(obj, fcl::Vector3f(vel[3 * j], vel[3 * j + 1], vel[3 * j + 2]));
        e.sign(j) = sign[j];
        e.path_len(j) = path_len[j];
    }
    for (int k = 0; k < steps; k++) e.updateEnvironment(dt);
    for (int j = 0; j < M; j++) {
        const auto obj = e.getObject(size_t(j));
        for (int c = 0; c < 3; c++) {
            pos[3 * j + c] = obj->getPosition()(c);
            vel[3 * j + c] = obj->getVelocity()(c);
            if (aabb_min) aabb_min[3 * j + c] = obj->getCollObject()->getAABB().min_(c);
            if (aabb_max) aabb_max[3 * j + c] = obj->getCollObject()->getAABB().max_(c);
        }
        sign[j] = e.sign(j);
        path_len[j] = e.path_len(j);
    }
    return 0;
}

// Environment::moveRandomDirections (Environment.cpp:289-333) after srand(seed): `steps` updates; pos_tl / vel_tl / acc_tl [steps x M x 3] record every
// object after every update, *next_rand = the value rand() returns right after the run (pins how many draws the run consumed).
extern "C" int envref_update_random(const float *env_params, int32_t ground_included, int32_t motion_type, int32_t M, const float *dim, const float *pos0,
                                    const float *vel0, const float *acc0, const float *max_vel, const uint8_t *dynamic, float dt, int32_t steps, uint32_t seed,
                                    float *pos_tl, float *vel_tl, float *acc_tl, int32_t *next_rand) {
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    env::Environment e(std::string("unused.yaml"), std::string(""));
    std::cout.rdbuf(keep);
    e.WS_center = fcl::Vector3f(env_params[0], env_params[1], env_params[2]);
    e.WS_radius = env_params[3];
    e.setBaseRadius(env_params[4]);
    e.setRobotMaxVel(env_params[5]);
    e.setGroundIncluded(size_t(ground_included));
    e.motion_type = static_cast<decltype(e.motion_type)>(motion_type);
    e.sign = Eigen::VectorXi::Ones(M);
    e.path_len = Eigen::VectorXf::Zero(M);
    std::cout.rdbuf(sink.rdbuf());
    for (int j = 0; j < M; j++) {
        auto obj = std::make_shared<env::Box>(fcl::Vector3f(dim[3 * j], dim[3 * j + 1], dim[3 * j + 2]), fcl::Vector3f(pos0[3 * j], pos0[3 * j + 1], pos0[3 * j + 2]),
                                              fcl::Quaternionf::Identity(), dynamic[j] ? "dynamic_obstacle" : "");
        obj->setMaxVel(max_vel[j]);
        e.addObject(obj, fcl::Vector3f(vel0[3 * j], vel0[3 * j + 1], vel0[3 * j + 2]), fcl::Vector3f(acc0[3 * j], acc0[3 * j + 1], acc0[3 * j + 2]));
    }
    std::cout.rdbuf(keep);
    std::srand(seed);
    for (int k = 0; k < steps; k++) {
        e.updateEnvironment(dt);
        for (int j = 0; j < M; j++) {
            const auto obj = e.getObject(size_t(j));
            for (int c = 0; c < 3; c++) {
                pos_tl[(size_t(k) * M + j) * 3 + c] = obj->getPosition()(c);
                vel_tl[(size_t(k) * M + j) * 3 + c] = obj->getVelocity()(c);
                acc_tl[(size_t(k) * M + j) * 3 + c] = obj->getAcceleration()(c);
            }
        }
    }
    *next_rand = std::rand();                                        // where the library's stream stands after the run
    return 0;
}
