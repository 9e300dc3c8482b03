// Stand-in for include/state_spaces/StateSpace.h (only the members the CUDA space uses or overrides).
#pragma once
#include "AbstractRobot.h"
#include "Environment.h"
#include "State.h"
namespace base {
class StateSpace {
public:
    base::StateSpaceType state_space_type = base::StateSpaceType::Abstract;
    size_t num_dimensions = 0;
    std::shared_ptr<robots::AbstractRobot> robot;
    std::shared_ptr<env::Environment> env;
    StateSpace(size_t n, const std::shared_ptr<robots::AbstractRobot> robot_, const std::shared_ptr<env::Environment> env_)
        : num_dimensions(n), robot(robot_), env(env_) {}
    virtual ~StateSpace() = default;
    inline void setStateSpaceType(base::StateSpaceType t) { state_space_type = t; }
    virtual std::shared_ptr<base::State> getNewState(const Eigen::VectorXf &coord) = 0;
    virtual bool isValid(const std::shared_ptr<base::State> q) = 0;
    virtual bool isValid(const std::shared_ptr<base::State> q1, const std::shared_ptr<base::State> q2) = 0;
    virtual float computeDistance(const std::shared_ptr<base::State> q, bool compute_again = false) = 0;
    virtual float computeDistanceUnderestimation(const std::shared_ptr<base::State> q,
                                                 const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points) = 0;
};
}  // namespace base
