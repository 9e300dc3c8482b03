// Stand-in for include/state_spaces/real_vector_space/RealVectorSpace.h. The CPU hot-path bodies are deliberately absent
// (they throw): in this repository nothing may fall back to a CPU implementation.
#pragma once
#include <stdexcept>
#include "StateSpace.h"
namespace base {
class RealVectorSpace : public base::StateSpace {
public:
    RealVectorSpace(size_t n, const std::shared_ptr<robots::AbstractRobot> robot_, const std::shared_ptr<env::Environment> env_)
        : StateSpace(n, robot_, env_) { setStateSpaceType(base::StateSpaceType::RealVectorSpace); }
    std::shared_ptr<base::State> getNewState(const Eigen::VectorXf &coord) override { return std::make_shared<base::RealVectorSpaceState>(coord); }
    bool isValid(const std::shared_ptr<base::State>) override { throw std::logic_error("stub: CPU path not available"); }
    bool isValid(const std::shared_ptr<base::State>, const std::shared_ptr<base::State>) override { throw std::logic_error("stub: CPU path not available"); }
    float computeDistance(const std::shared_ptr<base::State>, bool) override { throw std::logic_error("stub: CPU path not available"); }
    float computeDistanceUnderestimation(const std::shared_ptr<base::State>, const std::shared_ptr<std::vector<Eigen::MatrixXf>>) override {
        throw std::logic_error("stub: CPU path not available");
    }
};
}  // namespace base
