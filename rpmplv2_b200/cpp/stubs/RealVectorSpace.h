// Stand-in for include/state_spaces/real_vector_space/RealVectorSpace.h. The CPU hot-path bodies are deliberately absent
// (they throw): in this repository nothing may fall back to a CPU implementation.
#pragma once
#include <cmath>
#include <stdexcept>
#include "StateSpace.h"
namespace base {
class RealVectorSpace : public base::StateSpace {
public:
    RealVectorSpace(size_t n, const std::shared_ptr<robots::AbstractRobot> robot_, const std::shared_ptr<env::Environment> env_)
        : StateSpace(n, robot_, env_) { setStateSpaceType(base::StateSpaceType::RealVectorSpace); }
    std::shared_ptr<base::State> getNewState(const Eigen::VectorXf &coord) override { return std::make_shared<base::RealVectorSpaceState>(coord); }
    // host-side helpers the CUDA space inherits unchanged (RealVectorSpace.cpp:27-88); restated here only so that the stub links
    float getNorm(const std::shared_ptr<base::State> q1, const std::shared_ptr<base::State> q2) {
        float s = 0;
        for (Eigen::Index i = 0; i < q1->getCoord().size(); i++) { const float d = q1->getCoord()(i) - q2->getCoord()(i); s += d * d; }
        return std::sqrt(s);
    }
    std::shared_ptr<base::State> interpolateEdge(const std::shared_ptr<base::State> q1, const std::shared_ptr<base::State> q2, float step, float dist = -1) {
        if (dist < 0) dist = getNorm(q1, q2);
        Eigen::VectorXf c(q1->getCoord().size());
        for (Eigen::Index i = 0; i < c.size(); i++)
            c(i) = dist > 0 ? q1->getCoord()(i) + (q2->getCoord()(i) - q1->getCoord()(i)) * (step / dist) : q2->getCoord()(i);
        return getNewState(c);
    }
    bool isValid(const std::shared_ptr<base::State>) override { throw std::logic_error("stub: CPU path not available"); }
    bool isValid(const std::shared_ptr<base::State>, const std::shared_ptr<base::State>) override { throw std::logic_error("stub: CPU path not available"); }
    float computeDistance(const std::shared_ptr<base::State>, bool) override { throw std::logic_error("stub: CPU path not available"); }
    float computeDistanceUnderestimation(const std::shared_ptr<base::State>, const std::shared_ptr<std::vector<Eigen::MatrixXf>>) override {
        throw std::logic_error("stub: CPU path not available");
    }
};
}  // namespace base
