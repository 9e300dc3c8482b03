// Stand-in for include/state_spaces/State.h: the per-configuration cache fields the hot path reads and writes.
#pragma once
#include <eigen3/Eigen/Dense>
#include <memory>
#include <vector>
#include "StateSpaceType.h"
namespace base {
class State {
public:
    enum class Status { None, Advanced, Trapped, Reached, Orphan };
    explicit State(const Eigen::VectorXf &coord_) : num_dimensions(size_t(coord_.size())), coord(coord_) {}
    virtual ~State() = default;
    inline size_t getNumDimensions() const { return num_dimensions; }
    inline const Eigen::VectorXf &getCoord() const { return coord; }
    inline float getCoord(size_t idx) const { return coord(Eigen::Index(idx)); }
    inline float getDistance() const { return d_c; }
    inline const std::vector<float> &getDistanceProfile() const { return d_c_profile; }
    inline bool getIsRealDistance() const { return is_real_d_c; }
    inline std::shared_ptr<std::vector<Eigen::MatrixXf>> getNearestPoints() const { return nearest_points; }
    inline void setDistance(float d_c_) { d_c = d_c_; }
    inline void setDistanceProfile(const std::vector<float> &d_c_profile_) { d_c_profile = d_c_profile_; }
    inline void setIsRealDistance(bool is_real_d_c_) { is_real_d_c = is_real_d_c_; }
    inline void setNearestPoints(const std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points_) { nearest_points = nearest_points_; }
protected:
    size_t num_dimensions;
    Eigen::VectorXf coord;
    float d_c = -1;                       // State.cpp:10
    std::vector<float> d_c_profile;
    bool is_real_d_c = true;
    std::shared_ptr<std::vector<Eigen::MatrixXf>> nearest_points;
};
class RealVectorSpaceState : public State {
public:
    using State::State;
};
}  // namespace base
