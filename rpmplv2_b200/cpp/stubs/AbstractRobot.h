// Stand-in for include/robots/AbstractRobot.h: the getters RealVectorSpaceCUDA reads.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "State.h"
namespace robots {
class AbstractRobot {
public:
    virtual ~AbstractRobot() = default;
    inline const std::string &getType() const { return type; }
    inline size_t getNumDOFs() const { return num_DOFs; }
    inline size_t getNumLinks() const { return num_links; }
    inline const std::vector<std::pair<float, float>> &getLimits() const { return limits; }
    inline float getCapsuleRadius(size_t idx) const { return capsules_radius[idx]; }
    inline float getGripperLength() const { return gripper_length; }
    inline size_t getGroundIncluded() const { return ground_included; }
    std::string type;
    size_t num_DOFs = 0, num_links = 0;
    std::vector<std::pair<float, float>> limits;
    std::vector<float> capsules_radius;
    float gripper_length = 0;
    size_t ground_included = 0;
};
}  // namespace robots
