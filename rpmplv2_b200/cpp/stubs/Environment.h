// Stand-in for include/environments/Environment.h.
#pragma once
#include <memory>
#include <vector>
#include "Object.h"
namespace env {
class Environment {
public:
    inline const std::vector<std::shared_ptr<env::Object>> &getObjects() const { return objects; }
    inline std::shared_ptr<env::Object> getObject(size_t idx) const { return objects[idx]; }
    inline std::shared_ptr<fcl::CollisionObjectf> getCollObject(size_t idx) const { return objects[idx]->getCollObject(); }
    inline size_t getNumObjects() const { return objects.size(); }
    void addObject(const std::shared_ptr<env::Object> object) { objects.push_back(object); }
    inline void setBaseRadius(float r) { base_radius = r; }
    inline void setRobotMaxVel(float v) { robot_max_vel = v; }
    inline const fcl::Vector3f &getWSCenter() const { return WS_center; }
    inline float getWSRadius() const { return WS_radius; }
    inline float getBaseRadius() const { return base_radius; }
    inline float getRobotMaxVel() const { return robot_max_vel; }
    // stub only: the reference reads these from the scenario YAML in its constructor (Environment.cpp:73-79)
    void stubSetWorkspace(const fcl::Vector3f &c, float r) { WS_center = c; WS_radius = r; }
private:
    std::vector<std::shared_ptr<env::Object>> objects;
    fcl::Vector3f WS_center;
    float WS_radius = 0, base_radius = 0, robot_max_vel = 0;
};
}  // namespace env
