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
private:
    std::vector<std::shared_ptr<env::Object>> objects;
};
}  // namespace env
