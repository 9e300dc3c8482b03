// Stand-in for include/environments/Object.h.
#pragma once
#include <fcl/fcl.h>
#include <memory>
#include <string>
#include "RealVectorSpaceConfig.h"
namespace env {
class Object {
public:
    virtual ~Object() = default;
    inline const std::string &getLabel() const { return label; }
    inline std::shared_ptr<fcl::CollisionObject<float>> getCollObject() const { return coll_object; }
    inline float getMinDistTol() const { return min_dist_tol; }
    inline void setLabel(const std::string &label_) { label = label_; }
    inline void setCollObject(const std::shared_ptr<fcl::CollisionObject<float>> c) { coll_object = c; }
    inline void setMinDistTol(float t) { min_dist_tol = t; }
protected:
    std::string label;
    std::shared_ptr<fcl::CollisionObject<float>> coll_object;
    float min_dist_tol = RealVectorSpaceConfig::MAX_DISTANCE;
};
class Box : public Object {};
}  // namespace env
