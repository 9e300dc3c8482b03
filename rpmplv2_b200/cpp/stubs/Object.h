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
    inline const fcl::Vector3f &getPosition() const { return position; }
    inline const fcl::Vector3f &getVelocity() const { return velocity; }
    inline float getMaxVel() const { return max_vel; }
    // the reference: position = p; coll_object->setTranslation(p); coll_object->computeAABB()  (Object.cpp:21-26)
    void setPosition(const fcl::Vector3f &p) {
        fcl::AABB a = coll_object->getAABB();
        for (int k = 0; k < 3; k++) { const float half = 0.5f * (a.max_[k] - a.min_[k]); a.min_[k] = p[k] + (-half); a.max_[k] = p[k] + half; }
        coll_object->setAABB(a);
        position = p;
    }
    inline void setVelocity(const fcl::Vector3f &v) { velocity = v; }
    inline void setMaxVel(float v) { max_vel = v; }
    inline void setLabel(const std::string &label_) { label = label_; }
    inline void setCollObject(const std::shared_ptr<fcl::CollisionObject<float>> c) { coll_object = c; }
    inline void setMinDistTol(float t) { min_dist_tol = t; }
protected:
    std::string label;
    std::shared_ptr<fcl::CollisionObject<float>> coll_object;
    float min_dist_tol = RealVectorSpaceConfig::MAX_DISTANCE;
    fcl::Vector3f position, velocity;
    float max_vel = 0;
};
class Box : public Object {};
}  // namespace env
