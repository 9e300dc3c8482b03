// Minimal stand-in for the slice of FCL that RealVectorSpaceCUDA touches: CollisionObjectf::getAABB().min_/max_ and node type.
#pragma once
#include <memory>
namespace fcl {
struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() = default;
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
struct AABB { Vector3f min_, max_; };
enum NODE_TYPE { BV_UNKNOWN, GEOM_BOX, GEOM_SPHERE };
template <class S> class CollisionObject {
public:
    CollisionObject(const AABB &aabb, NODE_TYPE t = GEOM_BOX) : aabb_(aabb), type_(t) {}
    const AABB &getAABB() const { return aabb_; }
    NODE_TYPE getNodeType() const { return type_; }
    void setAABB(const AABB &a) { aabb_ = a; }
private:
    AABB aabb_;
    NODE_TYPE type_;
};
using CollisionObjectf = CollisionObject<float>;
}  // namespace fcl
