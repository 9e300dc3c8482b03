// oracle/env_shim/fcl/fcl.h — stand-in for the slice of FCL that RPMPLv2's env::Object / env::Box / env::Environment touch, so that the
// reference's own Environment.cpp, Object.cpp and Box.cpp compile UNMODIFIED from where they lie (oracle/Makefile, target `ref`) and
// pin the motion-law logic of oracle/rvs_oracle_env.hpp.  TEST INFRASTRUCTURE ONLY.  computeAABB follows FCL's published behaviour
// (SURVEY.md Q22): identity rotation -> local AABB (+-side/2) translated; otherwise the bounding-sphere cube around the centre.
#pragma once
#include <cmath>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include <Eigen/Dense>

namespace fcl {
using Vector3f = Eigen::Vector3f;
using Matrix3f = Eigen::Matrix3f;
struct Quaternionf {
    float w_, x_, y_, z_;
    Quaternionf(float w, float x, float y, float z) : w_(w), x_(x), y_(y), z_(z) {}
    static Quaternionf Identity() { return Quaternionf(1, 0, 0, 0); }
    Matrix3f matrix() const {
        Matrix3f m;
        const float tx = 2 * x_, ty = 2 * y_, tz = 2 * z_, twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_,
                    tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        m(0, 0) = 1 - (tyy + tzz); m(0, 1) = txy - twz; m(0, 2) = txz + twy;
        m(1, 0) = txy + twz; m(1, 1) = 1 - (txx + tzz); m(1, 2) = tyz - twx;
        m(2, 0) = txz - twy; m(2, 1) = tyz + twx; m(2, 2) = 1 - (txx + tyy);
        return m;
    }
};
template <class S> struct AABB { Vector3f min_, max_; };
enum NODE_TYPE { BV_UNKNOWN, GEOM_BOX, GEOM_SPHERE };
template <class S> struct CollisionGeometry {
    virtual ~CollisionGeometry() {}
    virtual NODE_TYPE getNodeType() const = 0;
    Vector3f half;            // local AABB half extents
    float radius = 0;         // local AABB radius
};
template <class S> struct Box : CollisionGeometry<S> {
    Vector3f side;
    explicit Box(const Vector3f &s) : side(s) {
        this->half = 0.5f * s;
        this->radius = this->half.norm();
    }
    NODE_TYPE getNodeType() const override { return GEOM_BOX; }
};
template <class S> class CollisionObject {
public:
    CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Matrix3f &R, const Vector3f &T) : geom(g), rot(R), pos(T) { computeAABB(); }
    void setTranslation(const Vector3f &T) { pos = T; }
    void computeAABB() {
        bool identity = true;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) identity = identity && rot(i, j) == (i == j ? 1.f : 0.f);
        if (identity) { aabb.min_ = pos + (-geom->half); aabb.max_ = pos + geom->half; }
        else { const Vector3f d(geom->radius, geom->radius, geom->radius); aabb.min_ = pos - d; aabb.max_ = pos + d; }
    }
    const AABB<S> &getAABB() const { return aabb; }
    void stubSetAABB(const Vector3f &mn, const Vector3f &mx) { aabb.min_ = mn; aabb.max_ = mx; }      // stand-in only: exact table values for the harness
    NODE_TYPE getNodeType() const { return geom->getNodeType(); }
private:
    std::shared_ptr<CollisionGeometry<S>> geom;
    Matrix3f rot;
    Vector3f pos;
    AABB<S> aabb;
};
using CollisionObjectf = CollisionObject<float>;
}  // namespace fcl
