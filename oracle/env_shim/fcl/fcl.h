// oracle/env_shim/fcl/fcl.h — stand-in for the slice of FCL that RPMPLv2's env::Object / env::Box / env::Environment touch, so that the
// reference's own Environment.cpp, Object.cpp and Box.cpp compile UNMODIFIED from where they lie (oracle/Makefile, target `ref`) and
// pin the motion-law logic of oracle/rvs_oracle_env.hpp.  TEST INFRASTRUCTURE ONLY.  computeAABB follows FCL's published behaviour
// (SURVEY.md Q22): identity rotation -> local AABB (+-side/2) translated; otherwise the bounding-sphere cube around the centre.
// Round 2 adds the slice the reference's ROBOT classes mention (src/robots/xArm6.cpp, Planar2DOF.cpp): Transform3f, BVHModel, Capsule,
// the broad-phase manager and DefaultCollisionFunction.  FCL's mesh narrow phase is third-party arithmetic that cannot be restated:
// the manager's collide() reports a collision UNCONDITIONALLY, so xArm6::checkRealSelfCollision (xArm6.cpp:426-464) answers "true" and
// xArm6::checkSelfCollision reduces to its capsule stage (computeCapsulesDistance < 0) — exactly the conservative contract the CUDA
// engine implements; the control flow around it (pair order, skip_checking, q2 truncation) is the reference's.
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
    explicit Quaternionf(const Matrix3f &m) {                  // (only fcl2KDL, off the path)
        const float t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > 0) { const float s = std::sqrt(t + 1.f) * 2; w_ = 0.25f * s; x_ = (m(2, 1) - m(1, 2)) / s; y_ = (m(0, 2) - m(2, 0)) / s; z_ = (m(1, 0) - m(0, 1)) / s; }
        else if (m(0, 0) > m(1, 1) && m(0, 0) > m(2, 2)) { const float s = std::sqrt(1.f + m(0, 0) - m(1, 1) - m(2, 2)) * 2; w_ = (m(2, 1) - m(1, 2)) / s; x_ = 0.25f * s; y_ = (m(0, 1) + m(1, 0)) / s; z_ = (m(0, 2) + m(2, 0)) / s; }
        else if (m(1, 1) > m(2, 2)) { const float s = std::sqrt(1.f + m(1, 1) - m(0, 0) - m(2, 2)) * 2; w_ = (m(0, 2) - m(2, 0)) / s; x_ = (m(0, 1) + m(1, 0)) / s; y_ = 0.25f * s; z_ = (m(1, 2) + m(2, 1)) / s; }
        else { const float s = std::sqrt(1.f + m(2, 2) - m(0, 0) - m(1, 1)) * 2; w_ = (m(1, 0) - m(0, 1)) / s; x_ = (m(0, 2) + m(2, 0)) / s; y_ = (m(1, 2) + m(2, 1)) / s; z_ = 0.25f * s; }
    }
    float w() const { return w_; }
    float x() const { return x_; }
    float y() const { return y_; }
    float z() const { return z_; }
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
struct Transform3f {                                           // Eigen::Transform<float, 3, Isometry> as far as the robot classes use it
    Matrix3f lin;
    Vector3f trans;
    Transform3f() { for (int i = 0; i < 3; i++) lin(i, i) = 1.f; }
    static Transform3f Identity() { return Transform3f(); }
    Matrix3f &linear() { return lin; }
    const Matrix3f &linear() const { return lin; }
    Vector3f &translation() { return trans; }
    const Vector3f &translation() const { return trans; }
    Matrix3f rotation() const { return lin; }
    Eigen::MatrixXf matrix() const {
        Eigen::MatrixXf m = Eigen::MatrixXf::Identity(4, 4);
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m(i, j) = lin(i, j); m(i, 3) = trans(i); }
        return m;
    }
};
template <class S> struct AABB { Vector3f min_, max_; };
enum NODE_TYPE { BV_UNKNOWN, GEOM_BOX, GEOM_SPHERE, GEOM_CAPSULE, BV_OBBRSS };
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
    Box(S x, S y, S z) : Box(Vector3f(x, y, z)) {}
    NODE_TYPE getNodeType() const override { return GEOM_BOX; }
};
template <class S> struct Capsule : CollisionGeometry<S> {
    Capsule(S r, S l) { this->half = Vector3f(r, r, 0.5f * l + r); this->radius = this->half.norm(); }
    NODE_TYPE getNodeType() const override { return GEOM_CAPSULE; }
};
template <class S> struct OBBRSS {};
template <class BV> struct BVHModel : CollisionGeometry<float> {   // triangles are accepted and dropped: the mesh narrow phase is not restated
    int num_tris = 0;
    void beginModel() {}
    void addTriangle(const Vector3f &, const Vector3f &, const Vector3f &) { num_tris++; }
    void endModel() {}
    NODE_TYPE getNodeType() const override { return BV_OBBRSS; }
};
template <class S> class CollisionObject {
public:
    CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Matrix3f &R, const Vector3f &T) : geom(g), rot(R), pos(T) { computeAABB(); }
    CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Transform3f &tf) : geom(g), rot(tf.linear()), pos(tf.translation()) { computeAABB(); }
    void setTranslation(const Vector3f &T) { pos = T; }
    void setRotation(const Matrix3f &R) { rot = R; }
    void setTransform(const Transform3f &tf) { rot = tf.linear(); pos = tf.translation(); }
    Transform3f getTransform() const { Transform3f tf; tf.linear() = rot; tf.translation() = pos; return tf; }
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
using CollisionGeometryf = CollisionGeometry<float>;
using Boxf = Box<float>;
using Capsulef = Capsule<float>;

// broad phase + default callback: a collide() between two managers that each hold an object reports a collision unconditionally
template <class S> struct CollisionResult { bool hit = false; bool isCollision() const { return hit; } };
template <class S> struct DefaultCollisionData { CollisionResult<S> result; bool done = false; };
template <class S> using CollisionCallBack = bool (*)(CollisionObject<S> *, CollisionObject<S> *, void *);
template <class S> bool DefaultCollisionFunction(CollisionObject<S> *, CollisionObject<S> *, void *data) {
    static_cast<DefaultCollisionData<S> *>(data)->result.hit = true;
    return true;
}
template <class S> class BroadPhaseCollisionManager {
public:
    virtual ~BroadPhaseCollisionManager() {}
    void clear() { objs.clear(); }
    void registerObject(CollisionObject<S> *o) { objs.push_back(o); }
    void setup() {}
    void collide(BroadPhaseCollisionManager *other, void *data, CollisionCallBack<S> cb) const {
        for (auto *a : objs) for (auto *b : other->objs) if (cb(a, b, data)) return;
    }
    std::vector<CollisionObject<S> *> objs;
};
template <class S> class DynamicAABBTreeCollisionManager : public BroadPhaseCollisionManager<S> {};
using BroadPhaseCollisionManagerf = BroadPhaseCollisionManager<float>;
using DynamicAABBTreeCollisionManagerf = DynamicAABBTreeCollisionManager<float>;

// only the robots' private test() helpers mention these
enum GJKSolverType { GST_LIBCCD, GST_INDEP };
template <class S> struct DistanceRequest { DistanceRequest(bool = false, S = 0, S = 0, GJKSolverType = GST_LIBCCD) {} };
template <class S> struct DistanceResult { S min_distance = 0; Vector3f nearest_points[3]; void clear() {} };
template <class S> S distance(const CollisionObject<S> *, const CollisionObject<S> *, const DistanceRequest<S> &, DistanceResult<S> &) {
    throw std::runtime_error("FCL stand-in: distance() is not restated");
}
}  // namespace fcl
