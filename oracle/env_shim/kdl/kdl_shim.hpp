// oracle/env_shim/kdl/kdl_shim.hpp — a work-alike of the slice of orocos-KDL that RPMPLv2's robot classes use
// (src/robots/xArm6.cpp, Planar2DOF.cpp: kdl_parser::treeFromFile -> KDL::Tree -> getChain -> TreeFkSolverPos_recursive::JntToCart,
// Frame / Rotation / Vector arithmetic).  TEST INFRASTRUCTURE ONLY: with it on the include path the reference's own robot sources
// compile UNMODIFIED from where they lie (oracle/Makefile, target `ref`), which pins the skeleton recipe (xArm6.cpp:209-235,
// Planar2DOF.cpp:117-130), computeEnclosingRadii (xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30) and the capsule
// stage of xArm6::checkSelfCollision (xArm6.cpp:284-424) to the real source.
//
// orocos-KDL is NOT in /root/reference (apt liborocos-kdl-dev on the reference's CI image, no version pin).  What is restated here is
// KDL's published algorithm, operation for operation as in its sources (frames.inl / frames.cpp / joint.cpp / segment.cpp / tree.cpp /
// treefksolverpos_recursive.cpp), all in double:
//   Rotation * Rotation, Rotation * Vector : three-term sums, left to right;   Frame * Frame = (M1 * M2, M1 * p2 + p1)
//   Rotation::Rot2(axis, angle)            : ct + vt*x*x, -z*st + vt*x*y, ...  (Rodrigues, KDL's operand order)
//   Rotation::Quaternion / RPY             : KDL's closed forms
//   Joint(origin, axis, RotAxis)::pose(q)  = Frame(Rot2(axis / |axis|, q), origin)
//   Segment(joint, f_tip)::pose(q)         = joint.pose(q) * (joint.pose(0).Inverse() * f_tip)
//   TreeFkSolverPos_recursive::JntToCart   = pose(root) * ... * pose(segment), composed from the root down (left associated)
// The inverse-kinematics solvers are declared and refuse to run (xArm6::computeInverseKinematics is not on the path).
// FK ARITHMETIC THEREFORE STILL RESTS ON A RESTATEMENT (of KDL, not of the reference): DESIGN.md keeps "KDL arithmetic" on the
// parity-unpinned list; everything the reference itself computes from the frames is pinned by its own code.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace KDL {

inline double sqr(double x) { return x * x; }

class Vector {
public:
    double data[3];
    Vector() : data{0, 0, 0} {}
    Vector(double x, double y, double z) : data{x, y, z} {}
    double x() const { return data[0]; }
    double y() const { return data[1]; }
    double z() const { return data[2]; }
    double operator()(int i) const { return data[i]; }
    double &operator()(int i) { return data[i]; }
    double operator[](int i) const { return data[i]; }
    double &operator[](int i) { return data[i]; }
    Vector &operator+=(const Vector &a) { data[0] += a.data[0]; data[1] += a.data[1]; data[2] += a.data[2]; return *this; }
    Vector &operator-=(const Vector &a) { data[0] -= a.data[0]; data[1] -= a.data[1]; data[2] -= a.data[2]; return *this; }
    static Vector Zero() { return Vector(0, 0, 0); }
    // frames.cpp: scaled norm (divides by the largest component before the square root)
    double Norm(double eps = 1e-6) const {
        double t1 = std::fabs(data[0]), t2 = std::fabs(data[1]);
        if (t1 >= t2) {
            t2 = std::fabs(data[2]);
            if (t1 >= t2) {
                if (t1 < eps) return 0;
                return t1 * std::sqrt(1 + sqr(data[1] / data[0]) + sqr(data[2] / data[0]));
            }
            return t2 * std::sqrt(1 + sqr(data[0] / data[2]) + sqr(data[1] / data[2]));
        }
        t1 = std::fabs(data[2]);
        if (t2 > t1) return t2 * std::sqrt(1 + sqr(data[0] / data[1]) + sqr(data[2] / data[1]));
        return t1 * std::sqrt(1 + sqr(data[0] / data[2]) + sqr(data[1] / data[2]));
    }
};
inline Vector operator+(const Vector &a, const Vector &b) { return Vector(a.data[0] + b.data[0], a.data[1] + b.data[1], a.data[2] + b.data[2]); }
inline Vector operator-(const Vector &a, const Vector &b) { return Vector(a.data[0] - b.data[0], a.data[1] - b.data[1], a.data[2] - b.data[2]); }
inline Vector operator-(const Vector &a) { return Vector(-a.data[0], -a.data[1], -a.data[2]); }
inline Vector operator*(const Vector &a, double s) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
inline Vector operator*(double s, const Vector &a) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
inline Vector operator/(const Vector &a, double s) { return Vector(a.data[0] / s, a.data[1] / s, a.data[2] / s); }

class Rotation {
public:
    double data[9];                                            // row-major
    Rotation() : data{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
    Rotation(double xx, double yx, double zx, double xy, double yy, double zy, double xz, double yz, double zz) : data{xx, yx, zx, xy, yy, zy, xz, yz, zz} {}
    static Rotation Identity() { return Rotation(); }
    double operator()(int i, int j) const { return data[3 * i + j]; }
    double &operator()(int i, int j) { return data[3 * i + j]; }
    Vector UnitX() const { return Vector(data[0], data[3], data[6]); }
    Vector UnitY() const { return Vector(data[1], data[4], data[7]); }
    Vector UnitZ() const { return Vector(data[2], data[5], data[8]); }
    Vector operator*(const Vector &v) const {
        return Vector(data[0] * v.data[0] + data[1] * v.data[1] + data[2] * v.data[2],
                      data[3] * v.data[0] + data[4] * v.data[1] + data[5] * v.data[2],
                      data[6] * v.data[0] + data[7] * v.data[1] + data[8] * v.data[2]);
    }
    Rotation Inverse() const { return Rotation(data[0], data[3], data[6], data[1], data[4], data[7], data[2], data[5], data[8]); }
    Vector Inverse(const Vector &v) const {
        return Vector(data[0] * v.data[0] + data[3] * v.data[1] + data[6] * v.data[2],
                      data[1] * v.data[0] + data[4] * v.data[1] + data[7] * v.data[2],
                      data[2] * v.data[0] + data[5] * v.data[1] + data[8] * v.data[2]);
    }
    // frames.cpp Rotation::Rot2: rotation about a UNIT vector
    static Rotation Rot2(const Vector &v, double angle) {
        const double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
        const double m_vt_0 = vt * v(0), m_vt_1 = vt * v(1), m_vt_2 = vt * v(2);
        const double m_st_0 = v(0) * st, m_st_1 = v(1) * st, m_st_2 = v(2) * st;
        const double m_vt_0_1 = m_vt_0 * v(1), m_vt_0_2 = m_vt_0 * v(2), m_vt_1_2 = m_vt_1 * v(2);
        return Rotation(ct + m_vt_0 * v(0), -m_st_2 + m_vt_0_1, m_st_1 + m_vt_0_2,
                        m_st_2 + m_vt_0_1, ct + m_vt_1 * v(1), -m_st_0 + m_vt_1_2,
                        -m_st_1 + m_vt_0_2, m_st_0 + m_vt_1_2, ct + m_vt_2 * v(2));
    }
    static Rotation Quaternion(double x, double y, double z, double w) {
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        return Rotation(w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
                        2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2);
    }
    // (xArm6.cpp:39 calls it through an instance: link_frame.M.RPY(...))
    static Rotation RPY(double roll, double pitch, double yaw) {
        const double ca1 = std::cos(yaw), sa1 = std::sin(yaw), cb1 = std::cos(pitch), sb1 = std::sin(pitch), cc1 = std::cos(roll), sc1 = std::sin(roll);
        return Rotation(ca1 * cb1, ca1 * sb1 * sc1 - sa1 * cc1, ca1 * sb1 * cc1 + sa1 * sc1,
                        sa1 * cb1, sa1 * sb1 * sc1 + ca1 * cc1, sa1 * sb1 * cc1 - ca1 * sc1,
                        -sb1, cb1 * sc1, cb1 * cc1);
    }
    // only feeds FCL link transforms (KDL2fcl), off the capsule path
    void GetQuaternion(double &x, double &y, double &z, double &w) const {
        const double trace = data[0] + data[4] + data[8];
        if (trace > 1e-12) {
            const double s = 0.5 / std::sqrt(trace + 1.0);
            w = 0.25 / s; x = (data[7] - data[5]) * s; y = (data[2] - data[6]) * s; z = (data[3] - data[1]) * s;
        } else if (data[0] > data[4] && data[0] > data[8]) {
            const double s = 2.0 * std::sqrt(1.0 + data[0] - data[4] - data[8]);
            w = (data[7] - data[5]) / s; x = 0.25 * s; y = (data[1] + data[3]) / s; z = (data[2] + data[6]) / s;
        } else if (data[4] > data[8]) {
            const double s = 2.0 * std::sqrt(1.0 + data[4] - data[0] - data[8]);
            w = (data[2] - data[6]) / s; x = (data[1] + data[3]) / s; y = 0.25 * s; z = (data[5] + data[7]) / s;
        } else {
            const double s = 2.0 * std::sqrt(1.0 + data[8] - data[0] - data[4]);
            w = (data[3] - data[1]) / s; x = (data[2] + data[6]) / s; y = (data[5] + data[7]) / s; z = 0.25 * s;
        }
    }
};
inline Rotation operator*(const Rotation &a, const Rotation &b) {
    return Rotation(a.data[0] * b.data[0] + a.data[1] * b.data[3] + a.data[2] * b.data[6], a.data[0] * b.data[1] + a.data[1] * b.data[4] + a.data[2] * b.data[7],
                    a.data[0] * b.data[2] + a.data[1] * b.data[5] + a.data[2] * b.data[8], a.data[3] * b.data[0] + a.data[4] * b.data[3] + a.data[5] * b.data[6],
                    a.data[3] * b.data[1] + a.data[4] * b.data[4] + a.data[5] * b.data[7], a.data[3] * b.data[2] + a.data[4] * b.data[5] + a.data[5] * b.data[8],
                    a.data[6] * b.data[0] + a.data[7] * b.data[3] + a.data[8] * b.data[6], a.data[6] * b.data[1] + a.data[7] * b.data[4] + a.data[8] * b.data[7],
                    a.data[6] * b.data[2] + a.data[7] * b.data[5] + a.data[8] * b.data[8]);
}

class Frame {
public:
    Vector p;
    Rotation M;
    Frame() {}
    Frame(const Rotation &R, const Vector &V) : p(V), M(R) {}
    explicit Frame(const Vector &V) : p(V) {}
    explicit Frame(const Rotation &R) : M(R) {}
    static Frame Identity() { return Frame(); }
    Frame Inverse() const { return Frame(M.Inverse(), -M.Inverse(p)); }
    Vector operator*(const Vector &v) const { return M * v + p; }
};
inline Frame operator*(const Frame &a, const Frame &b) { return Frame(a.M * b.M, a.M * b.p + a.p); }

inline std::ostream &operator<<(std::ostream &os, const Vector &v) { return os << "[" << v(0) << "," << v(1) << "," << v(2) << "]"; }
inline std::ostream &operator<<(std::ostream &os, const Rotation &R) {
    os << "[";
    for (int i = 0; i < 3; i++) os << R(i, 0) << "," << R(i, 1) << "," << R(i, 2) << (i < 2 ? ";\n " : "]");
    return os;
}
inline std::ostream &operator<<(std::ostream &os, const Frame &T) { return os << "[" << T.M << "\n" << T.p << "]"; }

class JntArray {
public:
    std::vector<double> data_;
    struct DataView {                                          // q_in.data(i) (Eigen::VectorXd in KDL)
        std::vector<double> *v;
        double &operator()(size_t i) { return (*v)[i]; }
        double operator()(size_t i) const { return (*v)[i]; }
    } data{&data_};
    JntArray() {}
    explicit JntArray(unsigned n) : data_(n, 0.0) {}
    JntArray(const JntArray &o) : data_(o.data_), data{&data_} {}
    JntArray &operator=(const JntArray &o) { data_ = o.data_; return *this; }
    double operator()(unsigned i) const { return data_[i]; }
    double &operator()(unsigned i) { return data_[i]; }
    unsigned rows() const { return unsigned(data_.size()); }
};

class Joint {
public:
    enum JointType { RotAxis, RotX, RotY, RotZ, TransAxis, TransX, TransY, TransZ, Fixed, None = Fixed };
    Joint() : type_(None) {}
    Joint(const std::string &name, JointType type) : name_(name), type_(type) {}
    Joint(const std::string &name, const Vector &origin, const Vector &axis, JointType type)
        : name_(name), type_(type), axis_(axis / axis.Norm()), origin_(origin) {
        if (type != RotAxis) throw std::runtime_error("KDL work-alike: only RotAxis joints carry an axis");
        joint_pose_.p = origin_;
        joint_pose_.M = Rotation::Rot2(axis_, 0.0);
        q_previous_ = 0;
    }
    Frame pose(double q) const {
        if (type_ == RotAxis) {
            if (q != q_previous_) {
                q_previous_ = q;
                joint_pose_.M = Rotation::Rot2(axis_, 1.0 * q + 0.0);      // scale * q + offset
            }
            return joint_pose_;
        }
        return Frame::Identity();
    }
    JointType getType() const { return type_; }
    const std::string &getName() const { return name_; }

private:
    std::string name_;
    JointType type_;
    Vector axis_, origin_;
    mutable Frame joint_pose_;
    mutable double q_previous_ = 0;
};

class Segment {
public:
    Segment() {}
    explicit Segment(const std::string &name, const Joint &joint = Joint(), const Frame &f_tip = Frame::Identity())
        : name_(name), joint_(joint), f_tip_(joint.pose(0).Inverse() * f_tip) {}
    Frame pose(double q) const { return joint_.pose(q) * f_tip_; }
    const std::string &getName() const { return name_; }
    const Joint &getJoint() const { return joint_; }

private:
    std::string name_;
    Joint joint_;
    Frame f_tip_;
};

class Chain {
public:
    std::vector<Segment> segments;
    void addSegment(const Segment &s) { segments.push_back(s); if (s.getJoint().getType() != Joint::None) nr_joints_++; }
    unsigned getNrOfJoints() const { return nr_joints_; }
    unsigned getNrOfSegments() const { return unsigned(segments.size()); }
    const Segment &getSegment(unsigned i) const { return segments[i]; }
    void clear() { segments.clear(); nr_joints_ = 0; }

private:
    unsigned nr_joints_ = 0;
};

struct TreeElement {
    Segment segment;
    unsigned q_nr = 0;
    std::string parent;                                        // "" for the root
};
typedef std::map<std::string, TreeElement> SegmentMap;

class Tree {
public:
    explicit Tree(const std::string &root_name = "root") : root_name_(root_name) { segments_[root_name].segment = Segment(root_name); }
    bool addSegment(const Segment &s, const std::string &hook_name) {
        if (!segments_.count(hook_name) || segments_.count(s.getName())) return false;
        TreeElement e;
        e.segment = s;
        e.parent = hook_name;
        e.q_nr = s.getJoint().getType() != Joint::None ? nr_joints_ : 0;
        segments_[s.getName()] = e;
        nr_segments_++;
        if (s.getJoint().getType() != Joint::None) nr_joints_++;
        return true;
    }
    unsigned getNrOfJoints() const { return nr_joints_; }
    unsigned getNrOfSegments() const { return nr_segments_; }
    const SegmentMap &getSegments() const { return segments_; }
    const std::string &getRootName() const { return root_name_; }
    // chain_root must be an ancestor of chain_tip (all the reference asks for)
    bool getChain(const std::string &chain_root, const std::string &chain_tip, Chain &chain) const {
        chain.clear();
        if (!segments_.count(chain_root) || !segments_.count(chain_tip)) return false;
        std::vector<const TreeElement *> up;
        for (std::string s = chain_tip; s != chain_root;) {
            const auto it = segments_.find(s);
            if (it == segments_.end() || it->second.parent.empty()) return false;
            up.push_back(&it->second);
            s = it->second.parent;
        }
        for (auto it = up.rbegin(); it != up.rend(); ++it) chain.addSegment((*it)->segment);
        return true;
    }

private:
    SegmentMap segments_;
    std::string root_name_;
    unsigned nr_joints_ = 0, nr_segments_ = 0;
};

class TreeFkSolverPos_recursive {
public:
    explicit TreeFkSolverPos_recursive(const Tree &tree) : tree_(tree) {}
    int JntToCart(const JntArray &q_in, Frame &p_out, const std::string &segment_name) {
        const auto it = tree_.getSegments().find(segment_name);
        if (q_in.rows() != tree_.getNrOfJoints()) return -1;
        if (it == tree_.getSegments().end()) return -2;
        p_out = recursiveFk(q_in, it);
        return 0;
    }

private:
    Frame recursiveFk(const JntArray &q_in, const SegmentMap::const_iterator &it) const {
        const TreeElement &e = it->second;
        const Frame current = e.segment.pose(q_in(e.q_nr));
        if (e.parent.empty()) return current;
        return recursiveFk(q_in, tree_.getSegments().find(e.parent)) * current;
    }
    const Tree &tree_;
};

// inverse kinematics (xArm6::computeInverseKinematics): declared so that the source compiles, refuses to run
class ChainFkSolverPos_recursive { public: explicit ChainFkSolverPos_recursive(const Chain &) {} };
class ChainIkSolverVel_pinv { public: explicit ChainIkSolverVel_pinv(const Chain &) {} };
class ChainIkSolverPos_NR {
public:
    ChainIkSolverPos_NR(const Chain &, ChainFkSolverPos_recursive &, ChainIkSolverVel_pinv &, unsigned = 100, double = 1e-6) {}
    int CartToJnt(const JntArray &, const Frame &, JntArray &) { throw std::runtime_error("KDL work-alike: inverse kinematics is not restated"); }
    int getError() const { return 0; }
};

}  // namespace KDL
