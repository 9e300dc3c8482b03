// oracle/env_shim/urdf/model.h — work-alike of the slice of urdfdom (urdf::Model) that RPMPLv2's robot classes and kdl_parser read:
// links (visual / collision geometry + origin), joints (type, origin, axis, limits), the link tree.  TEST INFRASTRUCTURE ONLY.
// urdfdom is not in /root/reference; restated from its published behaviour:
//   * <origin rpy> is stored as a QUATERNION (urdf::Rotation::setFromRPY with half angles, then normalize()); a missing attribute
//     means identity / zero;   * <axis> defaults to (1, 0, 0);   * Model::getLinks() walks a std::map, i.e. links come back SORTED BY NAME
//     (the reference relies on it: xArm6.cpp:31 indexes links_[i] for i < num_DOFs and gets link1..link6);
//   * only the top-level <link> / <joint> children of <robot> count (<transmission> blocks hold <joint> tags of their own).
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace urdf {

struct Vector3 { double x = 0, y = 0, z = 0; };
struct Rotation {
    double x = 0, y = 0, z = 0, w = 1;
    void normalize() {
        const double s = std::sqrt(x * x + y * y + z * z + w * w);
        if (s == 0.0) { x = 0; y = 0; z = 0; w = 1; } else { x /= s; y /= s; z /= s; w /= s; }
    }
    void setFromRPY(double roll, double pitch, double yaw) {
        const double phi = roll / 2.0, the = pitch / 2.0, psi = yaw / 2.0;
        x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
        y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
        z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
        w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
        normalize();
    }
};
struct Pose { Vector3 position; Rotation rotation; };

struct Geometry {
    enum { SPHERE, BOX, CYLINDER, MESH } type = SPHERE;
    virtual ~Geometry() {}
};
struct Sphere : Geometry { double radius = 0; Sphere() { type = SPHERE; } };
struct Box : Geometry { Vector3 dim; Box() { type = BOX; } };
struct Cylinder : Geometry { double length = 0, radius = 0; Cylinder() { type = CYLINDER; } };
struct Mesh : Geometry { std::string filename; Vector3 scale; Mesh() { type = MESH; scale.x = scale.y = scale.z = 1; } };
typedef std::shared_ptr<Geometry> GeometrySharedPtr;

struct Visual { Pose origin; GeometrySharedPtr geometry; };
struct Collision { Pose origin; GeometrySharedPtr geometry; };
struct Inertial { Pose origin; double mass = 0; };
struct JointLimits { double lower = 0, upper = 0, effort = 0, velocity = 0; };

struct Joint {
    enum { UNKNOWN, REVOLUTE, CONTINUOUS, PRISMATIC, FLOATING, PLANAR, FIXED } type = UNKNOWN;
    std::string name, parent_link_name, child_link_name;
    Vector3 axis;
    Pose parent_to_joint_origin_transform;
    std::shared_ptr<JointLimits> limits;
    Joint() { axis.x = 1; }
};
typedef std::shared_ptr<Joint> JointSharedPtr;
typedef std::shared_ptr<const Joint> JointConstSharedPtr;

struct Link;
typedef std::shared_ptr<Link> LinkSharedPtr;
typedef std::shared_ptr<const Link> LinkConstSharedPtr;
struct Link {
    std::string name;
    std::shared_ptr<Inertial> inertial;
    std::shared_ptr<Visual> visual;
    std::shared_ptr<Collision> collision;
    JointSharedPtr parent_joint;
    std::vector<JointSharedPtr> child_joints;
    std::vector<LinkSharedPtr> child_links;
};

namespace detail {
// a minimal XML element tree: tags, attributes, comments / declarations skipped, text ignored
struct Xml {
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<Xml> kids;
    const Xml *child(const std::string &t) const { for (const auto &k : kids) if (k.tag == t) return &k; return nullptr; }
    bool has(const std::string &a) const { return attr.count(a) != 0; }
    std::string get(const std::string &a) const { const auto it = attr.find(a); return it == attr.end() ? std::string() : it->second; }
};
inline void skip_ws(const std::string &s, size_t &i) { while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) i++; }
inline bool parse_element(const std::string &s, size_t &i, Xml &out) {
    for (;;) {                                                 // skip text, comments, declarations up to the next start tag
        const size_t lt = s.find('<', i);
        if (lt == std::string::npos) return false;
        if (s.compare(lt, 4, "<!--") == 0) { const size_t e = s.find("-->", lt); if (e == std::string::npos) return false; i = e + 3; continue; }
        if (s.compare(lt, 2, "<?") == 0) { const size_t e = s.find("?>", lt); if (e == std::string::npos) return false; i = e + 2; continue; }
        if (s.compare(lt, 2, "<!") == 0) { const size_t e = s.find('>', lt); if (e == std::string::npos) return false; i = e + 1; continue; }
        if (s.compare(lt, 2, "</") == 0) { i = lt; return false; }
        i = lt + 1;
        break;
    }
    size_t b = i;
    while (i < s.size() && !std::isspace(static_cast<unsigned char>(s[i])) && s[i] != '>' && s[i] != '/') i++;
    out.tag = s.substr(b, i - b);
    for (;;) {
        skip_ws(s, i);
        if (i >= s.size()) return false;
        if (s[i] == '/') { i = s.find('>', i) + 1; return true; }      // <tag ... />
        if (s[i] == '>') { i++; break; }
        b = i;
        while (i < s.size() && s[i] != '=' && !std::isspace(static_cast<unsigned char>(s[i]))) i++;
        const std::string key = s.substr(b, i - b);
        skip_ws(s, i);
        if (s[i] != '=') return false;
        i++;
        skip_ws(s, i);
        const char quote = s[i++];
        b = i;
        while (i < s.size() && s[i] != quote) i++;
        out.attr[key] = s.substr(b, i - b);
        i++;
    }
    for (;;) {                                                 // children until the matching end tag
        Xml kid;
        if (parse_element(s, i, kid)) { out.kids.push_back(kid); continue; }
        if (i < s.size() && s.compare(i, 2, "</") == 0) { i = s.find('>', i) + 1; return true; }
        return false;
    }
}
inline bool numbers(const std::string &text, double *out, int count) {
    std::istringstream ss(text);
    std::string tok;
    for (int k = 0; k < count; k++) { if (!(ss >> tok)) return false; out[k] = std::strtod(tok.c_str(), nullptr); }
    return true;
}
inline Pose parse_pose(const Xml *o) {
    Pose p;
    if (!o) return p;
    double v[3];
    if (o->has("xyz") && numbers(o->get("xyz"), v, 3)) { p.position.x = v[0]; p.position.y = v[1]; p.position.z = v[2]; }
    if (o->has("rpy") && numbers(o->get("rpy"), v, 3)) p.rotation.setFromRPY(v[0], v[1], v[2]);
    return p;
}
inline GeometrySharedPtr parse_geometry(const Xml *g) {
    if (!g || g->kids.empty()) return nullptr;
    const Xml &s = g->kids[0];
    double v[3];
    if (s.tag == "box") { auto b = std::make_shared<Box>(); if (numbers(s.get("size"), v, 3)) { b->dim.x = v[0]; b->dim.y = v[1]; b->dim.z = v[2]; } return b; }
    if (s.tag == "cylinder") { auto c = std::make_shared<Cylinder>(); c->length = std::strtod(s.get("length").c_str(), nullptr); c->radius = std::strtod(s.get("radius").c_str(), nullptr); return c; }
    if (s.tag == "sphere") { auto c = std::make_shared<Sphere>(); c->radius = std::strtod(s.get("radius").c_str(), nullptr); return c; }
    if (s.tag == "mesh") { auto m = std::make_shared<Mesh>(); m->filename = s.get("filename"); return m; }
    return nullptr;
}
}  // namespace detail

class Model {
public:
    std::map<std::string, LinkSharedPtr> links_;
    std::map<std::string, JointSharedPtr> joints_;
    LinkSharedPtr root_link_;
    std::string name_;

    bool initFile(const std::string &path) {
        std::ifstream in(path);
        if (!in) return false;
        std::stringstream buf;
        buf << in.rdbuf();
        return initString(buf.str());
    }
    bool initString(const std::string &text) {
        detail::Xml robot;
        size_t i = 0;
        if (!detail::parse_element(text, i, robot) || robot.tag != "robot") return false;
        name_ = robot.get("name");
        links_.clear(); joints_.clear(); root_link_.reset();
        for (const auto &e : robot.kids) {
            if (e.tag == "link") {
                auto l = std::make_shared<Link>();
                l->name = e.get("name");
                if (const auto *v = e.child("visual")) { l->visual = std::make_shared<Visual>(); l->visual->origin = detail::parse_pose(v->child("origin")); l->visual->geometry = detail::parse_geometry(v->child("geometry")); }
                if (const auto *c = e.child("collision")) { l->collision = std::make_shared<Collision>(); l->collision->origin = detail::parse_pose(c->child("origin")); l->collision->geometry = detail::parse_geometry(c->child("geometry")); }
                links_[l->name] = l;
            } else if (e.tag == "joint") {
                auto j = std::make_shared<Joint>();
                j->name = e.get("name");
                const std::string t = e.get("type");
                j->type = t == "revolute" ? Joint::REVOLUTE : t == "continuous" ? Joint::CONTINUOUS : t == "prismatic" ? Joint::PRISMATIC : t == "fixed" ? Joint::FIXED : Joint::UNKNOWN;
                j->parent_to_joint_origin_transform = detail::parse_pose(e.child("origin"));
                if (const auto *p = e.child("parent")) j->parent_link_name = p->get("link");
                if (const auto *c = e.child("child")) j->child_link_name = c->get("link");
                double v[3];
                if (const auto *a = e.child("axis")) if (detail::numbers(a->get("xyz"), v, 3)) { j->axis.x = v[0]; j->axis.y = v[1]; j->axis.z = v[2]; }
                if (const auto *lim = e.child("limit")) {
                    j->limits = std::make_shared<JointLimits>();
                    j->limits->lower = std::strtod(lim->get("lower").c_str(), nullptr);
                    j->limits->upper = std::strtod(lim->get("upper").c_str(), nullptr);
                    j->limits->effort = std::strtod(lim->get("effort").c_str(), nullptr);
                    j->limits->velocity = std::strtod(lim->get("velocity").c_str(), nullptr);
                }
                joints_[j->name] = j;
            }
        }
        // the tree (urdfdom's initTree / initRoot): children are attached in joint-name order
        std::map<std::string, bool> has_parent;
        for (auto &kv : joints_) {
            const auto p = links_.find(kv.second->parent_link_name), c = links_.find(kv.second->child_link_name);
            if (p == links_.end() || c == links_.end()) return false;
            c->second->parent_joint = kv.second;
            p->second->child_joints.push_back(kv.second);
            p->second->child_links.push_back(c->second);
            has_parent[c->first] = true;
        }
        for (auto &kv : links_)
            if (!has_parent.count(kv.first)) { if (root_link_) return false; root_link_ = kv.second; }
        return bool(root_link_);
    }
    const std::string &getName() const { return name_; }
    LinkConstSharedPtr getRoot() const { return root_link_; }
    void getLinks(std::vector<LinkSharedPtr> &out) const { for (const auto &kv : links_) out.push_back(kv.second); }
    JointConstSharedPtr getJoint(const std::string &name) const { const auto it = joints_.find(name); return it == joints_.end() ? nullptr : it->second; }
    LinkConstSharedPtr getLink(const std::string &name) const { const auto it = links_.find(name); return it == links_.end() ? nullptr : it->second; }
};

}  // namespace urdf
