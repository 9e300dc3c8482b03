// oracle/env_shim/kdl_parser — work-alike of kdl_parser::treeFromFile (urdf::Model -> KDL::Tree), restated from kdl_parser's published
// behaviour (kdl_parser.cpp: toKdl(Pose) = Frame(Rotation::Quaternion(x, y, z, w), position); a revolute joint becomes
// Joint(name, F_parent_jnt.p, F_parent_jnt.M * axis, RotAxis), a fixed joint Joint(name, None); every link becomes
// Segment(link name, joint, F_parent_jnt) hooked to its parent link; the tree root is the URDF root link).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../kdl/kdl_shim.hpp"
#include "../urdf/model.h"

namespace kdl_parser {
inline KDL::Vector toKdl(const urdf::Vector3 &v) { return KDL::Vector(v.x, v.y, v.z); }
inline KDL::Frame toKdl(const urdf::Pose &p) { return KDL::Frame(KDL::Rotation::Quaternion(p.rotation.x, p.rotation.y, p.rotation.z, p.rotation.w), toKdl(p.position)); }
inline KDL::Joint toKdl(const urdf::JointSharedPtr &jnt) {
    const KDL::Frame F_parent_jnt = toKdl(jnt->parent_to_joint_origin_transform);
    switch (jnt->type) {
        case urdf::Joint::FIXED: return KDL::Joint(jnt->name, KDL::Joint::None);
        case urdf::Joint::REVOLUTE:
        case urdf::Joint::CONTINUOUS: return KDL::Joint(jnt->name, F_parent_jnt.p, F_parent_jnt.M * toKdl(jnt->axis), KDL::Joint::RotAxis);
        default: throw std::runtime_error("kdl_parser work-alike: joint type not restated");
    }
}
inline bool addChildrenToTree(const urdf::LinkConstSharedPtr &root, KDL::Tree &tree) {
    const KDL::Segment sgm(root->name, toKdl(root->parent_joint), toKdl(root->parent_joint->parent_to_joint_origin_transform));
    if (!tree.addSegment(sgm, root->parent_joint->parent_link_name)) return false;
    for (const auto &c : root->child_links) if (!addChildrenToTree(c, tree)) return false;
    return true;
}
inline bool treeFromUrdfModel(const urdf::Model &model, KDL::Tree &tree) {
    if (!model.getRoot()) return false;
    tree = KDL::Tree(model.getRoot()->name);
    for (const auto &c : model.getRoot()->child_links) if (!addChildrenToTree(c, tree)) return false;
    return true;
}
inline bool treeFromFile(const std::string &file, KDL::Tree &tree) {
    urdf::Model model;
    return model.initFile(file) && treeFromUrdfModel(model, tree);
}
}  // namespace kdl_parser
