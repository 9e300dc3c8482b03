// rvs_urdf.h — dependency-free reader that flattens a URDF serial chain into the engine's rvs_robot_desc.
//
// It follows what the reference's robot constructors do with kdl_parser (src/robots/xArm6.cpp:7-36,
// src/robots/Planar2DOF.cpp:7-56 of RPMPLv2): take the chain `link_base -> link_eef` (xarm6) or `base_link -> tool`
// (planar_* / spatial_*), n revolute joints followed by one fixed tip joint, limits narrowed to float.
// The scenario-level fields (capsule radii, gripper_length, ground_included) are filled by the caller.
#pragma once
#include <string>

#include "../../include/rvs_cuda.h"

namespace rvs {

// robot_type is the scenario YAML's `robot.type` ("xarm6", "planar_2dof", "spatial_22dof", ...).
// Returns false and sets *error on malformed input; never throws.
bool chain_from_urdf_text(const std::string &urdf_xml, const std::string &robot_type, rvs_robot_desc *out, std::string *error);
bool chain_from_urdf_file(const std::string &path, const std::string &robot_type, rvs_robot_desc *out, std::string *error);

}  // namespace rvs
