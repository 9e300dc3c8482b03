// rvs_robot_host.h — host-side flattening of an rvs_robot_desc (URDF chain) into the HBM-resident RobotDev record.
#pragma once
#include <cmath>
#include <cstring>

#include "rvs_types.h"

namespace rvs {

inline void rpy_matrix(const double rpy[3], double m[9]) {      // KDL::Rotation::RPY = Rz(yaw) Ry(pitch) Rx(roll)
    const double ca = std::cos(rpy[2]), sa = std::sin(rpy[2]), cb = std::cos(rpy[1]), sb = std::sin(rpy[1]),
                 cc = std::cos(rpy[0]), sc = std::sin(rpy[0]);
    m[0] = ca * cb; m[1] = ca * sb * sc - sa * cc; m[2] = ca * sb * cc + sa * sc;
    m[3] = sa * cb; m[4] = sa * sb * sc + ca * cc; m[5] = sa * sb * cc - ca * sc;
    m[6] = -sb;     m[7] = cb * sc;                m[8] = cb * cc;
}

// Returns nullptr on success, else a static message describing the invalid field.
inline const char *fill_robot_dev(const rvs_robot_desc *rd, RobotDev *out) {
    if (rd->num_dofs < 1 || rd->num_dofs > RVS_MAX_DOFS) return "num_dofs out of range";
    if (rd->kind < RVS_ROBOT_XARM6 || rd->kind > RVS_ROBOT_SPATIAL) return "unknown robot kind";
    if (rd->kind == RVS_ROBOT_XARM6 && rd->num_dofs != 6) return "xarm6 needs 6 DOFs";
    RobotDev &r = *out;
    std::memset(&r, 0, sizeof(r));
    r.kind = rd->kind; r.n = rd->num_dofs; r.ground_included = rd->ground_included; r.gripper_length = rd->gripper_length;
    r.self_collision = 0; r.eps_step = 0.1f;                               // rvs_set_self_collision switches the xArm6 check on
    const int n = r.n;
    for (int k = 0; k <= n; k++) {
        for (int c = 0; c < 3; c++) r.xyz[k][c] = rd->joint_xyz[k][c];
        rpy_matrix(rd->joint_rpy[k], r.rfix[k]);
        r.rfix_identity[k] = (rd->joint_rpy[k][0] == 0.0 && rd->joint_rpy[k][1] == 0.0 && rd->joint_rpy[k][2] == 0.0) ? 1 : 0;
    }
    for (int k = 0; k < n; k++) {
        const double *a = rd->joint_axis[k];
        const double len = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        if (!(len > 0)) return "zero joint axis";
        for (int c = 0; c < 3; c++) r.axis[k][c] = a[c] / len;             // KDL::Joint normalises the axis
        r.axis_code[k] = (r.axis[k][0] == 1.0 && r.axis[k][1] == 0.0 && r.axis[k][2] == 0.0)   ? AXIS_X
                         : (r.axis[k][0] == 0.0 && r.axis[k][1] == 1.0 && r.axis[k][2] == 0.0) ? AXIS_Y
                         : (r.axis[k][0] == 0.0 && r.axis[k][1] == 0.0 && r.axis[k][2] == 1.0) ? AXIS_Z
                                                                                                : AXIS_GENERAL;
        r.radius[k] = rd->capsule_radius[k];
        r.radius_d[k] = double(rd->capsule_radius[k]);
        r.radius_in_d[k] = double(rd->capsule_radius[k]) * std::sqrt(3.0) / 3;      // CollisionAndDistance.cpp:7
        r.radius_in[k] = float(r.radius_in_d[k]);
        r.limit_lo[k] = rd->limit_lo[k];
        r.limit_hi[k] = rd->limit_hi[k];
    }
    for (int k = 0; k < n; k++)                                                   // xArm6.cpp:98-105
        r.radius_new[k] = (rd->kind == RVS_ROBOT_XARM6 && k < n - 1) ? std::fmax(r.radius[k], r.radius[k + 1]) : r.radius[k];
    return nullptr;
}

}  // namespace rvs
