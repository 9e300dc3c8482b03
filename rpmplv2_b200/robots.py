"""Robot chain descriptors for the RealVectorSpaceCUDA engine.

The engine receives a robot as a flat serial-revolute chain (``rvs_robot_desc`` in ``include/rvs_cuda.h``).
This module builds that descriptor either from the built-in tables below (the reference's URDF numbers,
SURVEY.md §8.2; always available, also on the GPU box) or by parsing a URDF the way the reference's robot
constructors do (``kdl_parser::treeFromFile`` + ``robot_tree.getChain``; /root/reference/src/robots/xArm6.cpp:7-29,
/root/reference/src/robots/Planar2DOF.cpp:7-56).
"""
from __future__ import annotations

import ctypes
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

RVS_MAX_DOFS = 32

ROBOT_XARM6, ROBOT_PLANAR, ROBOT_SPATIAL = 0, 1, 2
PREC_REF32, PREC_FAST32, PREC_F64 = 0, 1, 2
DIST_MIN, DIST_PROFILE, DIST_FULL = 0, 1, 2


class RvsRobotDesc(ctypes.Structure):
    """ctypes mirror of ``rvs_robot_desc``."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("num_dofs", ctypes.c_int32),
        ("ground_included", ctypes.c_int32),
        ("gripper_length", ctypes.c_float),
        ("joint_xyz", (ctypes.c_double * 3) * (RVS_MAX_DOFS + 1)),
        ("joint_rpy", (ctypes.c_double * 3) * (RVS_MAX_DOFS + 1)),
        ("joint_axis", (ctypes.c_double * 3) * RVS_MAX_DOFS),
        ("capsule_radius", ctypes.c_float * RVS_MAX_DOFS),
        ("limit_lo", ctypes.c_float * RVS_MAX_DOFS),
        ("limit_hi", ctypes.c_float * RVS_MAX_DOFS),
    ]


class RvsParams(ctypes.Structure):
    """ctypes mirror of ``rvs_params`` (RealVectorSpaceConfig made explicit)."""

    _fields_ = [
        ("equality_threshold", ctypes.c_float),
        ("max_distance", ctypes.c_float),
        ("num_interpolation_validity_checks", ctypes.c_int32),
        ("precision", ctypes.c_int32),
    ]


class RvsSpineParams(ctypes.Structure):
    _fields_ = [
        ("num_layers", ctypes.c_int32),
        ("num_iter_spine", ctypes.c_int32),
        ("use_expanded_bubble", ctypes.c_int32),
        ("d_crit", ctypes.c_float),
    ]


def default_params(precision: int = PREC_REF32, *, equality_threshold: float = 1e-4,
                   max_distance: float = 1000.0, num_checks: int = 10) -> RvsParams:
    """Shipped YAML values (/root/reference/data/configurations/configuration_realvectorspace.yaml:1-3)."""
    return RvsParams(equality_threshold, max_distance, num_checks, precision)


@dataclass
class Robot:
    """A serial chain: ``n`` revolute joints followed by one fixed tip transform."""

    name: str
    kind: int
    xyz: np.ndarray            # (n+1, 3) float64
    rpy: np.ndarray            # (n+1, 3) float64
    axis: np.ndarray           # (n, 3) float64
    radius: np.ndarray         # (n,) float32   AbstractRobot::capsules_radius
    limits: np.ndarray         # (n, 2) float32
    ground_included: int = 0
    gripper_length: float = 0.0
    self_collision_checking: bool = False
    extra: dict = field(default_factory=dict)

    @property
    def num_dofs(self) -> int:
        return int(self.axis.shape[0])

    @property
    def num_links(self) -> int:
        return self.num_dofs

    def to_desc(self) -> RvsRobotDesc:
        n = self.num_dofs
        if not 1 <= n <= RVS_MAX_DOFS:
            raise ValueError(f"num_dofs {n} outside 1..{RVS_MAX_DOFS}")
        d = RvsRobotDesc()
        d.kind, d.num_dofs = int(self.kind), n
        d.ground_included = int(self.ground_included)
        d.gripper_length = float(self.gripper_length)
        for k in range(n + 1):
            for c in range(3):
                d.joint_xyz[k][c] = float(self.xyz[k, c])
                d.joint_rpy[k][c] = float(self.rpy[k, c])
        for k in range(n):
            for c in range(3):
                d.joint_axis[k][c] = float(self.axis[k, c])
            d.capsule_radius[k] = float(self.radius[k])
            d.limit_lo[k] = float(self.limits[k, 0])
            d.limit_hi[k] = float(self.limits[k, 1])
        return d


def _broadcast_radius(radius, n: int, default: float) -> np.ndarray:
    """Scenario.cpp:36-51: one value is broadcast, otherwise exactly n values."""
    if radius is None:
        return np.full(n, default, np.float32)
    r = np.atleast_1d(np.asarray(radius, np.float32))
    if r.size == 1:
        return np.full(n, r[0], np.float32)
    if r.size != n:
        raise ValueError("Number of capsules is not correct!")
    return r.copy()


# --- built-in tables (numbers from /root/reference/data/**.urdf) ---------------------------

_XARM6_XYZ = [(0, 0, 0.267), (0, 0, 0), (0.0535, -0.2845, 0), (0.0775, 0.3425, 0), (0, 0, 0), (0.076, 0.097, 0), (0, 0, 0)]
_XARM6_RPY = [(0, 0, 0), (-1.5708, 0, 0), (0, 0, 0), (-1.5708, 0, 0), (1.5708, 0, 0), (-1.5708, 0, 0), (0, 0, 0)]
_XARM6_LIMITS = [(-3.14159265359, 3.14159265359), (-2.059, 2.0944), (-3.927, 0.19198),
                 (-3.14159265359, 3.14159265359), (-1.69297, 3.14159265359), (-3.14159265359, 3.14159265359)]
XARM6_DEFAULT_RADII = (0.047, 0.12, 0.11, 0.09, 0.05, 0.05)


def xarm6(capsules_radius: Optional[Sequence[float]] = XARM6_DEFAULT_RADII, gripper_length: float = 0.17,
          ground_included: int = 2) -> Robot:
    """robots::xArm6 (/root/reference/data/xarm6/xarm6.urdf:69-222, chain link_base -> link_eef)."""
    n = 6
    return Robot("xarm6", ROBOT_XARM6, np.array(_XARM6_XYZ, np.float64), np.array(_XARM6_RPY, np.float64),
                 np.tile(np.array([[0.0, 0.0, 1.0]]), (n, 1)), _broadcast_radius(capsules_radius, n, 0.05),
                 np.array(_XARM6_LIMITS, np.float32), ground_included, gripper_length, True)


def planar(n: int, link_length: Optional[float] = None, capsules_radius=None) -> Robot:
    """robots::Planar2DOF / Planar10DOF: z-axis joints along x (/root/reference/data/planar_2dof/planar_2dof.urdf:50-72,
    /root/reference/data/planar_10dof/planar_10dof.urdf:146-231). Radius default = box.y / 2 = 0.025 (Planar2DOF.cpp:51)."""
    if link_length is None:
        link_length = 1.0 if n == 2 else 0.2
    xyz = np.zeros((n + 1, 3))
    xyz[1:, 0] = link_length
    limits = np.tile(np.array([[-3.14159, 3.14159]], np.float32), (n, 1))
    return Robot(f"planar_{n}dof", ROBOT_PLANAR, xyz, np.zeros((n + 1, 3)), np.tile(np.array([[0.0, 0.0, 1.0]]), (n, 1)),
                 _broadcast_radius(capsules_radius, n, 0.025), limits, 0, 0.0, False)


_SPATIAL_H = {10: 0.1, 14: 0.0692, 18: 0.05294, 22: 0.042857}


def spatial(n: int, capsules_radius=(0.025,), ground_included: int = 2, h: Optional[float] = None) -> Robot:
    """robots::Spatial{10,14,18,22}DOF: joint 1 about z at the origin, joints 2..n about x, offsets (0,0,h)
    (/root/reference/data/spatial_10dof/spatial_10dof.urdf:243-320 and siblings)."""
    if h is None:
        h = _SPATIAL_H[n]
    xyz = np.zeros((n + 1, 3))
    xyz[1:, 2] = h
    xyz[1, 2] = 0.1                       # joint2 sits 0.1 above joint1 in all four URDFs
    axis = np.tile(np.array([[1.0, 0.0, 0.0]]), (n, 1))
    axis[0] = (0.0, 0.0, 1.0)
    limits = np.tile(np.array([[-3.14159, 3.14159]], np.float32), (n, 1))
    return Robot(f"spatial_{n}dof", ROBOT_SPATIAL, xyz, np.zeros((n + 1, 3)), axis,
                 _broadcast_radius(capsules_radius, n, 0.025), limits, ground_included, 0.0, False)


def by_type(robot_type: str, **kw) -> Robot:
    """The `robot.type` switch of scenario::Scenario (/root/reference/src/scenario/Scenario.cpp:13-34)."""
    if robot_type == "xarm6":
        return xarm6(**kw)
    if robot_type.startswith("planar_") and robot_type.endswith("dof"):
        return planar(int(robot_type[len("planar_"):-3]), **kw)
    if robot_type.startswith("spatial_") and robot_type.endswith("dof"):
        return spatial(int(robot_type[len("spatial_"):-3]), **kw)
    raise ValueError("Robot type is not correct! Please check whether your robot is added to RPMPLv2 library.")


# --- URDF parsing (reads the reference's own files in place when they are present) -------------------------------

def _vec(text: Optional[str], default=(0.0, 0.0, 0.0)):
    if text is None:
        return tuple(default)
    return tuple(float(t) for t in text.split())


def from_urdf(path: str, robot_type: str, *, capsules_radius=None, gripper_length: float = 0.0,
              ground_included: int = 0) -> Robot:
    """Flatten the URDF chain the reference walks: xarm6 `link_base -> link_eef`, others `base_link -> tool`."""
    root = ET.parse(path).getroot()
    joints = {}
    for j in root.findall("joint"):
        if j.find("parent") is None or j.find("child") is None:
            continue                      # <transmission><joint name=.../> entries
        origin = j.find("origin")
        axis = j.find("axis")
        limit = j.find("limit")
        joints[j.find("child").get("link")] = dict(
            name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"),
            xyz=_vec(origin.get("xyz") if origin is not None else None),
            rpy=_vec(origin.get("rpy") if origin is not None else None),
            axis=_vec(axis.get("xyz") if axis is not None else None, (1.0, 0.0, 0.0)),
            lower=float(limit.get("lower", 0)) if limit is not None else 0.0,
            upper=float(limit.get("upper", 0)) if limit is not None else 0.0)
    base, tip = ("link_base", "link_eef") if robot_type == "xarm6" else ("base_link", "tool")
    chain, link = [], tip
    while link != base:
        chain.append(joints[link])
        link = joints[link]["parent"]
    chain.reverse()
    moving = [j for j in chain if j["type"] in ("revolute", "continuous")]
    if chain[-1]["type"] != "fixed" or len(moving) != len(chain) - 1:
        raise ValueError("expected n revolute joints followed by one fixed tip joint")
    n = len(moving)
    xyz = np.array([j["xyz"] for j in chain], np.float64)
    rpy = np.array([j["rpy"] for j in chain], np.float64)
    axis = np.array([j["axis"] for j in moving], np.float64)
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    # limits are narrowed to float exactly like `float lower = ...->limits->lower` (xArm6.cpp:30-35)
    limits = np.array([(j["lower"], j["upper"]) for j in moving], np.float64).astype(np.float32)
    if robot_type == "xarm6":
        kind, default_r, selfc = ROBOT_XARM6, 0.05, True
    elif robot_type.startswith("planar"):
        kind, default_r, selfc = ROBOT_PLANAR, 0.025, False
    else:
        kind, default_r, selfc = ROBOT_SPATIAL, 0.025, False
    return Robot(robot_type, kind, xyz, rpy, axis, _broadcast_radius(capsules_radius, n, default_r), limits,
                 ground_included, gripper_length, selfc)


def rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """KDL::Rotation::RPY = Rz(yaw) Ry(pitch) Rx(roll) (float64)."""
    ca, sa, cb, sb, cc, sc = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    return np.array([[ca * cb, ca * sb * sc - sa * cc, ca * sb * cc + sa * sc],
                     [sa * cb, sa * sb * sc + ca * cc, sa * sb * cc - ca * sc],
                     [-sb, cb * sc, cb * cc]], np.float64)


def to_urdf(robot: Robot) -> str:
    """URDF text of a chain (inverse of from_urdf): lets the tests exercise the URDF readers without /root/reference."""
    base, tip = ("link_base", "link_eef") if robot.kind == ROBOT_XARM6 else ("base_link", "tool")
    n = robot.num_dofs
    links = [base] + [f"link{k + 1}" for k in range(n)] + [tip]
    out = ['<?xml version="1.0"?>', f'<robot name="{robot.name}">', '  <!-- generated by rpmplv2_b200.robots.to_urdf -->']
    out += [f'  <link name="{name}"/>' for name in links]
    for k in range(n + 1):
        xyz = " ".join(repr(float(v)) for v in robot.xyz[k])
        rpy = " ".join(repr(float(v)) for v in robot.rpy[k])
        if k < n:
            axis = " ".join(repr(float(v)) for v in robot.axis[k])
            lo, hi = (repr(float(np.float64(v))) for v in robot.limits[k])
            out += [f'  <joint name="joint{k + 1}" type="revolute">', f'    <parent link="{links[k]}"/>', f'    <child link="{links[k + 1]}"/>',
                    f'    <origin xyz="{xyz}" rpy="{rpy}"/>', f'    <axis xyz="{axis}"/>',
                    f'    <limit lower="{lo}" upper="{hi}" effort="50" velocity="3.14"/>', '  </joint>']
        else:
            out += [f'  <joint name="joint_tip" type="fixed">', f'    <parent link="{links[k]}"/>', f'    <child link="{links[k + 1]}"/>',
                    f'    <origin xyz="{xyz}" rpy="{rpy}"/>', '  </joint>']
    out += ['  <transmission name="t1"><joint name="joint1"/></transmission>', '</robot>']
    return "\n".join(out) + "\n"
