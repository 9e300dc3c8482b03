// oracle/env_shim/kdl — stand-in for the orocos-KDL types that RPMPLv2's State.h / AbstractRobot.h mention (never used by the
// RealVectorSpace hot path itself).  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace KDL {
struct Vector { double data[3] = {0, 0, 0}; };
struct Rotation { double data[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; };
struct Frame { Rotation M; Vector p; };
}  // namespace KDL
