#!/usr/bin/env python
"""tests/golden/make_robot_golden.py — golden vectors for the ROBOT-level rows of the hot path, produced by
the REFERENCE's own robot classes — src/robots/{xArm6,Planar2DOF,Planar10DOF,Spatial10DOF,Spatial14DOF,Spatial18DOF,Spatial22DOF}.cpp,
compiled unmodified against the KDL / kdl_parser / urdfdom / FCL / glog work-alikes of oracle/env_shim (oracle/Makefile target `ref`) and
constructed from the reference's own URDFs the way scenario::Scenario does (Scenario.cpp:13-51):
  frames    computeForwardKinematics        (xArm6.cpp:107-134, Planar2DOF.cpp:82-108)   [double; KDL arithmetic = work-alike]
  skeleton  computeSkeleton                 (xArm6.cpp:209-235, Planar2DOF.cpp:117-130)
  radii     computeEnclosingRadii           (xArm6.cpp:237-275, Planar2DOF.cpp:132-149, Spatial10DOF.cpp:11-30)
  caps_d    computeCapsulesDistance, pairs {0-5, 1-5, 0-4, 1-4}                          (xArm6.cpp:417-424)
  self_hit  checkSelfCollision(q), capsule stage                                          (xArm6.cpp:374-411)
  edge_*    checkSelfCollision(q1, q2&): verdict and q2 after the call                    (xArm6.cpp:284-368)
The FCL mesh confirmation (checkRealSelfCollision) is third-party arithmetic: the FCL stand-in answers "collide", so the xArm6 verdicts
are those of the capsule stage — the conservative contract of the CUDA engine (rvs_self_collision_capsules / rvs_self_collision_edges).
Run in the build container (needs /root/reference); writes tests/golden/robot_golden.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refharness  # noqa: E402
from rpmplv2_b200 import robots, synthetic  # noqa: E402

N = 192
EDGE_SETS = ((0.1, 0.3, 0x51), (0.1, 1.5, 0x52), (0.25, 3.0, 0x53), (0.1, 0.05, 0x54))      # (EPS_STEP, edge length, seed)
E = 400


def robots_under_test():
    return {"xarm6": robots.xarm6(), "xarm6_nogripper": robots.xarm6(gripper_length=0.0, ground_included=0, capsules_radius=(0.06,)),
            "planar_2dof": robots.planar(2), "planar_10dof": robots.planar(10), "spatial_10dof": robots.spatial(10),
            "spatial_14dof": robots.spatial(14), "spatial_18dof": robots.spatial(18), "spatial_22dof": robots.spatial(22)}


def configs(rb):
    q = synthetic.configs(rb, N, seed=0xC0)
    q[0] = 0                                                      # the URDF's zero pose (SURVEY.md §8.2 known answers)
    q[1] = np.clip(np.float32(1.0), rb.limits[:, 0], rb.limits[:, 1])
    return q


def edge_inputs(rb, length, seed):
    q1, q2 = synthetic.edges(rb, E, length, seed=seed)
    q2[:4] = q1[:4]                                               # zero-length motions: the bubble step divides by zero (inf / NaN path)
    far = synthetic.configs(rb, 8, seed=seed + 100)
    q2[4:12] = far                                                # and a few long ones
    return q1, q2


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for name, rb in robots_under_test().items():
        h, used = refharness.create(lib, rb, 1e-4)
        assert used, name
        info = refharness.robot_info(lib, h, rb.num_dofs)
        q = configs(rb)
        data[f"{name}__limits"], data[f"{name}__radii"] = info["limits"], info["radii"]
        data[f"{name}__frames"] = refharness.frames(lib, h, q[:32])
        data[f"{name}__skeleton"] = refharness.skeleton(lib, h, q)
        data[f"{name}__radii_matrix"] = refharness.enclosing_radii(lib, h, q)
        if rb.kind == robots.ROBOT_XARM6:
            q_many = synthetic.configs(rb, 4000, seed=0xC1)
            d, hit = refharness.self_collision(lib, h, q_many)
            data[f"{name}__caps_d"], data[f"{name}__self_hit"] = d, hit
            for k, (eps, length, seed) in enumerate(EDGE_SETS):
                q1, q2 = edge_inputs(rb, length, seed)
                ehit, q2_out = refharness.self_collision_edges(lib, h, q1, q2, eps)
                data[f"{name}__edge{k}_hit"], data[f"{name}__edge{k}_q2"] = ehit, q2_out
                print(name, "edges", eps, length, "hits", int(ehit.sum()), "of", E)
            lib.rvsref_set_self_collision(h, 0)                   # self_collision_checking: false -> constant false, q2 untouched
            ehit, q2_out = refharness.self_collision_edges(lib, h, q1, q2, eps)
            assert not ehit.any() and np.array_equal(q2_out, q2)
            print(name, "self-collisions", int(hit.sum()), "of", len(hit))
        else:
            q1, q2 = edge_inputs(rb, 0.5, 0x55)
            ehit, q2_out = refharness.self_collision_edges(lib, h, q1, q2, 0.1)
            assert not ehit.any() and np.array_equal(q2_out, q2)  # the other classes' stubs (Planar2DOF.cpp:151-161, Spatial10DOF.cpp:32-43)
        lib.rvsref_destroy(h)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "robot_golden.npz"), **data)
    print("wrote", len(robots_under_test()), "robots,", sum(v.nbytes for v in data.values()) // 1024, "KiB raw")
