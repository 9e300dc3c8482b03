#!/usr/bin/env python
"""tests/golden/make_motion_golden.py — golden vectors for the dynamic-obstacle motion check, produced by the
REFERENCE's own planning::trajectory::MotionValidity::check(q_previous, q_current) (src/planners/trajectory/MotionValidity.cpp:30-51)
on top of its own Environment::updateEnvironment and RealVectorSpace::isValid, all compiled unmodified against oracle/eigen_shim +
oracle/env_shim (oracle/Makefile target `ref` -> oracle/_ref/librvs_ref.so, entry rvsref_motion_run).  Every case is a run of T
consecutive checks on ONE environment, as DRGBT makes them: each check advances the obstacles by the updates it made before returning.
Run in the build container (needs /root/reference); writes tests/golden/motion_golden.npz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refharness  # noqa: E402
import scenes  # noqa: E402
from rpmplv2_b200 import robots, synthetic  # noqa: E402
from rpmplv2_b200.engine import MOTION_TYPES  # noqa: E402

f32p, u8p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32)
P = lambda a, t: a.ctypes.data_as(t)
ENV = dict(ws_center=scenes.WS_CENTER, ws_radius=1.0, base_radius=0.3, robot_max_vel=1.5)
ROBOTS = {"xarm6": robots.xarm6, "spatial22": lambda: robots.spatial(22), "planar10": lambda: robots.planar(10)}
# (robot, dynamic obstacles, motion law, T motions, K validity points, max_iter_time)
CASES = [("xarm6", 5, "light_directions", 60, 10, 0.05), ("xarm6", 50, "light_directions", 40, 10, 0.05), ("xarm6", 5, "straight", 60, 10, 0.05),
         ("xarm6", 8, "circular", 60, 7, 0.05), ("spatial22", 5, "light_directions", 60, 10, 0.05), ("planar10", 4, "two_tunnels", 80, 10, 0.05),
         ("xarm6", 12, "light_directions", 60, 2, 0.02)]


def case_inputs(robot, M_dyn, law, T, seed=21):
    rb = ROBOTS[robot]()
    reach = 1.0 if robot == "xarm6" else float(np.abs(rb.xyz).sum())
    sc = scenes.dynamic_scene(rb, M_dyn, seed, r_lo=0.45 * reach, r_hi=0.95 * reach, size=0.3 * reach)
    if rb.ground_included == 0:                      # planar chains live in the z = 0 plane, without a ground box
        sc = {k: v[1:].copy() for k, v in sc.items()}
        sc["pos"][:, 2] = 0.0
    q1, q2 = synthetic.edges(rb, T, length=0.3, seed=0x99)
    return rb, sc, q1, q2


def trajectory_points(q1, q2, K):
    """K explicit points per motion, as a spline / Ruckig trajectory would hand them over (MotionValidity.cpp:60): the straight edge
    bent by a sine bump, [T, K, n] float32."""
    s = (np.arange(1, K + 1, dtype=np.float32) / np.float32(K))[None, :, None]
    bump = np.float32(1.5) * np.sin(np.float32(np.pi) * s) * np.roll(q2 - q1, 1, axis=1)[:, None, :]
    return np.ascontiguousarray(q1[:, None, :] + (q2 - q1)[:, None, :] * s + bump, np.float32)


def delta_time(K, max_iter_time):
    """MotionValidity.cpp:34: float delta_time = max_iter_time / NUM_VALIDITY_POINTS_CHECK."""
    return float(np.float32(max_iter_time) / np.float32(K))


def run_reference(lib, rb, sc, law, q1, q2, K, max_iter_time, points=None):
    # MotionValidity::check also asks robot->checkSelfCollision(q) (MotionValidity.cpp:42,71): kept off here, the engine's motion check
    # covers the obstacle part (the capsule self-collision of the K points is a separate engine call)
    h, _ = refharness.create(lib, rb, 1e-4, self_collision=False)
    prm = np.array(list(ENV["ws_center"]) + [ENV["ws_radius"], ENV["base_radius"], ENV["robot_max_vel"]], np.float32)
    T, M = len(q1), len(sc["dim"])
    c = lambda k, dt: np.ascontiguousarray(sc[k], dt)
    r = dict(valid=np.zeros(T, np.uint8), passed=np.zeros(T, np.int32), pos=np.zeros((T, M, 3), np.float32), vel=np.zeros((T, M, 3), np.float32))
    env = (P(prm, f32p), MOTION_TYPES[law], M, P(c("dim", np.float32), f32p), P(c("pos", np.float32), f32p), P(c("vel", np.float32), f32p),
           P(c("max_vel", np.float32), f32p), P(c("is_dynamic", np.uint8), u8p), P(c("is_ground", np.uint8), u8p), P(c("min_dist_tol", np.float32), f32p))
    out = (ctypes.c_float(max_iter_time), P(r["valid"], u8p), P(r["passed"], i32p), P(r["pos"], f32p), P(r["vel"], f32p))
    if points is None:
        assert lib.rvsref_motion_run(h, *env, P(q1, f32p), P(q2, f32p), T, K, *out) == 0
    else:
        assert lib.rvsref_motion_points_run(h, *env, P(points, f32p), T, K, *out) == 0
    lib.rvsref_destroy(h)
    return r


if __name__ == "__main__":
    lib = refharness.load()
    data = {}
    for k, (robot, M_dyn, law, T, K, mit) in enumerate(CASES):
        rb, sc, q1, q2 = case_inputs(robot, M_dyn, law, T)
        r = run_reference(lib, rb, sc, law, q1, q2, K, mit)
        print(robot, M_dyn, law, "valid", int(r["valid"].sum()), "of", T)
        for name, v in r.items():
            data[f"c{k}_{name}"] = v
        r = run_reference(lib, rb, sc, law, q1, q2, K, mit, points=trajectory_points(q1, q2, K))
        print("   explicit points: valid", int(r["valid"].sum()), "of", T)
        for name, v in r.items():
            data[f"c{k}_pts_{name}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "motion_golden.npz"), **data)
