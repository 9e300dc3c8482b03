"""tests/golden/refharness.py — ctypes front of oracle/_ref/librvs_ref.so for the golden-vector generators (run in the build container only:
needs /root/reference).  `create()` builds the harness around the REFERENCE'S OWN robot class (robots::xArm6, Planar2DOF, Planar10DOF,
Spatial{10,14,18,22}DOF constructed from the reference's URDF like scenario::Scenario does, Scenario.cpp:13-51) whenever the chain is one
the reference ships; other chains (generic test chains) fall back to the descriptor-only harness robot (oracle FK + oracle radii)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_DATA = "/root/reference/data"
f32p, f64p, u8p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32)
P = lambda a, t: a.ctypes.data_as(t)
REF_TYPES = ("xarm6", "planar_2dof", "planar_10dof", "spatial_10dof", "spatial_14dof", "spatial_18dof", "spatial_22dof")


def load(name="librvs_ref.so"):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", name))
    lib.rvsref_create.restype = ctypes.c_void_p
    lib.rvsref_create_robot.restype = ctypes.c_void_p
    lib.rvsref_pattern_tree.restype = ctypes.c_int64
    return lib


def urdf_of(rb):
    """Path of the reference URDF this chain was tabulated from, or None (generic chain / reference data absent)."""
    if rb.name not in REF_TYPES:
        return None
    path = os.path.join(REF_DATA, rb.name, rb.name + ".urdf")
    return path if os.path.exists(path) else None


def create(lib, rb, thr=1e-4, max_dist=1000.0, num_checks=10, self_collision=None, force_descriptor=False):
    """-> (handle, used_reference_class)."""
    desc = rb.to_desc()
    urdf = None if force_descriptor else urdf_of(rb)
    if urdf is not None:
        sc = rb.self_collision_checking if self_collision is None else self_collision
        h = lib.rvsref_create_robot(ctypes.byref(desc), rb.name.encode(), urdf.encode(), int(bool(sc)), ctypes.c_float(thr), ctypes.c_float(max_dist), int(num_checks))
        if not h:
            raise RuntimeError(f"rvsref_create_robot({rb.name}) failed")
        return ctypes.c_void_p(h), True
    h = lib.rvsref_create(ctypes.byref(desc), ctypes.c_float(thr), ctypes.c_float(max_dist), int(num_checks))
    if not h:
        raise RuntimeError("rvsref_create failed")
    return ctypes.c_void_p(h), False


def set_obstacles(lib, h, mn, mx, g, tol):
    mn, mx, g, tol = [np.ascontiguousarray(a) for a in (mn, mx, g, tol)]
    lib.rvsref_set_obstacles(h, P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), len(g))


def frames(lib, h, q):
    q = np.ascontiguousarray(q, np.float32)
    N, n = q.shape
    out = np.zeros((N, n, 12), np.float64)
    assert lib.rvsref_frames(h, P(q, f32p), ctypes.c_int64(N), P(out, f64p)) == 0
    return out


def skeleton(lib, h, q):
    q = np.ascontiguousarray(q, np.float32)
    N, n = q.shape
    out = np.zeros((N, n + 1, 3), np.float32)
    assert lib.rvsref_skeleton(h, P(q, f32p), ctypes.c_int64(N), P(out, f32p)) == 0
    return out


def enclosing_radii(lib, h, q):
    q = np.ascontiguousarray(q, np.float32)
    N, n = q.shape
    out = np.zeros((N, n, n + 1), np.float32)
    assert lib.rvsref_enclosing_radii(h, P(q, f32p), ctypes.c_int64(N), P(out, f32p)) == 0
    return out


def self_collision(lib, h, q):
    q = np.ascontiguousarray(q, np.float32)
    N = len(q)
    d, hit = np.zeros((N, 4), np.float32), np.zeros(N, np.uint8)
    assert lib.rvsref_self_collision(h, P(q, f32p), ctypes.c_int64(N), P(d, f32p), P(hit, u8p)) == 0
    return d, hit


def self_collision_edges(lib, h, q1, q2, eps_step=0.1):
    q1, q2 = np.ascontiguousarray(q1, np.float32), np.ascontiguousarray(q2, np.float32)
    E = len(q1)
    hit, q2_out = np.zeros(E, np.uint8), np.zeros_like(q2)
    assert lib.rvsref_self_collision_edges(h, P(q1, f32p), P(q2, f32p), ctypes.c_int64(E), ctypes.c_float(eps_step), P(hit, u8p), P(q2_out, f32p)) == 0
    return hit, q2_out


def robot_info(lib, h, n):
    lim, rad = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    g, sc, nl, gl = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_float()
    nd = lib.rvsref_robot_info(h, P(lim, f32p), P(rad, f32p), ctypes.byref(g), ctypes.byref(gl), ctypes.byref(sc), ctypes.byref(nl))
    return dict(num_dofs=nd, limits=lim, radii=rad, ground_included=g.value, gripper_length=gl.value, self_collision=sc.value, num_links=nl.value)
