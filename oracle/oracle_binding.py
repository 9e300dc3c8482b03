"""ctypes binding of the CPU oracle (oracle/_build/librvs_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs. No product module may import this file.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

from rpmplv2_b200.robots import DIST_FULL, DIST_MIN, DIST_PROFILE, Robot, RvsParams, RvsRobotDesc, default_params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librvs_oracle.so")
_lib = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (no FMA contraction, like the reference's x86-64 Release build)."""
    srcs = [os.path.join(_HERE, f) for f in ("rvs_oracle_capi.cpp", "rvs_oracle.hpp")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s", "oracle"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_create.restype = ctypes.c_void_p
        _lib.orc_create.argtypes = [ctypes.POINTER(RvsRobotDesc), ctypes.POINTER(RvsParams)]
        _lib.orc_destroy.argtypes = [ctypes.c_void_p]
        _lib.orc_pair_count.restype = ctypes.c_uint64
        _lib.orc_pair_count.argtypes = [ctypes.c_void_p, ctypes.c_int]
    return _lib


def _p(a: Optional[np.ndarray], typ):
    return None if a is None else a.ctypes.data_as(typ)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def lib_norm(a, b) -> float:
    """(q1 - q2).norm() in Eigen's summation order (rvs_oracle.hpp coord_norm) through the library."""
    a, b = _f32(a), _f32(b)
    fn = lib().orc_coord_norm
    fn.restype = ctypes.c_float
    return float(fn(_p(a, c_f32p), _p(b, c_f32p), ctypes.c_int32(len(a))))


class Oracle:
    """CPU oracle with the same call surface as rpmplv2_b200.engine.Engine (host arrays)."""

    def __init__(self, robot: Robot, params: Optional[RvsParams] = None):
        self.robot = robot
        self.params = params or default_params()
        self.n = robot.num_dofs
        self.M = 0
        self._desc = robot.to_desc()
        self._h = ctypes.c_void_p(lib().orc_create(ctypes.byref(self._desc), ctypes.byref(self.params)))
        if not self._h:
            raise RuntimeError("orc_create failed")
        # robots::xArm6's constructor switches self-collision checking on (xArm6.cpp:67); every other class has none
        if robot.self_collision_checking:
            self.set_self_collision(True, 0.1)

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    def pair_count(self, reset: bool = False) -> int:
        return int(lib().orc_pair_count(self._h, 1 if reset else 0))

    def set_obstacles(self, aabb_min, aabb_max, is_ground=None, min_dist_tol=None):
        mn, mx = _f32(aabb_min, (-1, 3)), _f32(aabb_max, (-1, 3))
        self.M = mn.shape[0]
        g = None if is_ground is None else np.ascontiguousarray(is_ground, np.uint8)
        t = None if min_dist_tol is None else _f32(min_dist_tol)
        lib().orc_set_obstacles(self._h, _p(mn, c_f32p), _p(mx, c_f32p), _p(g, c_u8p), _p(t, c_f32p), ctypes.c_int32(self.M))

    # ---- dynamic environment (row f-2): the state arrays live on this object, like env::Environment's objects -------------
    def env_set(self, ep, dim, pos, vel=None, max_vel=None, is_dynamic=None, is_ground=None, min_dist_tol=None, rotated=None):
        dim, pos = _f32(dim, (-1, 3)), _f32(pos, (-1, 3)).copy()
        M = dim.shape[0]
        half = (np.float32(0.5) * dim).astype(np.float32)
        if rotated is not None:
            # FCL's bounding-sphere cube for rotated boxes, summed as the engine's host code does: (x*x + y*y) + z*z
            rad = np.sqrt(((half[:, 0] * half[:, 0] + half[:, 1] * half[:, 1]) + half[:, 2] * half[:, 2]).astype(np.float32)).astype(np.float32)
            half = np.where(np.asarray(rotated, bool)[:, None], rad[:, None], half).astype(np.float32)
        self._env = dict(ep=ep, M=M, half=np.ascontiguousarray(half), pos=pos,
                         vel=np.zeros((M, 3), np.float32) if vel is None else _f32(vel, (M, 3)).copy(),
                         max_vel=np.zeros(M, np.float32) if max_vel is None else _f32(max_vel, (M,)).copy(),
                         sign=np.ones(M, np.int32), path_len=np.zeros(M, np.float32),
                         dynamic=np.zeros(M, np.uint8) if is_dynamic is None else np.ascontiguousarray(is_dynamic, np.uint8).copy())
        self._env["ground"], self._env["tol"] = is_ground, min_dist_tol
        self.set_obstacles(pos + (-half), pos + half, is_ground, min_dist_tol)

    def _env_args(self):
        e = self._env
        ep = e["ep"]
        prm = np.array([ep.ws_center[0], ep.ws_center[1], ep.ws_center[2], ep.ws_radius, ep.base_radius, ep.robot_max_vel], np.float32)
        return prm, [ctypes.c_int32(ep.motion_type), ctypes.c_int32(e["M"]), _p(e["half"], c_f32p), _p(e["pos"], c_f32p), _p(e["vel"], c_f32p),
                     _p(e["max_vel"], c_f32p), _p(e["sign"], c_i32p), _p(e["path_len"], c_f32p), _p(e["dynamic"], c_u8p)]

    def env_set_random_directions(self, acc=None, ground_included: int = -1, seed: int = -1):
        """Extra state of Environment::moveRandomDirections (Environment.cpp:289-333): accelerations, Environment::ground_included, srand(seed)."""
        e = self._env
        if "acc" not in e:
            e["acc"], e["ground_included"], e["rand"] = np.zeros((e["M"], 3), np.float32), int(self.robot.ground_included), np.zeros(36, np.uint32)
            lib().orc_libc_srand(ctypes.c_uint32(1), _p(e["rand"], ctypes.POINTER(ctypes.c_uint32)))
        if acc is not None:
            e["acc"] = _f32(acc, (e["M"], 3)).copy()
        if ground_included >= 0:
            e["ground_included"] = int(ground_included)
        if seed >= 0:
            lib().orc_libc_srand(ctypes.c_uint32(seed), _p(e["rand"], ctypes.POINTER(ctypes.c_uint32)))

    def env_get_random_directions(self):
        return self._env["acc"].copy(), int(self._env["rand"][35])

    def env_step(self, delta_time: float, steps: int = 1):
        prm, args = self._env_args()
        M = self._env["M"]
        mn, mx = np.empty((M, 3), np.float32), np.empty((M, 3), np.float32)
        if self._env["ep"].motion_type == 3:
            if "acc" not in self._env:
                self.env_set_random_directions()
            e = self._env
            rc = lib().orc_env_update_random(_p(prm, c_f32p), ctypes.c_int32(e["ground_included"]), ctypes.c_int32(M), _p(e["half"], c_f32p), _p(e["pos"], c_f32p),
                                             _p(e["vel"], c_f32p), _p(e["acc"], c_f32p), _p(e["max_vel"], c_f32p), _p(e["dynamic"], c_u8p), ctypes.c_float(delta_time),
                                             ctypes.c_int32(steps), _p(e["rand"], ctypes.POINTER(ctypes.c_uint32)), _p(mn, c_f32p), _p(mx, c_f32p))
            if rc:
                raise RuntimeError(f"orc_env_update_random rc={rc}")
            self.set_obstacles(mn, mx, e["ground"], e["tol"])
            return
        rc = lib().orc_env_update(_p(prm, c_f32p), *args, ctypes.c_float(delta_time), ctypes.c_int32(steps), _p(mn, c_f32p), _p(mx, c_f32p))
        if rc:
            raise RuntimeError(f"orc_env_update rc={rc}")
        self.set_obstacles(mn, mx, self._env["ground"], self._env["tol"])

    def env_get(self):
        return self._env["pos"].copy(), self._env["vel"].copy()

    def _motion_validity_random(self, q_prev, q_cur, K, delta_time, commit_edge):
        """MotionValidity::check under random_directions as a composition of this oracle's own pieces: K x (one update; interpolate; isValid)."""
        e = self._env
        keep = {k: e[k].copy() for k in ("pos", "vel", "acc", "rand")}
        E = q_prev.shape[0]
        first = np.zeros(E, np.int32)
        states = []
        dist = np.array([np.float32(lib_norm(q_prev[i], q_cur[i])) for i in range(E)], np.float32)
        for k in range(1, K + 1):
            self.env_step(delta_time, 1)
            states.append({kk: e[kk].copy() for kk in ("pos", "vel", "acc", "rand")})
            step = (dist * np.float32(k) / np.float32(K)).astype(np.float32)
            with np.errstate(invalid="ignore", divide="ignore"):
                f = (step / dist).astype(np.float32)
            pts = np.where((dist > 0)[:, None], q_prev + (q_cur - q_prev) * f[:, None], q_cur).astype(np.float32)
            ok = self.is_valid(pts)
            first[(first == 0) & ~ok.astype(bool)] = k
        kc = None if commit_edge < 0 else (int(first[commit_edge]) or K)
        st = keep if kc is None else states[kc - 1]
        for kk in ("pos", "vel", "acc", "rand"):
            e[kk] = st[kk].copy()
        half = e["half"]
        self.set_obstacles(e["pos"] + (-half), e["pos"] + half, e["ground"], e["tol"])
        return first == 0, first

    def motion_validity(self, q_prev, q_cur, K: int, delta_time: float, commit_edge: int = -1):
        q_prev, q_cur = _f32(q_prev, (-1, self.n)), _f32(q_cur, (-1, self.n))
        E = q_prev.shape[0]
        if self._env["ep"].motion_type == 3:
            return self._motion_validity_random(q_prev, q_cur, K, delta_time, commit_edge)
        valid, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        prm, args = self._env_args()
        rc = lib().orc_motion_validity(self._h, _p(prm, c_f32p), *args, _p(q_prev, c_f32p), _p(q_cur, c_f32p), ctypes.c_int64(E), ctypes.c_int32(K),
                                       ctypes.c_float(delta_time), ctypes.c_int32(commit_edge), _p(valid, c_u8p), _p(first, c_i32p))
        if rc:
            raise RuntimeError(f"orc_motion_validity rc={rc}")
        return valid.astype(bool), first

    def motion_validity_points(self, points, delta_time: float, commit_edge: int = -1):
        points = np.ascontiguousarray(points, np.float32)
        if points.ndim == 2:
            points = points[None]
        E, K, _ = points.shape
        valid, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        prm, args = self._env_args()
        rc = lib().orc_motion_validity_points(self._h, _p(prm, c_f32p), *args, _p(points, c_f32p), ctypes.c_int64(E), ctypes.c_int32(K),
                                              ctypes.c_float(delta_time), ctypes.c_int32(commit_edge), _p(valid, c_u8p), _p(first, c_i32p))
        if rc:
            raise RuntimeError(f"orc_motion_validity_points rc={rc}")
        return valid.astype(bool), first

    def skeleton(self, q, f64=False):
        q = _f32(q, (-1, self.n))
        N = q.shape[0]
        if f64:
            sk = np.empty((N, self.n + 1, 3), np.float64)
            lib().orc_skeleton_f64(self._h, _p(q, c_f32p), ctypes.c_int64(N), _p(sk, c_f64p))
        else:
            sk = np.empty((N, self.n + 1, 3), np.float32)
            lib().orc_skeleton_f32(self._h, _p(q, c_f32p), ctypes.c_int64(N), _p(sk, c_f32p))
        return sk

    def is_valid(self, q, f64=False, threads=1):
        q = _f32(q, (-1, self.n))
        out = np.empty(q.shape[0], np.uint8)
        fn = lib().orc_is_valid_f64 if f64 else lib().orc_is_valid_f32
        fn(self._h, _p(q, c_f32p), ctypes.c_int64(q.shape[0]), _p(out, c_u8p), ctypes.c_int(threads))
        return out

    def is_valid_edges(self, q1, q2, threads=1):
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        out = np.empty(q1.shape[0], np.uint8)
        lib().orc_is_valid_edges_f32(self._h, _p(q1, c_f32p), _p(q2, c_f32p), ctypes.c_int64(q1.shape[0]), _p(out, c_u8p),
                                     ctypes.c_int(threads))
        return out

    def distance(self, q, mode=DIST_PROFILE, f64=False, threads=1):
        q = _f32(q, (-1, self.n))
        N = q.shape[0]
        dt, pt, fn = (np.float64, c_f64p, lib().orc_distance_f64) if f64 else (np.float32, c_f32p, lib().orc_distance_f32)
        d_c = np.empty(N, dt)
        profile = np.zeros((N, self.n), dt) if mode >= DIST_PROFILE else None
        nearest = np.zeros((N, self.M, self.n, 6), dt) if mode == DIST_FULL else None
        fn(self._h, _p(q, c_f32p), ctypes.c_int64(N), ctypes.c_int32(mode), _p(d_c, pt), _p(profile, pt), _p(nearest, pt),
           ctypes.c_int(threads))
        return d_c, profile, nearest

    def underestimation(self, q, planes, plane_idx=None):
        q = _f32(q, (-1, self.n))
        N = q.shape[0]
        planes = _f32(planes, (-1, self.M, self.n, 6))
        idx = None if plane_idx is None else np.ascontiguousarray(plane_idx, np.int32)
        d_c, profile, status = np.empty(N, np.float32), np.empty((N, self.n), np.float32), np.empty(N, np.uint8)
        rc = lib().orc_underestimation_f32(self._h, _p(q, c_f32p), ctypes.c_int64(N), _p(planes, c_f32p), _p(idx, c_i32p),
                                           ctypes.c_int32(planes.shape[0]), _p(d_c, c_f32p), _p(profile, c_f32p), _p(status, c_u8p))
        if rc:
            raise ValueError("plane_idx out of range")
        return d_c, profile, status

    def enclosing_radii(self, q):
        q = _f32(q, (-1, self.n))
        out = np.empty((q.shape[0], self.n, self.n + 1), np.float32)
        lib().orc_enclosing_radii_f32(self._h, _p(q, c_f32p), ctypes.c_int64(q.shape[0]), _p(out, c_f32p))
        return out

    def extend_spines(self, q_root, root_profile, root_d_c, root_planes, q_targets, spine_params):
        """RGBTConnect::extendGenSpine (distance-to-planes variant) from one root towards every target."""
        q_root, root_profile = _f32(q_root, (self.n,)), _f32(root_profile, (self.n,))
        q_targets, root_planes = _f32(q_targets, (-1, self.n)), _f32(root_planes)
        S = q_targets.shape[0]
        q_out, d_out, status = np.empty((S, self.n), np.float32), np.empty(S, np.float32), np.empty(S, np.uint8)
        lib().orc_extend_spines_f32(self._h, _p(q_root, c_f32p), _p(root_profile, c_f32p), ctypes.c_float(root_d_c), _p(root_planes, c_f32p),
                                    _p(q_targets, c_f32p), ctypes.c_int32(S), ctypes.byref(spine_params), _p(q_out, c_f32p), _p(d_out, c_f32p),
                                    _p(status, ctypes.POINTER(ctypes.c_uint8)))
        return q_out, d_out, status

    def extend_spines_layers(self, q_roots, root_profiles, root_d_c, root_planes, root_idx, q_targets, spine_params, gen2=False):
        """extendGenSpine / extendGenSpine2 with per-spine roots, every layer kept. Returns (q_layers[S,layers,n], n_layers, d_c_last, status)."""
        q_roots, root_profiles = _f32(q_roots, (-1, self.n)), _f32(root_profiles, (-1, self.n))
        root_d_c, root_planes = _f32(root_d_c, (-1,)), _f32(root_planes)
        root_idx = np.ascontiguousarray(root_idx, np.int32)
        q_targets = _f32(q_targets, (-1, self.n))
        S = q_targets.shape[0]
        q_layers = np.zeros((S, spine_params.num_layers, self.n), np.float32)
        n_layers, d_last, status = np.empty(S, np.int32), np.empty(S, np.float32), np.empty(S, np.uint8)
        lib().orc_extend_spines_layers_f32(self._h, _p(q_roots, c_f32p), _p(root_profiles, c_f32p), _p(root_d_c, c_f32p), _p(root_planes, c_f32p),
                                           _p(root_idx, c_i32p), _p(q_targets, c_f32p), ctypes.c_int32(S), ctypes.byref(spine_params),
                                           ctypes.c_int32(1 if gen2 else 0), _p(q_layers, c_f32p), _p(n_layers, c_i32p), _p(d_last, c_f32p),
                                           _p(status, ctypes.POINTER(ctypes.c_uint8)))
        return q_layers, n_layers, d_last, status

    def self_collision_capsules(self, q):
        q = _f32(q, (-1, self.n))
        out = np.empty(q.shape[0], np.uint8)
        lib().orc_self_collision_capsules_f32(self._h, _p(q, c_f32p), ctypes.c_int64(q.shape[0]), _p(out, c_u8p))
        return out

    def set_self_collision(self, enabled: bool, eps_step: float = 0.1):
        """AbstractRobot::setSelfCollisionChecking + RRTConnectConfig::EPS_STEP (xArm6.cpp:340-341)."""
        lib().orc_set_self_collision(self._h, ctypes.c_int32(1 if enabled else 0), ctypes.c_float(eps_step))

    def self_collision_edges(self, q1, q2):
        """xArm6::checkSelfCollision(q1, q2&) capsule stage (xArm6.cpp:284-368) -> (hit[E], q2 after the call)."""
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        hit, out = np.empty(q1.shape[0], np.uint8), np.empty_like(q2)
        lib().orc_self_collision_edges_f32(self._h, _p(q1, c_f32p), _p(q2, c_f32p), ctypes.c_int64(q1.shape[0]), _p(hit, c_u8p), _p(out, c_f32p))
        return hit, out

    def generate_configs(self, seed: int, first: int, N: int):
        q = np.empty((N, self.n), np.float32)
        lib().orc_generate_configs(self._h, ctypes.c_uint64(seed), ctypes.c_int64(first), ctypes.c_int64(N), _p(q, c_f32p))
        return q


# --- primitives -------------------------------------------------------------------------------------------------

def seg_seg(ABCD, thr=1e-4, maxd=1000.0, f64=False):
    dt, pt = (np.float64, c_f64p) if f64 else (np.float32, c_f32p)
    a = np.ascontiguousarray(ABCD, dt).reshape(-1, 12)
    out = np.empty((a.shape[0], 8), dt)
    ct = ctypes.c_double if f64 else ctypes.c_float
    (lib().orc_seg_seg_f64 if f64 else lib().orc_seg_seg_f32)(ctypes.c_int64(a.shape[0]), _p(a, pt), ct(thr), ct(maxd), _p(out, pt))
    return out


def seg_point(ABC, thr=1e-4, maxd=1000.0):
    a = _f32(ABC, (-1, 9))
    out = np.empty((a.shape[0], 8), np.float32)
    lib().orc_seg_point_f32(ctypes.c_int64(a.shape[0]), _p(a, c_f32p), ctypes.c_float(thr), ctypes.c_float(maxd), _p(out, c_f32p))
    return out


def capsule_box_distance(AB, radius, obs, thr=1e-4, maxd=1000.0, f64=False):
    dt, pt = (np.float64, c_f64p) if f64 else (np.float32, c_f32p)
    ct = ctypes.c_double if f64 else ctypes.c_float
    a = np.ascontiguousarray(AB, dt).reshape(-1, 6)
    o = np.ascontiguousarray(obs, dt).reshape(-1, 6)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dt), (a.shape[0],)))
    out = np.empty((a.shape[0], 8), dt)
    fn = lib().orc_capsule_box_distance_f64 if f64 else lib().orc_capsule_box_distance_f32
    fn(ctypes.c_int64(a.shape[0]), _p(a, pt), _p(r, pt), _p(o, pt), ct(thr), ct(maxd), _p(out, pt))
    return out


def capsule_box_collision(AB, radius, obs, thr=1e-4, maxd=1000.0):
    a, o = _f32(AB, (-1, 6)), _f32(obs, (-1, 6))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (a.shape[0],)))
    out = np.empty(a.shape[0], np.uint8)
    lib().orc_capsule_box_collision_f32(ctypes.c_int64(a.shape[0]), _p(a, c_f32p), _p(r, c_f32p), _p(o, c_f32p),
                                        ctypes.c_float(thr), ctypes.c_float(maxd), _p(out, c_u8p))
    return out


def capsule_rect_collision(AB, radius, obs, face, thr=1e-4, maxd=1000.0):
    a, o = _f32(AB, (-1, 6)), _f32(obs, (-1, 6))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (a.shape[0],)))
    f = np.ascontiguousarray(np.broadcast_to(np.asarray(face, np.int32), (a.shape[0],)))
    out = np.empty(a.shape[0], np.uint8)
    lib().orc_capsule_rect_collision_f32(ctypes.c_int64(a.shape[0]), _p(a, c_f32p), _p(r, c_f32p), _p(o, c_f32p), _p(f, c_i32p),
                                         ctypes.c_float(thr), ctypes.c_float(maxd), _p(out, c_u8p))
    return out


# --- planners driven by the oracle (checker of tests/test_planners.py, CPU baseline of bench.py) ---------------------------

_plan_lib = None


def plan_lib():
    global _plan_lib
    if _plan_lib is None:
        from rpmplv2_b200.planning import PlanConfig, PlanResult
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liborc_planning.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found: run `make -C oracle oracle`")
        _plan_lib = ctypes.CDLL(path)
        _plan_lib.orcplan_solve_batch.restype = ctypes.c_int
        _plan_lib.orcplan_solve_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(PlanConfig), c_f32p, c_f32p, ctypes.POINTER(ctypes.c_uint64),
                                                  ctypes.c_int32, ctypes.POINTER(PlanResult), c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.POINTER(ctypes.c_int64)]
        _plan_lib.orcplan_last_error.restype = ctypes.c_char_p
        _plan_lib.orcplan_solve_batch_trees.restype = ctypes.c_int
        _plan_lib.orcplan_solve_batch_trees.argtypes = [ctypes.c_void_p, ctypes.POINTER(PlanConfig), c_f32p, c_f32p, ctypes.POINTER(ctypes.c_uint64),
                                                        ctypes.c_int32, ctypes.POINTER(PlanResult), c_f32p, ctypes.c_int32, c_f32p, c_i32p, c_i32p, ctypes.c_int32]
        from rpmplv2_b200.planning import SYMBOLS as _PS
        _plan_lib.orcplan_rrtx_grow.restype, _plan_lib.orcplan_rrtx_grow.argtypes = _PS["rvsplan_rrtx_grow"]
    return _plan_lib


def plan_solve_batch(orc: "Oracle", cfg, q_start, q_goal, seeds=None, max_path: int = 512, threads: int = 1, lockstep: bool = False):
    """Same planner code as rpmplv2_b200.planning.solve_batch, state-space queries answered by the CPU oracle."""
    from rpmplv2_b200.planning import PlanResult, unpack_results
    n = orc.n
    qs, qg = np.atleast_2d(np.asarray(q_start, np.float32)), np.atleast_2d(np.asarray(q_goal, np.float32))
    P = len(seeds) if seeds is not None else qs.shape[0]
    qs = np.ascontiguousarray(np.broadcast_to(qs, (P, n)))
    qg = np.ascontiguousarray(np.broadcast_to(qg, (P, n)))
    sd = np.ascontiguousarray(np.arange(P) if seeds is None else seeds, np.uint64)
    results = (PlanResult * P)()
    paths = np.zeros((P, max_path, n), np.float32)
    stats = (ctypes.c_int64 * 3)()
    lib_ = plan_lib()
    # the engine handle of the oracle library is a plain heap object: it can be shared with the planning library, which is built
    # from the same header (rvs_oracle_engine.hpp)
    rc = lib_.orcplan_solve_batch(orc._h, ctypes.byref(cfg), _p(qs, c_f32p), _p(qg, c_f32p), sd.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), P,
                                  results, _p(paths, c_f32p), max_path, threads, 1 if lockstep else 0, stats)
    if rc != 0:
        raise RuntimeError((lib_.orcplan_last_error() or b"").decode())
    return unpack_results(results, paths, n), {"rounds": int(stats[0]), "backend_calls": int(stats[1])}


def plan_solve_batch_trees(orc: "Oracle", cfg, q_start, q_goal, seeds, max_path: int = 512, trees: int = 4096):
    """plan_solve_batch (one pool) that also returns both trees of every run: the CPU twin of planning.solve_batch(..., trees=T)."""
    from rpmplv2_b200.planning import PlanResult, attach_trees, unpack_results
    n, P = orc.n, len(seeds)
    qs = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(q_start, np.float32)), (P, n)))
    qg = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(q_goal, np.float32)), (P, n)))
    sd = np.ascontiguousarray(seeds, np.uint64)
    results, paths = (PlanResult * P)(), np.zeros((P, max_path, n), np.float32)
    tc, tp, ts = np.zeros((2 * P, trees, n), np.float32), np.zeros((2 * P, trees), np.int32), np.zeros(2 * P, np.int32)
    rc = plan_lib().orcplan_solve_batch_trees(orc._h, ctypes.byref(cfg), _p(qs, c_f32p), _p(qg, c_f32p), sd.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), P,
                                              results, _p(paths, c_f32p), max_path, _p(tc, c_f32p), _p(tp, c_i32p), _p(ts, c_i32p), trees)
    if rc != 0:
        raise RuntimeError((plan_lib().orcplan_last_error() or b"").decode())
    out = unpack_results(results, paths, n)
    attach_trees(out, tc, tp, ts)
    return out


def plan_rrtx_grow(orc: "Oracle", cfg, q_root, samples, path_start: int = -1, max_states=None) -> dict:
    """Same RRTx code as rpmplv2_b200.planning.rrtx_grow (cpp/planning/Rrtx.cpp), edges answered by the CPU oracle, host neighbour scans."""
    from rpmplv2_b200.planning import rrtx_call
    lib_ = plan_lib()
    rc, out = rrtx_call(lib_.orcplan_rrtx_grow, orc._h, cfg, orc.n, q_root, samples, path_start, max_states)
    if rc != 0:
        raise RuntimeError((lib_.orcplan_last_error() or b"").decode())
    return out
