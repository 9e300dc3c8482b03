"""ctypes binding of librvs_cuda.so (include/rvs_cuda.h) — the thin host layer every Python caller goes through.

There is NO CPU fallback here: if the shared library is missing, or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from .robots import (DIST_FULL, DIST_MIN, DIST_PROFILE, PREC_F64, PREC_REF32, Robot, RvsParams, RvsRobotDesc, RvsSpineParams,
                     default_params)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RVS_CUDA_LIB") or os.path.join(_PKG, "librvs_cuda.so")   # RVS_CUDA_LIB: a differently tuned build of the same ABI

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)

MOTION_STRAIGHT, MOTION_CIRCULAR, MOTION_TWO_TUNNELS, MOTION_RANDOM_DIRECTIONS, MOTION_LIGHT_DIRECTIONS = range(5)
MOTION_TYPES = {"straight": MOTION_STRAIGHT, "circular": MOTION_CIRCULAR, "two_tunnels": MOTION_TWO_TUNNELS,
                "random_directions": MOTION_RANDOM_DIRECTIONS, "light_directions": MOTION_LIGHT_DIRECTIONS}


class RvsEnvParams(ctypes.Structure):
    """rvs_env_params (include/rvs_cuda.h)."""
    _fields_ = [("motion_type", ctypes.c_int32), ("ws_center", ctypes.c_float * 3), ("ws_radius", ctypes.c_float),
                ("base_radius", ctypes.c_float), ("robot_max_vel", ctypes.c_float)]


def env_params(motion_type, ws_center=(0.0, 0.0, 0.267), ws_radius=1.0, base_radius=0.3, robot_max_vel=1.0) -> RvsEnvParams:
    mt = MOTION_TYPES[motion_type] if isinstance(motion_type, str) else int(motion_type)
    return RvsEnvParams(mt, (ctypes.c_float * 3)(*[float(x) for x in ws_center]), float(ws_radius), float(base_radius), float(robot_max_vel))

#: every symbol include/rvs_cuda.h declares: name -> (restype, argtypes)
_H = ctypes.c_void_p
SYMBOLS = {
    "rvs_abi_version": (ctypes.c_int, []),
    "rvs_device_count": (ctypes.c_int, []),
    "rvs_create": (ctypes.c_int, [ctypes.POINTER(RvsRobotDesc), ctypes.POINTER(RvsParams), ctypes.c_int, ctypes.POINTER(_H)]),
    "rvs_destroy": (ctypes.c_int, [_H]),
    "rvs_last_error": (ctypes.c_char_p, [_H]),
    "rvs_set_params": (ctypes.c_int, [_H, ctypes.POINTER(RvsParams)]),
    "rvs_get_params": (ctypes.c_int, [_H, ctypes.POINTER(RvsParams)]),
    "rvs_num_dofs": (ctypes.c_int, [_H]),
    "rvs_get_limits": (ctypes.c_int, [_H, c_f32p, c_f32p]),
    "rvs_set_stream": (ctypes.c_int, [_H, ctypes.c_void_p]),
    "rvs_get_stream": (ctypes.c_void_p, [_H]),
    "rvs_sync": (ctypes.c_int, [_H]),
    "rvs_launch_count": (ctypes.c_int64, [_H]),
    "rvs_measure_fma_peak": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]),
    "rvs_pair_count": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)]),
    "rvs_work_stats": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)]),
    "rvs_set_obstacles": (ctypes.c_int, [_H, c_f32p, c_f32p, c_u8p, c_f32p, ctypes.c_int32]),
    "rvs_update_obstacle_aabbs": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int32]),
    "rvs_num_obstacles": (ctypes.c_int, [_H]),
    "rvs_is_valid": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_is_valid_edges": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_distance": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, ctypes.c_int32, c_f32p, c_f32p, c_f32p]),
    "rvs_underestimation": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_f32p, c_i32p, ctypes.c_int32, c_f32p, c_f32p, c_u8p]),
    "rvs_skeleton": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_f32p]),
    "rvs_enclosing_radii": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_f32p]),
    "rvs_self_collision_capsules": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_set_self_collision": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_float]),
    "rvs_get_self_collision": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)]),
    "rvs_tree_create": (ctypes.c_int, [_H, c_i32p]),
    "rvs_tree_clear": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_tree_release": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_tree_size": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_tree_append": (ctypes.c_int, [_H, ctypes.c_int32, c_f32p, ctypes.c_int32]),
    "rvs_tree_nearest": (ctypes.c_int, [_H, c_i32p, c_f32p, ctypes.c_int32, c_i32p, c_f32p]),
    "rvs_tree_radius": (ctypes.c_int, [_H, ctypes.c_int32, c_f32p, ctypes.c_float, ctypes.c_int32, ctypes.c_int32, c_i32p, c_f32p, c_i32p]),
    "rvs_check_states": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_check_edges": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_self_collision_edges": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int64, c_u8p, c_f32p]),
    "rvs_check_edges_detailed": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int64, c_u8p, c_u8p, c_f32p]),
    "rvs_self_collision_edges_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "rvs_extend_spines": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_float, c_f32p, c_f32p, ctypes.c_int32,
                                         ctypes.POINTER(RvsSpineParams), c_f32p, c_f32p, c_u8p]),
    "rvs_store_reserve": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_store_capacity": (ctypes.c_int, [_H]),
    "rvs_store_distance": (ctypes.c_int, [_H, c_f32p, ctypes.c_int32, c_i32p, c_f32p, c_f32p]),
    "rvs_store_extend_spines": (ctypes.c_int, [_H, c_i32p, c_f32p, ctypes.c_int32, ctypes.POINTER(RvsSpineParams), ctypes.c_int32,
                                               c_f32p, c_i32p, c_f32p, c_u8p]),
    "rvs_store_connect": (ctypes.c_int, [_H, c_i32p, c_f32p, c_i32p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(RvsSpineParams), c_f32p, c_i32p, c_f32p, c_u8p]),
    "rvs_store_read": (ctypes.c_int, [_H, ctypes.c_int32, c_f32p, c_f32p, c_f32p, c_f32p]),
    "rvs_capsule_box_distance": (ctypes.c_int, [_H, c_f32p, c_f32p, c_f32p, ctypes.c_int64, c_f32p]),
    "rvs_capsule_box_collision": (ctypes.c_int, [_H, c_f32p, c_f32p, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_distance_f64": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, ctypes.c_int32, c_f64p, c_f64p, c_f64p]),
    "rvs_is_valid_f64": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, c_u8p]),
    "rvs_is_valid_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "rvs_is_valid_edges_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "rvs_distance_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "rvs_distance_f64_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "rvs_underestimation_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "rvs_generate_configs_dev": (ctypes.c_int, [_H, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "rvs_set_config_sort": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_set_distance_field": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_second_pass_warps": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)]),
    "rvs_set_culling": (ctypes.c_int, [_H, ctypes.c_int32]),
    "rvs_env_set": (ctypes.c_int, [_H, ctypes.POINTER(RvsEnvParams), c_f32p, c_f32p, c_u8p, c_f32p, c_f32p, c_u8p, c_u8p, c_f32p, ctypes.c_int32]),
    "rvs_env_set_random_directions": (ctypes.c_int, [_H, c_f32p, ctypes.c_int32, ctypes.c_int64]),
    "rvs_env_get_random_directions": (ctypes.c_int, [_H, c_f32p, ctypes.POINTER(ctypes.c_uint32)]),
    "rvs_env_step": (ctypes.c_int, [_H, ctypes.c_float, ctypes.c_int32]),
    "rvs_env_get": (ctypes.c_int, [_H, c_f32p, c_f32p]),
    "rvs_motion_validity": (ctypes.c_int, [_H, c_f32p, c_f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, c_u8p, c_i32p]),
    "rvs_motion_validity_points": (ctypes.c_int, [_H, c_f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, c_u8p, c_i32p]),
}

_lib = None


class RvsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rvs error {code}: {message}")
        self.code = code


def load_library(path: Optional[str] = None):
    """dlopen librvs_cuda.so and bind every declared symbol. Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.rvs_abi_version() != 1:
        raise RuntimeError("librvs_cuda.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def _p(a: Optional[np.ndarray], typ):
    return None if a is None else a.ctypes.data_as(typ)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


class Engine:
    """One handle = one robot + one resident obstacle table on one device."""

    def __init__(self, robot: Robot, params: Optional[RvsParams] = None, device: int = 0):
        self._lib = load_library()
        self.robot = robot
        self.params = params or default_params()
        self.n = robot.num_dofs
        self.M = 0
        self.device = device
        self._desc = robot.to_desc()
        h = _H()
        rc = self._lib.rvs_create(ctypes.byref(self._desc), ctypes.byref(self.params), device, ctypes.byref(h))
        if rc != 0:
            raise RvsError(rc, (self._lib.rvs_last_error(None) or b"").decode())
        self._h = h
        # robots::xArm6's constructor switches self-collision checking on (xArm6.cpp:67; the scenario YAML may override it,
        # Scenario.cpp:101-103); EPS_STEP 0.1 is the shipped RRTConnectConfig value.  Precision F64 has no self-collision kernels.
        if robot.self_collision_checking and self.params.precision != PREC_F64:
            self.set_self_collision(True, 0.1)

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RvsError(rc, (self._lib.rvs_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rvs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def is_f64(self) -> bool:
        return self.params.precision == PREC_F64

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.rvs_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._check(self._lib.rvs_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.rvs_launch_count(self._h))

    def pair_count(self, reset: bool = False) -> int:
        """Pair queries the reference's own loop would have made for everything passed to distance*() so far."""
        v = ctypes.c_uint64(0)
        self._check(self._lib.rvs_pair_count(self._h, 1 if reset else 0, ctypes.byref(v)))
        return int(v.value)

    def work_stats(self, reset: bool = False) -> dict:
        """Reference-order pair queries and what the work-queue kernel executed for them (rvs_work_stats)."""
        v = (ctypes.c_uint64 * 4)()
        self._check(self._lib.rvs_work_stats(self._h, 1 if reset else 0, v))
        return {"pair_queries": int(v[0]), "scan_survivors": int(v[1]), "classified": int(v[2]), "edges": int(v[3])}

    def measure_fma_peak(self, fp64: bool = False) -> float:
        """FFMA / DFMA issue peak of this device in TFLOP/s (register-only microbenchmark)."""
        v = ctypes.c_double(0)
        self._check(self._lib.rvs_measure_fma_peak(self._h, 1 if fp64 else 0, ctypes.byref(v)))
        return float(v.value)

    def set_culling(self, on: bool):
        self._check(self._lib.rvs_set_culling(self._h, 1 if on else 0))

    def set_distance_field(self, on: bool):
        """Bootstrap thresholds of large distance batches from a coarse distance field (results identical on / off)."""
        self._check(self._lib.rvs_set_distance_field(self._h, 1 if on else 0))

    def set_config_sort(self, on: bool):
        """Coherent-warp ordering of large batches (results identical on / off)."""
        self._check(self._lib.rvs_set_config_sort(self._h, 1 if on else 0))

    def second_pass_warps(self, reset: bool = False) -> int:
        out = ctypes.c_uint64(0)
        self._check(self._lib.rvs_second_pass_warps(self._h, 1 if reset else 0, ctypes.byref(out)))
        return int(out.value)

    def set_params(self, params: RvsParams):
        self._check(self._lib.rvs_set_params(self._h, ctypes.byref(params)))
        self.params = params

    # -- obstacles ---------------------------------------------------------------------------------------------
    def set_obstacles(self, aabb_min, aabb_max, is_ground=None, min_dist_tol=None):
        mn, mx = _f32(aabb_min, (-1, 3)), _f32(aabb_max, (-1, 3))
        if mn.shape != mx.shape:
            raise ValueError("aabb_min / aabb_max shape mismatch")
        M = mn.shape[0]
        g = None if is_ground is None else np.ascontiguousarray(is_ground, np.uint8)
        t = None if min_dist_tol is None else _f32(min_dist_tol)
        self._check(self._lib.rvs_set_obstacles(self._h, _p(mn, c_f32p), _p(mx, c_f32p), _p(g, c_u8p), _p(t, c_f32p), M))
        self.M = M

    def update_obstacle_aabbs(self, aabb_min, aabb_max):
        mn, mx = _f32(aabb_min, (-1, 3)), _f32(aabb_max, (-1, 3))
        self._check(self._lib.rvs_update_obstacle_aabbs(self._h, _p(mn, c_f32p), _p(mx, c_f32p), mn.shape[0]))

    # -- hot path, host arrays ----------------------------------------------------------------------------------
    def is_valid(self, q) -> np.ndarray:
        q = _f32(q, (-1, self.n))
        out = np.empty(q.shape[0], np.uint8)
        fn = self._lib.rvs_is_valid_f64 if self.is_f64 else self._lib.rvs_is_valid
        self._check(fn(self._h, _p(q, c_f32p), q.shape[0], _p(out, c_u8p)))
        return out

    def is_valid_edges(self, q1, q2) -> np.ndarray:
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        if q1.shape != q2.shape:
            raise ValueError("q1 / q2 shape mismatch")
        out = np.empty(q1.shape[0], np.uint8)
        self._check(self._lib.rvs_is_valid_edges(self._h, _p(q1, c_f32p), _p(q2, c_f32p), q1.shape[0], _p(out, c_u8p)))
        return out

    def distance(self, q, mode: int = DIST_PROFILE):
        """Returns (d_c[N], profile[N,L] or None, nearest[N,M,L,6] or None); float64 arrays on an F64 handle."""
        q = _f32(q, (-1, self.n))
        N = q.shape[0]
        dt, pt, fn = ((np.float64, c_f64p, self._lib.rvs_distance_f64) if self.is_f64
                      else (np.float32, c_f32p, self._lib.rvs_distance))
        d_c = np.empty(N, dt)
        profile = np.empty((N, self.n), dt) if mode >= DIST_PROFILE else None
        nearest = np.zeros((N, self.M, self.n, 6), dt) if mode == DIST_FULL else None
        self._check(fn(self._h, _p(q, c_f32p), N, mode, _p(d_c, pt), _p(profile, pt), _p(nearest, pt)))
        return d_c, profile, nearest

    def underestimation(self, q, planes, plane_idx=None):
        q = _f32(q, (-1, self.n))
        N = q.shape[0]
        planes = _f32(planes, (-1, self.M, self.n, 6))
        idx = None if plane_idx is None else np.ascontiguousarray(plane_idx, np.int32)
        d_c, profile, status = np.empty(N, np.float32), np.empty((N, self.n), np.float32), np.empty(N, np.uint8)
        self._check(self._lib.rvs_underestimation(self._h, _p(q, c_f32p), N, _p(planes, c_f32p), _p(idx, c_i32p), planes.shape[0],
                                                  _p(d_c, c_f32p), _p(profile, c_f32p), _p(status, c_u8p)))
        return d_c, profile, status

    def skeleton(self, q) -> np.ndarray:
        q = _f32(q, (-1, self.n))
        out = np.empty((q.shape[0], self.n + 1, 3), np.float32)
        self._check(self._lib.rvs_skeleton(self._h, _p(q, c_f32p), q.shape[0], _p(out, c_f32p)))
        return out

    def enclosing_radii(self, q) -> np.ndarray:
        q = _f32(q, (-1, self.n))
        out = np.empty((q.shape[0], self.n, self.n + 1), np.float32)
        self._check(self._lib.rvs_enclosing_radii(self._h, _p(q, c_f32p), q.shape[0], _p(out, c_f32p)))
        return out

    def extend_spines(self, q_root, root_profile, root_d_c: float, root_planes, q_targets, spine_params: Optional[RvsSpineParams] = None):
        """Generalized-bur spines from one root towards S targets (rvs_extend_spines). Returns (q_reached[S,n], d_c[S], status[S])."""
        sp = spine_params or RvsSpineParams(5, 1, 1, 0.05)
        q_root, root_profile = _f32(q_root, (self.n,)), _f32(root_profile, (self.n,))
        q_targets, root_planes = _f32(q_targets, (-1, self.n)), _f32(root_planes)
        S = q_targets.shape[0]
        q_out, d_out, status = np.empty((S, self.n), np.float32), np.empty(S, np.float32), np.empty(S, np.uint8)
        self._check(self._lib.rvs_extend_spines(self._h, _p(q_root, c_f32p), _p(root_profile, c_f32p), ctypes.c_float(root_d_c), _p(root_planes, c_f32p),
                                                _p(q_targets, c_f32p), S, ctypes.byref(sp), _p(q_out, c_f32p), _p(d_out, c_f32p), _p(status, c_u8p)))
        return q_out, d_out, status

    # -- device-resident State cache ----------------------------------------------------------------------------------
    def store_reserve(self, capacity: int):
        self._check(self._lib.rvs_store_reserve(self._h, capacity))

    def store_distance(self, q, slots, want_profile: bool = True):
        """computeDistance with nearest points for N configurations, cached in slots[i]. Returns (d_c[N], profile[N,L] or None)."""
        q = _f32(q, (-1, self.n))
        slots = np.ascontiguousarray(slots, np.int32)
        N = q.shape[0]
        d_c = np.empty(N, np.float32)
        profile = np.empty((N, self.n), np.float32) if want_profile else None
        self._check(self._lib.rvs_store_distance(self._h, _p(q, c_f32p), N, _p(slots, c_i32p), _p(d_c, c_f32p), _p(profile, c_f32p)))
        return d_c, profile

    def store_extend_spines(self, root_slots, q_targets, spine_params: Optional[RvsSpineParams] = None, gen2: bool = False):
        """Generalized-bur spines rooted at cached states. Returns (q_layers[S,layers,n], n_layers[S], d_c_last[S], status[S])."""
        sp = spine_params or RvsSpineParams(5, 1, 1, 0.05)
        root_slots = np.ascontiguousarray(root_slots, np.int32)
        q_targets = _f32(q_targets, (-1, self.n))
        S = q_targets.shape[0]
        q_layers = np.zeros((S, sp.num_layers, self.n), np.float32)
        n_layers, d_last, status = np.empty(S, np.int32), np.empty(S, np.float32), np.empty(S, np.uint8)
        self._check(self._lib.rvs_store_extend_spines(self._h, _p(root_slots, c_i32p), _p(q_targets, c_f32p), S, ctypes.byref(sp), 1 if gen2 else 0,
                                                      _p(q_layers, c_f32p), _p(n_layers, c_i32p), _p(d_last, c_f32p), _p(status, c_u8p)))
        return q_layers, n_layers, d_last, status

    def store_connect(self, root_slots, q_targets, new_slots, spine_params: Optional[RvsSpineParams] = None):
        """connectGenSpine's loop for R requests in one call (rvs_store_connect). new_slots: [R, steps] (-1 = fewer steps).
        Returns (q_layers[R,steps,layers,n], n_layers[R,steps] (-1: step not executed), d_c_new[R,steps], status[R,steps])."""
        sp = spine_params or RvsSpineParams(5, 1, 1, 0.05)
        root_slots = np.ascontiguousarray(root_slots, np.int32)
        new_slots = np.ascontiguousarray(new_slots, np.int32).reshape(len(root_slots), -1)
        q_targets = _f32(q_targets, (len(root_slots), self.n))
        R, steps = new_slots.shape
        q_layers = np.zeros((R, steps, sp.num_layers, self.n), np.float32)
        n_layers, d_new, status = np.empty((R, steps), np.int32), np.empty((R, steps), np.float32), np.empty((R, steps), np.uint8)
        self._check(self._lib.rvs_store_connect(self._h, _p(root_slots, c_i32p), _p(q_targets, c_f32p), _p(new_slots, c_i32p), R, steps, ctypes.byref(sp),
                                                _p(q_layers, c_f32p), _p(n_layers, c_i32p), _p(d_new, c_f32p), _p(status, c_u8p)))
        return q_layers, n_layers, d_new, status

    def store_read(self, slot: int):
        """(q[n], d_c, profile[L], planes[M,L,6]) of one cached state."""
        q, prof, planes = np.empty(self.n, np.float32), np.empty(self.n, np.float32), np.empty((self.M, self.n, 6), np.float32)
        d = ctypes.c_float(0)
        self._check(self._lib.rvs_store_read(self._h, slot, _p(q, c_f32p), ctypes.byref(d), _p(prof, c_f32p), _p(planes, c_f32p)))
        return q, float(d.value), prof, planes

    # ---- dynamic environment ---------------------------------------------------------------------
    def env_set(self, ep: RvsEnvParams, dim, pos, vel=None, max_vel=None, is_dynamic=None, is_ground=None, min_dist_tol=None, rotated=None):
        """env::Environment with moving objects; replaces set_obstacles (Environment.cpp:13-109, 194-377)."""
        dim, pos = _f32(dim, (-1, 3)), _f32(pos, (-1, 3))
        M = dim.shape[0]
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8).reshape(M)
        f = lambda a, shape: None if a is None else _f32(a, shape)
        vel, max_vel, tol = f(vel, (M, 3)), f(max_vel, (M,)), f(min_dist_tol, (M,))
        rot, dyn, gr = u8(rotated), u8(is_dynamic), u8(is_ground)
        self._env_keep = (dim, pos, vel, max_vel, tol, rot, dyn, gr)
        self._check(self._lib.rvs_env_set(self._h, ctypes.byref(ep), _p(dim, c_f32p), _p(pos, c_f32p), _p(rot, c_u8p), _p(vel, c_f32p),
                                         _p(max_vel, c_f32p), _p(dyn, c_u8p), _p(gr, c_u8p), _p(tol, c_f32p), M))
        self.M = M

    def env_step(self, delta_time: float, steps: int = 1):
        self._check(self._lib.rvs_env_step(self._h, float(delta_time), int(steps)))

    def env_set_random_directions(self, acc=None, ground_included: int = -1, seed: int = -1):
        """Extra state of Environment::moveRandomDirections: accelerations [M, 3], Environment::ground_included, srand(seed)."""
        acc = None if acc is None else _f32(acc, (self.M, 3))
        self._check(self._lib.rvs_env_set_random_directions(self._h, _p(acc, c_f32p), int(ground_included), int(seed)))

    def env_get_random_directions(self):
        """-> (accelerations [M, 3], rand() calls consumed since the last seed)."""
        acc, draws = np.empty((self.M, 3), np.float32), ctypes.c_uint32(0)
        self._check(self._lib.rvs_env_get_random_directions(self._h, _p(acc, c_f32p), ctypes.byref(draws)))
        return acc, int(draws.value)

    def env_get(self):
        pos, vel = np.empty((self.M, 3), np.float32), np.empty((self.M, 3), np.float32)
        self._check(self._lib.rvs_env_get(self._h, _p(pos, c_f32p), _p(vel, c_f32p)))
        return pos, vel

    def motion_validity(self, q_prev, q_cur, K: int, delta_time: float, commit_edge: int = -1):
        """MotionValidity::check for E motions over one obstacle timeline -> (valid[E] bool, first_invalid[E])."""
        q_prev, q_cur = _f32(q_prev, (-1, self.n)), _f32(q_cur, (-1, self.n))
        E = q_prev.shape[0]
        valid, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        self._check(self._lib.rvs_motion_validity(self._h, _p(q_prev, c_f32p), _p(q_cur, c_f32p), E, int(K), float(delta_time), int(commit_edge),
                                                 _p(valid, c_u8p), _p(first, c_i32p)))
        return valid.astype(bool), first

    def motion_validity_points(self, points, delta_time: float, commit_edge: int = -1):
        """MotionValidity::check(traj_points) for E trajectories of K explicit points (points: [E, K, n]) -> (valid[E] bool, first_invalid[E])."""
        points = np.ascontiguousarray(points, np.float32)
        if points.ndim == 2:
            points = points[None]
        E, K, n = points.shape
        if n != self.n:
            raise ValueError("points must be [E, K, num_dofs]")
        valid, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        self._check(self._lib.rvs_motion_validity_points(self._h, _p(points, c_f32p), E, int(K), float(delta_time), int(commit_edge),
                                                        _p(valid, c_u8p), _p(first, c_i32p)))
        return valid.astype(bool), first

    def self_collision_capsules(self, q) -> np.ndarray:
        q = _f32(q, (-1, self.n))
        out = np.empty(q.shape[0], np.uint8)
        self._check(self._lib.rvs_self_collision_capsules(self._h, _p(q, c_f32p), q.shape[0], _p(out, c_u8p)))
        return out

    def set_self_collision(self, enabled: bool, eps_step: float = 0.1) -> None:
        """AbstractRobot::setSelfCollisionChecking (Scenario.cpp:101-103) + RRTConnectConfig::EPS_STEP (xArm6.cpp:340-341).
        While on, the spine entry points apply the reference's self-collision truncation (RBTConnect.cpp:129-131)."""
        self._check(self._lib.rvs_set_self_collision(self._h, 1 if enabled else 0, float(eps_step)))

    def self_collision_edges(self, q1, q2):
        """xArm6::checkSelfCollision(q1, q2&), capsule stage (xArm6.cpp:284-368) -> (hit[E], q2 after the call)."""
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        hit, out = np.empty(q1.shape[0], np.uint8), np.empty_like(q2)
        self._check(self._lib.rvs_self_collision_edges(self._h, _p(q1, c_f32p), _p(q2, c_f32p), q1.shape[0], _p(hit, c_u8p), _p(out, c_f32p)))
        return hit, out

    # -- device-resident planner trees (base::Tree's nearest neighbour, RRTx::findNeighbors) ---------------------
    def tree_create(self) -> int:
        t = ctypes.c_int32(-1)
        self._check(self._lib.rvs_tree_create(self._h, ctypes.byref(t)))
        return int(t.value)

    def tree_append(self, tree: int, q) -> None:
        q = _f32(q, (-1, self.n))
        self._check(self._lib.rvs_tree_append(self._h, tree, _p(q, c_f32p), q.shape[0]))

    def tree_clear(self, tree: int) -> None:
        self._check(self._lib.rvs_tree_clear(self._h, tree))

    def tree_release(self, tree: int) -> None:
        self._check(self._lib.rvs_tree_release(self._h, tree))

    def tree_size(self, tree: int) -> int:
        return int(self._lib.rvs_tree_size(self._h, tree))

    def tree_nearest(self, queries, trees=None):
        """Tree::getNearestState for Q queries in one launch -> (index[Q], squared distance[Q])."""
        queries = _f32(queries, (-1, self.n))
        Q = queries.shape[0]
        idx, d2 = np.empty(Q, np.int32), np.empty(Q, np.float32)
        tr = None if trees is None else np.ascontiguousarray(trees, np.int32)
        self._check(self._lib.rvs_tree_nearest(self._h, _p(tr, c_i32p), _p(queries, c_f32p), Q, _p(idx, c_i32p), _p(d2, c_f32p)))
        return idx, d2

    def tree_radius(self, tree: int, query, radius: float, exclude: int = -1, capacity: int = 4096):
        """RRTx::findNeighbors: (indices in tree order, distances) of the states within `radius` (ss->getNorm) of `query`."""
        query = _f32(query, (self.n,))
        while True:
            idx, dist, cnt = np.empty(capacity, np.int32), np.empty(capacity, np.float32), ctypes.c_int32(0)
            self._check(self._lib.rvs_tree_radius(self._h, tree, _p(query, c_f32p), float(radius), exclude, capacity, _p(idx, c_i32p), _p(dist, c_f32p), ctypes.byref(cnt)))
            if cnt.value <= capacity:
                return idx[:cnt.value].copy(), dist[:cnt.value].copy()
            capacity = int(cnt.value)

    def check_edges_detailed(self, q1, q2):
        """isValid(q1,q2) and checkSelfCollision(q1,q2) in one call, reported apart -> (obstacle_ok[E] bool, self_hit[E] bool, q2 after the call)."""
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        ok, hit, out = np.empty(q1.shape[0], np.uint8), np.empty(q1.shape[0], np.uint8), np.empty_like(q2)
        self._check(self._lib.rvs_check_edges_detailed(self._h, _p(q1, c_f32p), _p(q2, c_f32p), q1.shape[0], _p(ok, c_u8p), _p(hit, c_u8p), _p(out, c_f32p)))
        return ok.astype(bool), hit.astype(bool), out

    def check_states(self, q) -> np.ndarray:
        """isValid(q) && !robot->checkSelfCollision(q) in one call (RRTConnect.cpp:13-15, MotionValidity.cpp:42)."""
        q = _f32(q, (-1, self.n))
        out = np.empty(q.shape[0], np.uint8)
        self._check(self._lib.rvs_check_states(self._h, _p(q, c_f32p), q.shape[0], _p(out, c_u8p)))
        return out.astype(bool)

    def check_edges(self, q1, q2) -> np.ndarray:
        """isValid(q1, q2) && !robot->checkSelfCollision(q1, q2) in one call (RRTConnect.cpp:100, RRTx.cpp:467-468)."""
        q1, q2 = _f32(q1, (-1, self.n)), _f32(q2, (-1, self.n))
        out = np.empty(q1.shape[0], np.uint8)
        self._check(self._lib.rvs_check_edges(self._h, _p(q1, c_f32p), _p(q2, c_f32p), q1.shape[0], _p(out, c_u8p)))
        return out.astype(bool)

    def capsule_box_distance(self, AB, radius, obs) -> np.ndarray:
        """Batched CollisionAndDistance::distanceCapsuleToBox: rows (d, has_points, P1 xyz, P2 xyz)."""
        AB, obs = _f32(AB, (-1, 6)), _f32(obs, (-1, 6))
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (AB.shape[0],)))
        out = np.empty((AB.shape[0], 8), np.float32)
        self._check(self._lib.rvs_capsule_box_distance(self._h, _p(AB, c_f32p), _p(r, c_f32p), _p(obs, c_f32p), AB.shape[0], _p(out, c_f32p)))
        return out

    def capsule_box_collision(self, AB, radius, obs) -> np.ndarray:
        """Batched CollisionAndDistance::collisionCapsuleToBox."""
        AB, obs = _f32(AB, (-1, 6)), _f32(obs, (-1, 6))
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float32), (AB.shape[0],)))
        out = np.empty(AB.shape[0], np.uint8)
        self._check(self._lib.rvs_capsule_box_collision(self._h, _p(AB, c_f32p), _p(r, c_f32p), _p(obs, c_f32p), AB.shape[0], _p(out, c_u8p)))
        return out

    # -- device pointers (ints, e.g. torch.Tensor.data_ptr()), stream ordered ----------------------------------------
    def distance_dev(self, d_q: int, N: int, mode: int, d_dc: int, d_profile: int = 0, d_nearest: int = 0):
        fn = self._lib.rvs_distance_f64_dev if self.is_f64 else self._lib.rvs_distance_dev
        self._check(fn(self._h, d_q, N, mode, d_dc, d_profile or None, d_nearest or None))

    def is_valid_dev(self, d_q: int, N: int, d_valid: int):
        self._check(self._lib.rvs_is_valid_dev(self._h, d_q, N, d_valid))

    def is_valid_edges_dev(self, d_q1: int, d_q2: int, E: int, d_valid: int):
        self._check(self._lib.rvs_is_valid_edges_dev(self._h, d_q1, d_q2, E, d_valid))

    def underestimation_dev(self, d_q: int, N: int, d_planes: int, d_plane_idx: int, R: int, d_dc: int, d_profile: int = 0,
                            d_status: int = 0):
        self._check(self._lib.rvs_underestimation_dev(self._h, d_q, N, d_planes, d_plane_idx or None, R, d_dc, d_profile or None,
                                                      d_status or None))

    def generate_configs_dev(self, seed: int, first: int, N: int, d_q: int):
        self._check(self._lib.rvs_generate_configs_dev(self._h, ctypes.c_uint64(seed), first, N, d_q))
