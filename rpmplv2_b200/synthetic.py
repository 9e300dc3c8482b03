"""Stateless synthetic workloads (SURVEY.md §8(d), S1 "sweep" and S2 "planner-shaped").

Everything is a pure function of (seed, index) through splitmix64, so the CPU oracle, the host harness and the
device-side generator (`rvs_generate_configs_dev`) produce bit-identical inputs without shipping arrays around.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

from .robots import Robot

SEED_CONFIGS = 0x5EED0001
SEED_OBSTACLES = 0x5EED0002
SEED_ROTATIONS = 0x5EED0003
SEED_EDGES = 0x5EED0010

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def unit24(seed: int, counter: np.ndarray) -> np.ndarray:
    """Top 24 bits of splitmix64(seed ^ counter) scaled to [0, 1) as float32 (exact)."""
    h = splitmix64(np.uint64(seed) ^ np.asarray(counter, np.uint64))
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def configs(robot: Robot, N: int, seed: int = SEED_CONFIGS, first: int = 0) -> np.ndarray:
    """q[i][j] = lo_j + (hi_j - lo_j) * u24(seed ^ (i*32 + j)), all in float32 (matches the device generator)."""
    n = robot.num_dofs
    i = np.arange(first, first + N, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    u = unit24(seed, i * np.uint64(32) + j)
    lo, hi = robot.limits[:, 0][None, :], robot.limits[:, 1][None, :]
    return (lo + (hi - lo) * u).astype(np.float32)


def box_aabbs(dim: np.ndarray, pos: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """env::Box + FCL computeAABB for an identity rotation: pos -/+ 0.5*dim in float32
    (/root/reference/src/environments/objects/Box.cpp:3-8)."""
    dim, pos = np.asarray(dim, np.float32), np.asarray(pos, np.float32)
    half = np.float32(0.5) * dim
    return (pos + (-half)).astype(np.float32), (pos + half).astype(np.float32)


def rotated_box_aabbs(dim: np.ndarray, pos: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """FCL CollisionObject::computeAABB for a NON-identity rotation: a cube of half-size |0.5*dim| (the bounding
    sphere radius) around the centre. The rotation itself does not matter."""
    dim, pos = np.asarray(dim, np.float32), np.asarray(pos, np.float32)
    half = np.float32(0.5) * dim
    rad = np.sqrt((half * half).sum(axis=-1, keepdims=True)).astype(np.float32)
    return (pos - rad).astype(np.float32), (pos + rad).astype(np.float32)


def sweep_obstacles(M: int, dist: str = "D2", seed: int = SEED_OBSTACLES, ws_center=(0.0, 0.0, 0.267), ws_radius: float = 1.5,
                    ground: bool = True, rotated: bool = False, base_clearance: float = 0.2, r_min: float = 0.0, r_max: float = None):
    """Obstacle table of the S1 sweep. Object 0 is the ground box (dim 1.5 x 1.5 x 0.1 at z = -0.05, min_dist_tol 0.05,
    label "ground"); the others have centres drawn like initRandomObstacles (/root/reference/apps/CommonFunctions.h:95-100):
    r = u*R, phi = 2*pi*u, theta = pi*u around WS_center. Sizes: D1 = 0.01^3 cubes, D2 = each side uniform [0.05, 0.30].
    `r_min` > 0 draws r = r_min + u*(R - r_min) instead (a shell: obstacles kept out of the inner workspace so that most
    random configurations stay collision-free and every pair goes through the narrow phase); `r_max` < R keeps them inside a
    smaller sphere (near field: every obstacle within the arm's reach).
    Returns (aabb_min[M,3], aabb_max[M,3], is_ground[M], min_dist_tol[M])."""
    if M < 1:
        raise ValueError("M >= 1")
    if dist not in ("D1", "D2"):
        raise ValueError("dist must be D1 or D2")
    k = np.arange(M, dtype=np.uint64)
    pos = np.zeros((M, 3), np.float32)
    dim = np.zeros((M, 3), np.float32)
    todo = np.ones(M, bool)
    # Like initRandomObstacles, a candidate that would make the (fixed) base of the robot collide is drawn again
    # (CommonFunctions.h:102-110 removes obstacles for which isValid(start/goal) fails): keep every box at least
    # `base_clearance` away from the base column (0,0,0)-(0,0,WS_center.z).
    for attempt in range(64):
        if not todo.any():
            break
        ctr = (k[todo] + np.uint64(attempt) * np.uint64(1 << 32)) * np.uint64(8)
        u = [unit24(seed, ctr + np.uint64(c)).astype(np.float64) for c in range(6)]
        r, phi, theta = r_min + u[0] * ((ws_radius if r_max is None else r_max) - r_min), 2.0 * math.pi * u[1], math.pi * u[2]
        p = np.stack([ws_center[0] + r * np.cos(phi) * np.sin(theta), ws_center[1] + r * np.sin(phi) * np.sin(theta),
                      ws_center[2] + r * np.cos(theta)], axis=1).astype(np.float32)
        if dist == "D1":
            d = np.full(p.shape, 0.01, np.float32)
        else:
            d = np.stack([0.05 + 0.25 * u[3], 0.05 + 0.25 * u[4], 0.05 + 0.25 * u[5]], axis=1).astype(np.float32)
        lo, hi = box_aabbs(d, p)
        gap_xy = np.maximum(np.maximum(lo[:, :2], -hi[:, :2]), 0.0)                  # distance of the z axis to the box, in xy
        gap_z = np.maximum(np.maximum(lo[:, 2] - ws_center[2], 0.0 - hi[:, 2]), 0.0)   # and along z to the column [0, WS_center.z]
        ok = np.sqrt((gap_xy ** 2).sum(1) + gap_z ** 2) >= base_clearance
        idx = np.flatnonzero(todo)[ok]
        pos[idx], dim[idx] = p[ok], d[ok]
        todo[idx] = False
    if todo.any():
        raise RuntimeError("could not place all obstacles clear of the robot base")
    is_ground = np.zeros(M, np.uint8)
    tol = np.full(M, 1000.0, np.float32)
    if ground:
        dim[0], pos[0] = (1.5, 1.5, 0.1), (0.0, 0.0, -0.05)
        is_ground[0], tol[0] = 1, 0.05
    mn, mx = box_aabbs(dim, pos)
    if rotated:
        rmn, rmx = rotated_box_aabbs(dim, pos)
        keep = is_ground.astype(bool)[:, None]
        mn, mx = np.where(keep, mn, rmn), np.where(keep, mx, rmx)
    return mn, mx, is_ground, tol


def edges(robot: Robot, E: int, length: float = 0.1, seed: int = SEED_EDGES):
    """S2(a): random edges of C-space length `length` (RRTConnectConfig::EPS_STEP) starting at random configurations."""
    q1 = configs(robot, E, seed)
    d = configs(robot, E, seed + 1) - configs(robot, E, seed + 2)
    norm = np.sqrt((d.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    q2 = (q1 + (d / np.maximum(norm, 1e-9) * length)).astype(np.float32)
    return q1, q2


def pair_queries(robot: Robot, is_ground: np.ndarray) -> int:
    """Capsule-obstacle pair queries per configuration: L*M minus the ground pairs the reference skips
    (RealVectorSpace.cpp:285-286,336-341)."""
    L, M = robot.num_links, int(len(is_ground))
    skipped = int(np.count_nonzero(is_ground)) * min(int(robot.ground_included), L)
    return L * M - skipped


def horizon_states(robot: Robot, q_current: np.ndarray, H: int, seed: int, delta: float = math.pi) -> np.ndarray:
    """H states RBTConnectConfig::DELTA away from q_current in random directions, clipped to the joint limits: the shape of DRGBT's
    lateral / random horizon states (/root/reference/src/planners/drbt/DRGBT.cpp:358-420)."""
    rng = np.random.default_rng(seed)
    step = rng.normal(0, 1.0, (H, robot.num_dofs)).astype(np.float32)
    step *= np.float32(delta) / np.linalg.norm(step, axis=1, keepdims=True)
    return np.clip(q_current + step, robot.limits[:, 0], robot.limits[:, 1]).astype(np.float32)
