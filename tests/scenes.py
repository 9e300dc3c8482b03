"""Shared scene builders for the tests (inputs only; all synthetic, nothing read from /root/reference at run time)."""
import numpy as np

from rpmplv2_b200 import robots, synthetic


def xarm6_scene(M, dist="D2", r_min=0.0):
    return (robots.xarm6(),) + synthetic.sweep_obstacles(M, dist, r_min=r_min)


def planar2_scenario_test():
    """planar_2dof scenario_test geometry: four 0.5 x 0.5 x 0.1 boxes at (+-1, +-1, 0)
    (/root/reference/data/planar_2dof/scenario_test/scenario_test.yaml:1-13); no ground, default radii."""
    dim = np.tile(np.array([[0.5, 0.5, 0.1]], np.float32), (4, 1))
    pos = np.array([[1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0]], np.float32)
    mn, mx = synthetic.box_aabbs(dim, pos)
    return robots.planar(2), mn, mx, np.zeros(4, np.uint8), np.full(4, 1000.0, np.float32)


def xarm6_scenario1():
    """xarm6 scenario1: ground + four 0.5 x 0.5 x 1 pillars (/root/reference/data/xarm6/scenario1/scenario1.yaml:1-18)."""
    dim = np.array([[1.5, 1.5, 0.1]] + [[0.5, 0.5, 1.0]] * 4, np.float32)
    pos = np.array([[0, 0, -0.05], [0.55, 0.55, 0.5], [-0.55, 0.55, 0.5], [0.55, -0.55, 0.5], [-0.55, -0.55, 0.5]], np.float32)
    mn, mx = synthetic.box_aabbs(dim, pos)
    return robots.xarm6(), mn, mx, np.array([1, 0, 0, 0, 0], np.uint8), np.array([0.05, 1000, 1000, 1000, 1000], np.float32)


def spatial10_scenario1():
    """spatial_10dof scenario1: the same ground + four pillars around a 10-DOF spatial chain, radii 0.025, ground_included 2
    (/root/reference/data/spatial_10dof/scenario1/scenario1.yaml:1-40)."""
    _, mn, mx, g, tol = xarm6_scenario1()
    return robots.spatial(10), mn, mx, g, tol


def chain_scene(rb, M, seed=77):
    """Random boxes around a planar / spatial chain whose reach is ~ n * link length."""
    rng = np.random.default_rng(seed)
    reach = float(np.abs(rb.xyz).sum())
    c = rng.uniform(-reach, reach, (M, 3)).astype(np.float32)
    if rb.kind == robots.ROBOT_PLANAR:
        c[:, 2] = rng.uniform(-0.05, 0.05, M)
    d = rng.uniform(0.05, 0.3, (M, 3)).astype(np.float32) * np.float32(max(reach, 0.5))
    keep = np.linalg.norm(c, axis=1) > 0.25 * reach + 0.2
    c[~keep] += np.sign(c[~keep] + 1e-3) * (0.3 * reach + 0.2)
    mn, mx = synthetic.box_aabbs(d, c)
    ground = np.zeros(M, np.uint8)
    tol = np.full(M, 1000.0, np.float32)
    if rb.ground_included > 0:
        gmn, gmx = synthetic.box_aabbs(np.array([[3.0, 3.0, 0.1]], np.float32), np.array([[0, 0, -0.05]], np.float32))
        mn[0], mx[0], ground[0], tol[0] = gmn[0], gmx[0], 1, 0.05
    return rb, mn, mx, ground, tol


WS_CENTER = (0.0, 0.0, 0.267)


def dynamic_scene(rb, M_dyn, seed, speed=0.8, size=0.3, r_lo=0.45, r_hi=0.95):
    """Ground box (static) + M_dyn cubes of side `size` in a shell around WS_center with random velocities (|v| <= max_vel = speed),
    the shape of data/xarm6/scenario_random_obstacles and data/spatial_22dof/scenario_random_obstacles."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(M_dyn, 3))
    d[:, 2] = np.abs(d[:, 2])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(r_lo, r_hi, (M_dyn, 1))
    pos = np.vstack([[0.0, 0.0, -0.05], np.asarray(WS_CENTER) + d * r]).astype(np.float32)
    dim = np.vstack([[1.5, 1.5, 0.1], np.full((M_dyn, 3), size)]).astype(np.float32)
    v = rng.normal(size=(M_dyn, 3))
    v *= (rng.uniform(0.3, 1.0, (M_dyn, 1)) * speed) / np.linalg.norm(v, axis=1, keepdims=True)
    vel = np.vstack([[0, 0, 0], v]).astype(np.float32)
    max_vel = np.array([0.0] + [speed] * M_dyn, np.float32)
    dyn = np.array([0] + [1] * M_dyn, np.uint8)
    ground = np.array([1] + [0] * M_dyn, np.uint8)
    tol = np.array([0.05] + [1000.0] * M_dyn, np.float32)
    return dict(dim=dim, pos=pos, vel=vel, max_vel=max_vel, is_dynamic=dyn, is_ground=ground, min_dist_tol=tol)
