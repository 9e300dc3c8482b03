"""The C++ host layer above the C ABI: base::RealVectorSpaceCUDA (rpmplv2_b200/cpp) and the URDF chain reader.

The class is compiled against the reference's REAL headers (State.h, StateSpace.h, RealVectorSpace.h, AbstractRobot.h, Environment.h,
Object.h, Box.h, RRTConnectConfig.h, ... from /root/reference/include; Eigen / FCL / KDL / yaml-cpp through the work-alikes of
oracle/eigen_shim + oracle/env_shim) into oracle/_ref/libdropin_ref.so by `make -C oracle ref` — the same library tests/test_dropin.py
uses to run the reference's planners on top of the class.  The drv_* entry points of oracle/rvs_ref_capi.cpp drive the class the way a
maintainer's code would (URDF constructor, single-query overrides, batched overloads, State cache rules, staleness guard, dynamic
environment).  The library is built where /root/reference exists and travels to the GPU box prebuilt.
"""
import ctypes
import os

import numpy as np
import pytest

import scenes
from conftest import REFERENCE_DATA, ROOT
from rpmplv2_b200 import robots, synthetic

SO = os.path.join(ROOT, "oracle", "_ref", "libdropin_ref.so")
f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)


@pytest.fixture(scope="session")
def driver():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libdropin_ref.so is built where /root/reference exists (make -C oracle ref)")
    lib = ctypes.CDLL(SO)                     # resolves librvs_cuda.so through its rpath; loading needs no GPU
    lib.drv_last_error.restype = ctypes.c_char_p
    return lib


def _write_urdf(tmp_path, rb):
    p = tmp_path / f"{rb.name}.urdf"
    p.write_text(robots.to_urdf(rb))
    return str(p)


def _desc_equal(a, b, n):
    assert a.kind == b.kind and a.num_dofs == b.num_dofs
    for k in range(n + 1):
        assert list(a.joint_xyz[k]) == list(b.joint_xyz[k]) and list(a.joint_rpy[k]) == list(b.joint_rpy[k])
    for k in range(n):
        assert list(a.joint_axis[k]) == list(b.joint_axis[k])
        assert a.limit_lo[k] == b.limit_lo[k] and a.limit_hi[k] == b.limit_hi[k]


@pytest.mark.parametrize("rb", [robots.xarm6(), robots.planar(2), robots.planar(10), robots.spatial(22)], ids=lambda r: r.name)
def test_cpp_urdf_reader_round_trip(driver, tmp_path, rb):
    """class builds (compile check of RealVectorSpaceCUDA against the reference's API surface) and the C++ URDF reader agrees with
    the built-in chain tables and with the Python reader."""
    path = _write_urdf(tmp_path, rb)
    got = robots.RvsRobotDesc()
    assert driver.drv_parse_urdf(path.encode(), rb.name.encode(), ctypes.byref(got)) == 0, driver.drv_last_error()
    _desc_equal(got, rb.to_desc(), rb.num_dofs)
    py = robots.from_urdf(path, rb.name)
    assert np.array_equal(py.xyz, rb.xyz) and np.array_equal(py.rpy, rb.rpy) and np.array_equal(py.axis, rb.axis) and np.array_equal(py.limits, rb.limits)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference data only exists in the build container")
def test_cpp_urdf_reader_on_the_reference_urdfs(driver):
    for rtype, rel in [("xarm6", "xarm6/xarm6.urdf"), ("planar_2dof", "planar_2dof/planar_2dof.urdf"), ("planar_10dof", "planar_10dof/planar_10dof.urdf"),
                       ("spatial_10dof", "spatial_10dof/spatial_10dof.urdf"), ("spatial_22dof", "spatial_22dof/spatial_22dof.urdf")]:
        got = robots.RvsRobotDesc()
        assert driver.drv_parse_urdf(os.path.join(REFERENCE_DATA, rel).encode(), rtype.encode(), ctypes.byref(got)) == 0
        _desc_equal(got, robots.by_type(rtype).to_desc(), got.num_dofs)


def test_cpp_urdf_reader_rejects_malformed_input(driver, tmp_path):
    p = tmp_path / "bad.urdf"
    p.write_text('<robot name="x"><joint name="j" type="revolute"><parent link="base_link"/><child link="l1"/></joint></robot>')
    got = robots.RvsRobotDesc()
    assert driver.drv_parse_urdf(str(p).encode(), b"planar_2dof", ctypes.byref(got)) == 1
    assert driver.drv_parse_urdf(b"/nonexistent.urdf", b"xarm6", ctypes.byref(got)) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["xarm6", "planar2", "spatial10"])
def test_cpp_class_matches_oracle_and_keeps_reference_state_semantics(driver, ob, tmp_path, case):
    if case == "xarm6":
        rb, mn, mx, g, tol = scenes.xarm6_scene(30, "D2", 0.45)
    elif case == "planar2":
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    else:
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(10), 6)
    n, M, N = rb.num_dofs, len(g), 300
    q = synthetic.configs(rb, N)
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(mn, mx, g, tol)
    d_o, p_o, near_o = o.distance(q, robots.DIST_FULL)
    root = int(np.flatnonzero(d_o > 0)[0])
    q[0], q[root] = q[root].copy(), q[0].copy()                     # make configuration 0 the (collision-free) root
    d_o, p_o, near_o = o.distance(q, robots.DIST_FULL)
    valid = np.zeros(N, np.uint8); valid_b = np.zeros(N, np.uint8); has_near = np.zeros(N, np.uint8); is_real = np.zeros(N, np.uint8)
    d = np.zeros(N, np.float32); d_b = np.zeros(N, np.float32); prof = np.zeros((N, n), np.float32); near = np.zeros((N, M, n, 6), np.float32)
    ev = np.zeros(N - 1, np.uint8); ev_b = np.zeros(N - 1, np.uint8); un = np.zeros(N, np.float32); un_b = np.zeros(N, np.float32)
    P = lambda a, t: a.ctypes.data_as(t)
    desc = rb.to_desc()
    rc = driver.drv_run(_write_urdf(tmp_path, rb).encode(), ctypes.byref(desc), P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), M, ctypes.c_float(1e-4), 10, P(q, f32p), N, 0,
                        P(valid, u8p), P(d, f32p), P(prof, f32p), P(near, f32p), P(has_near, u8p), P(d_b, f32p), P(valid_b, u8p), P(ev, u8p), P(ev_b, u8p),
                        P(un, f32p), P(un_b, f32p), P(is_real, u8p))
    assert rc == 0, driver.drv_last_error()
    assert (valid == o.is_valid(q)).all() and (valid_b == valid).all()
    assert ((d == 0) == (d_o == 0)).all() and np.abs(d - d_o).max() <= 1e-4 and (d_b == d).all()
    assert np.abs(prof - p_o).max() <= 1e-4
    free = d_o > 0
    assert (has_near == free).all()                                   # nearest_points == nullptr exactly after a collision
    assert np.abs(near[free] - near_o[free]).max() <= 1e-4
    e_o = o.is_valid_edges(q[:-1], q[1:])
    assert (ev == e_o).all() and (ev_b == e_o).all()
    du_o, _, st_o = o.underestimation(q, near_o[0:1], np.zeros(N, np.int32))
    assert np.abs(un - du_o).max() <= 1e-4 and (un_b == un).all()
    assert ((un == 0) == (st_o == 1)).all()
    assert (is_real[st_o == 0] == 0).all() and (is_real[st_o == 1] == 1).all()      # accepted: is_real_d_c = false; rejected: untouched


@pytest.mark.gpu
@pytest.mark.parametrize("law", ["light_directions", "straight"])
def test_cpp_class_motion_check_moves_the_environment_like_the_reference_loop(driver, ob, tmp_path, law):
    """attachDynamicEnvironment + checkMotion / checkMotionBatch (row f-2) against the oracle's MotionValidity::check."""
    from rpmplv2_b200.engine import MOTION_TYPES, env_params
    rb = robots.xarm6()
    sc = scenes.dynamic_scene(rb, 5, 41)
    n, M, T, K, dt = rb.num_dofs, len(sc["max_vel"]), 24, 10, 0.005
    o = ob.Oracle(rb, robots.default_params())
    ep = env_params(law, ws_center=scenes.WS_CENTER, ws_radius=1.0, base_radius=0.3, robot_max_vel=1.5)
    o.env_set(ep, **sc)
    cand = synthetic.configs(rb, 400, seed=0xC3)
    d0 = o.distance(cand)[0]
    q = synthetic.configs(rb, T + 1, seed=0xC4)
    q[0] = cand[int(np.argmax(d0))]                                                  # start with clearance, then a slow random walk
    q[1:] = q[:1] + np.cumsum((q[1:] - q[:-1]) * np.float32(0.03), axis=0)
    ws = np.array(list(scenes.WS_CENTER) + [1.0, 0.3, 1.5], np.float32)
    valid = np.zeros(T, np.uint8); passed = np.zeros(T, np.int32); pos_after = np.zeros((T, M, 3), np.float32); bvalid = np.zeros(T, np.uint8)
    P = lambda a, t: a.ctypes.data_as(t)
    i32p = ctypes.POINTER(ctypes.c_int32)
    desc = rb.to_desc()
    rc = driver.drv_motion(_write_urdf(tmp_path, rb).encode(), ctypes.byref(desc), P(sc["dim"], f32p), P(sc["pos"], f32p), P(sc["vel"], f32p), P(sc["max_vel"], f32p), P(sc["is_dynamic"], u8p),
                           P(sc["is_ground"], u8p), P(sc["min_dist_tol"], f32p), M, MOTION_TYPES[law], P(ws, f32p), P(q, f32p), T, K, ctypes.c_float(dt),
                           P(valid, u8p), P(passed, i32p), P(pos_after, f32p), P(bvalid, u8p))
    assert rc == 0, driver.drv_last_error()
    for t in range(T):
        v, f = o.motion_validity(q[t:t + 1], q[t + 1:t + 2], K, dt, commit_edge=0)
        assert bool(valid[t]) == bool(v[0]), t
        assert passed[t] == (K if v[0] else f[0] - 1)
        assert np.array_equal(pos_after[t], o.env_get()[0]), t                          # env::Object positions mirrored back
    vb, _ = o.motion_validity(q[:-1], q[1:], K, dt, commit_edge=-1)
    assert np.array_equal(bvalid.astype(bool), vb)
    assert valid.sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("self_collision", [0, 1])
def test_cpp_class_spine_sets_and_self_collision_vs_oracle(driver, ob, tmp_path, self_collision):
    """extendSpinesBatch / checkEdgesBatch / checkSelfCollisionEdges of the class (xArm6) against the oracle, with the robot's
    self_collision_checking switch mirrored into the engine by the class itself."""
    rb, mn, mx, g, tol = scenes.xarm6_scene(40, "D2", 0.45)
    tol = np.full_like(tol, 1000.0)
    n, M, S, E = rb.num_dofs, len(g), 48, 300
    o = ob.Oracle(rb, robots.default_params())
    o.set_self_collision(bool(self_collision), 0.1)
    o.set_obstacles(mn, mx, g, tol)
    cand = synthetic.configs(rb, 400)
    d, p, near = o.distance(cand, robots.DIST_FULL)
    k = int(np.argmax(d))
    targets = synthetic.configs(rb, S, seed=0x77)
    sp = robots.RvsSpineParams(5, 2, 1, 0.03)
    q_want, d_want, st_want = o.extend_spines(cand[k], p[k], float(d[k]), near[k], targets, sp)
    e1, e2 = synthetic.edges(rb, E, 0.6, seed=0x78)
    ok_want = o.is_valid_edges(e1, e2).astype(bool)
    hit_want, q2_want = o.self_collision_edges(e1, e2)
    ok_want &= hit_want == 0
    P = lambda a, t: a.ctypes.data_as(t)
    desc = rb.to_desc()
    q_got, d_got, st_got = np.zeros((S, n), np.float32), np.zeros(S, np.float32), np.zeros(S, np.uint8)
    ok_got, hit_got, q2_got = np.zeros(E, np.uint8), np.zeros(E, np.uint8), np.zeros((E, n), np.float32)
    root = np.ascontiguousarray(cand[k])
    rc = driver.drv_spines(_write_urdf(tmp_path, rb).encode(), ctypes.byref(desc), self_collision, P(mn, f32p), P(mx, f32p), P(g, u8p), P(tol, f32p), M,
                           P(root, f32p), P(targets, f32p), S, sp.num_layers, sp.num_iter_spine, sp.use_expanded_bubble, ctypes.c_float(sp.d_crit),
                           P(q_got, f32p), P(d_got, f32p), P(st_got, u8p), P(e1, f32p), P(e2, f32p), E, P(ok_got, u8p), P(hit_got, u8p), P(q2_got, f32p))
    assert rc == 0, driver.drv_last_error()
    same = st_got == st_want
    assert same.mean() >= 0.95                                        # a step within an ulp of 1 may flip Advanced / Reached (device FK vs host FK)
    assert np.abs(q_got[same] - q_want[same]).max() <= 1e-4
    live = same & (d_want > 0)
    assert np.abs(d_got[live] - d_want[live]).max() <= 1e-4
    assert (ok_got.astype(bool) == ok_want).mean() >= 0.99
    assert (hit_got == hit_want).mean() >= 0.99
    agree = hit_got == hit_want
    assert (np.abs(q2_got[agree] - q2_want[agree]).max(axis=1) <= 1e-5).mean() >= 0.99
    if self_collision:
        assert hit_want.sum() > 10
    else:
        assert hit_got.sum() == 0 and np.array_equal(q2_got, e2)
