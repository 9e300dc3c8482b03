"""Dynamic environment: Environment::updateEnvironment's motion laws and MotionValidity::check.

CPU part: the oracle restatement (oracle/rvs_oracle_env.hpp) against closed-form / invariant known answers and against an
independent composition of its own primitives.  GPU part (-m gpu): rvs_env_* / rvs_motion_validity through the C ABI against the
oracle on the same seeded inputs — bit-exact for the straight, two_tunnels and light_directions laws (add / mul / div / sqrt only),
1e-5 m for the circular law (atan2f / sinf / cosf differ in the last ulp between CUDA and glibc).
Reference: /root/reference/src/environments/Environment.cpp:194-377, src/planners/trajectory/MotionValidity.cpp:30-51.
"""
import os

import numpy as np
import pytest

from rpmplv2_b200 import robots, synthetic
from rpmplv2_b200.engine import env_params
from conftest import ROOT
from scenes import WS_CENTER, dynamic_scene

LAWS = ["straight", "circular", "two_tunnels", "light_directions"]


def params_for(law):
    return env_params(law, ws_center=WS_CENTER, ws_radius=1.0, base_radius=0.3, robot_max_vel=1.5)


def random_motions(rb, E, seed, length=0.3):
    q1, q2 = synthetic.edges(rb, E, length=length, seed=seed)
    return q1, q2


# ---------------------------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_straight_law_is_the_float_recurrence(ob):
    sc = dynamic_scene(robots.xarm6(), 4, 1)
    o = ob.Oracle(robots.xarm6())
    o.env_set(params_for("straight"), **sc)
    dt, steps = np.float32(0.013), 57
    o.env_step(float(dt), steps)
    pos, vel = o.env_get()
    want = sc["pos"].copy()
    for _ in range(steps):
        want[1:, 0] = want[1:, 0] - dt * sc["max_vel"][1:]                # pos.x() -= delta_time * max_vel   (Environment.cpp:233)
    assert np.array_equal(pos, want)
    assert np.array_equal(vel, sc["vel"])                                   # the straight law never touches the velocity
    assert np.array_equal(pos[0], sc["pos"][0])                             # the ground is not a dynamic_obstacle


def test_oracle_circular_law_keeps_height_speed_and_spirals_out_by_the_euler_step(ob):
    sc = dynamic_scene(robots.xarm6(), 5, 2)
    o = ob.Oracle(robots.xarm6())
    o.env_set(params_for("circular"), **sc)
    dt = 0.004
    r0 = np.hypot(sc["pos"][1:, 0], sc["pos"][1:, 1]).astype(np.float64)
    for k in range(1, 40):
        o.env_step(dt, 1)
        pos, vel = o.env_get()
        assert np.array_equal(pos[1:, 2], sc["pos"][1:, 2])
        assert np.allclose(np.linalg.norm(vel[1:], axis=1), sc["max_vel"][1:], rtol=1e-6)
        # explicit Euler along the tangent: r_k^2 = r_0^2 + k (v dt)^2
        want = np.sqrt(r0 ** 2 + k * (sc["max_vel"][1:].astype(np.float64) * dt) ** 2)
        assert np.allclose(np.hypot(pos[1:, 0], pos[1:, 1]), want, atol=2e-6)
        # velocity is tangential (perpendicular to the radius before the step)
    assert np.all(np.abs(np.einsum("ij,ij->i", vel[1:, :2], (pos[1:, :2] - vel[1:, :2] * np.float32(dt)))) < 1e-5)


def test_oracle_two_tunnels_law_oscillates_along_the_axis_chosen_by_the_sign_of_y(ob):
    sc = dynamic_scene(robots.xarm6(), 6, 3, speed=0.5)
    o = ob.Oracle(robots.xarm6())
    o.env_set(params_for("two_tunnels"), **sc)
    p0 = sc["pos"].copy()
    lo, hi = p0.copy(), p0.copy()
    for _ in range(400):
        o.env_step(0.01, 1)
        pos, vel = o.env_get()
        lo, hi = np.minimum(lo, pos), np.maximum(hi, pos)
    along_x = p0[1:, 1] > 0
    span = hi[1:] - lo[1:]
    step = 0.5 * 0.01
    checked = 0
    for j in range(6):
        a = 0 if along_x[j] else 1                                           # y > 0: along x, otherwise along y (Environment.cpp:276-279)
        if not along_x[j] and p0[j + 1, 1] > -0.45:
            continue                                                         # would cross y = 0 and change its axis on the way
        checked += 1
        # first leg 0.2 forwards (+ one step of overshoot), then legs of 0.4 backwards / forwards
        assert 0.4 <= span[j, a] <= 0.4 + 3 * step
        assert hi[j + 1, a] - p0[j + 1, a] <= 0.2 + 2 * step
        assert np.all(span[j, [c for c in range(3) if c != a]] == 0)
    assert checked >= 3


def test_oracle_light_law_preserves_speed_and_keeps_objects_inside_the_workspace_shell(ob):
    sc = dynamic_scene(robots.xarm6(), 12, 4, speed=0.9)
    o = ob.Oracle(robots.xarm6())
    ep = params_for("light_directions")
    o.env_set(ep, **sc)
    speed0 = np.linalg.norm(sc["vel"][1:], axis=1)
    dt = 0.01
    c = np.asarray(WS_CENTER, np.float32)
    bounced = False
    for _ in range(600):
        o.env_step(dt, 1)
        pos, vel = o.env_get()
        assert np.allclose(np.linalg.norm(vel[1:], axis=1), speed0, rtol=2e-3)        # a mirror reflection keeps the speed
        bounced |= not np.allclose(vel[1:], sc["vel"][1:])
        rc = np.linalg.norm(pos[1:] - c, axis=1)
        assert np.all(rc <= ep.ws_radius + 1e-3) and np.all(pos[1:, 2] >= -1e-3)         # reflected at the sphere and at z = 0
        tol = np.maximum(speed0 / ep.robot_max_vel, ep.base_radius)
        assert np.all(rc >= tol - 2 * 0.9 * dt)                                           # and at the robot-base exclusion zone
    assert bounced


def test_oracle_motion_validity_is_update_then_is_valid_point_by_point(ob):
    """MotionValidity::check composed by hand from env_step + isValid (an independent path through the oracle's own primitives)."""
    rb = robots.xarm6()
    sc = dynamic_scene(rb, 5, 5)
    K, dt = 8, 0.006
    q1, q2 = random_motions(rb, 40, 0x77)
    o = ob.Oracle(rb)
    o.env_set(params_for("light_directions"), **sc)
    valid, first = o.motion_validity(q1, q2, K, dt, commit_edge=-1)
    p_before, _ = o.env_get()
    assert np.array_equal(p_before, sc["pos"])                                          # what-if query: nothing moved
    o2 = ob.Oracle(rb)
    o2.env_set(params_for("light_directions"), **sc)
    want_first = np.zeros(40, np.int32)
    dist = np.sqrt(((q2.astype(np.float64) - q1) ** 2).sum(axis=1))
    for k in range(1, K + 1):
        o2.env_step(dt, 1)
        f = (dist * k / K) / np.maximum(dist, 1e-30)
        qk = (q1 + (q2 - q1) * f[:, None].astype(np.float32)).astype(np.float32)
        ok = o2.is_valid(qk).astype(bool)
        want_first[(want_first == 0) & ~ok] = k
    # the interpolation factor above is computed in double: compare away from verdict flips caused by the last ulp of q
    assert (first == want_first).mean() >= 0.95
    assert np.array_equal(valid, first == 0)
    assert 0 < valid.sum() < 40                                                         # the case set exercises both outcomes
    # committing motion e leaves the environment after first_invalid[e] (or K) updates
    e = int(np.flatnonzero(~valid)[0])
    o.motion_validity(q1, q2, K, dt, commit_edge=e)
    o3 = ob.Oracle(rb)
    o3.env_set(params_for("light_directions"), **sc)
    o3.env_step(dt, int(first[e]))
    assert np.array_equal(o.env_get()[0], o3.env_get()[0])


# ---------------------------------------------------------------------------------------------------------------- GPU parity
def _pair(rb, ob_mod, law, sc, precision=None):
    from rpmplv2_b200.engine import Engine
    prm = robots.default_params() if precision is None else robots.default_params(precision=precision)
    eng, orc = Engine(rb, prm), ob_mod.Oracle(rb, prm)
    eng.env_set(params_for(law), **sc)
    orc.env_set(params_for(law), **sc)
    return eng, orc


@pytest.mark.gpu
@pytest.mark.parametrize("law", LAWS)
def test_gpu_motion_laws_match_the_oracle(ob, law):
    rb = robots.xarm6()
    sc = dynamic_scene(rb, 50, 11)
    eng, orc = _pair(rb, ob, law, sc)
    q = synthetic.configs(rb, 2048, seed=0xE1)
    for steps, dt in [(1, 0.01), (7, 0.004), (120, 0.01), (33, 0.02)]:
        eng.env_step(dt, steps)
        orc.env_step(dt, steps)
        (pe, ve), (po, vo) = eng.env_get(), orc.env_get()
        if law == "circular":
            assert np.allclose(pe, po, atol=1e-5) and np.allclose(ve, vo, atol=1e-5)
        else:
            assert np.array_equal(pe, po), np.abs(pe - po).max()
            assert np.array_equal(ve, vo)
    if law != "circular":
        # the resident obstacle table (both copies: caller order and re-ranked scan order) followed the objects
        d_e, p_e = eng.distance(q)[:2]
        d_o, p_o = orc.distance(q)[:2]
        assert np.array_equal(d_e, d_o) and np.array_equal(p_e, p_o)
        assert np.array_equal(eng.is_valid(q), orc.is_valid(q))


@pytest.mark.gpu
@pytest.mark.parametrize("robot,M_dyn,law", [("xarm6", 5, "light_directions"), ("xarm6", 50, "light_directions"), ("xarm6", 5, "straight"),
                                             ("spatial22", 5, "light_directions"), ("planar10", 4, "two_tunnels")])
def test_gpu_motion_validity_matches_the_oracle(ob, robot, M_dyn, law):
    rb = {"xarm6": robots.xarm6, "spatial22": lambda: robots.spatial(22), "planar10": lambda: robots.planar(10)}[robot]()
    reach = 1.0 if robot == "xarm6" else float(np.abs(rb.xyz).sum())
    sc = dynamic_scene(rb, M_dyn, 21, r_lo=0.45 * reach, r_hi=0.95 * reach, size=0.3 * reach)
    if rb.ground_included == 0:                                             # planar chains live in the z = 0 plane, without a ground box
        sc = {k: v[1:].copy() for k, v in sc.items()}
        sc["pos"][:, 2] = 0.0
    eng, orc = _pair(rb, ob, law, sc)
    K, dt = 10, 0.005
    q1, q2 = random_motions(rb, 512, 0x99)
    v_e, f_e = eng.motion_validity(q1, q2, K, dt)
    v_o, f_o = orc.motion_validity(q1, q2, K, dt)
    assert np.array_equal(v_e, v_o) and np.array_equal(f_e, f_o)
    assert 0 < v_o.sum() < len(v_o)
    assert np.array_equal(eng.env_get()[0], sc["pos"])                                  # what-if: the environment did not move
    # a DRGBT-like run: the committed motion decides how far the environment advances, iteration after iteration
    for it in range(6):
        e = it * 37 % 512
        v_e, f_e = eng.motion_validity(q1, q2, K, dt, commit_edge=e)
        v_o, f_o = orc.motion_validity(q1, q2, K, dt, commit_edge=e)
        assert np.array_equal(f_e, f_o), it
        assert np.array_equal(eng.env_get()[0], orc.env_get()[0]) and np.array_equal(eng.env_get()[1], orc.env_get()[1])
    q = synthetic.configs(rb, 1024, seed=0xE2)
    assert np.array_equal(eng.distance(q)[0], orc.distance(q)[0])


@pytest.mark.gpu
def test_gpu_env_argument_errors(ob):
    from rpmplv2_b200.engine import Engine, RvsError
    rb = robots.xarm6()
    sc = dynamic_scene(rb, 3, 31)
    eng = Engine(rb)
    with pytest.raises(RvsError):
        eng.env_step(0.01, 1)                                               # no environment yet
    with pytest.raises(RvsError):
        eng.env_set_random_directions(seed=3)                               # no environment yet
    eng.env_set(params_for("straight"), **sc)
    with pytest.raises(RvsError):
        eng.update_obstacle_aabbs(sc["pos"], sc["pos"])                     # the table is driven by the environment now
    with pytest.raises(RvsError):
        eng.motion_validity(np.zeros((2, 6), np.float32), np.zeros((2, 6), np.float32), 0, 0.01)
    v, f = eng.motion_validity(np.zeros((0, 6), np.float32), np.zeros((0, 6), np.float32), 5, 0.01)
    assert len(v) == 0


@pytest.mark.gpu
def test_gpu_env_edge_cases(ob):
    """Empty environment, a single update / validity point, objects at rest, and a handle switching between a plain table and an
    environment."""
    from rpmplv2_b200.engine import Engine
    rb = robots.xarm6()
    q1, q2 = random_motions(rb, 17, 0x31)
    eng = Engine(rb)
    empty = dict(dim=np.zeros((0, 3), np.float32), pos=np.zeros((0, 3), np.float32))
    eng.env_set(params_for("straight"), **empty)
    v, f = eng.motion_validity(q1, q2, 3, 0.01, commit_edge=0)
    assert v.all() and (f == 0).all()                                        # nothing to collide with
    eng.env_step(0.01, 5)
    assert eng.env_get()[0].shape == (0, 3)
    sc = dynamic_scene(rb, 4, 51)
    sc["vel"][:] = 0
    sc["max_vel"][:] = 0
    eng.env_set(params_for("light_directions"), **sc)
    orc = ob.Oracle(rb)
    orc.env_set(params_for("light_directions"), **sc)
    for K in (1, 2, 25):
        v_e, f_e = eng.motion_validity(q1, q2, K, 0.02, commit_edge=3)
        v_o, f_o = orc.motion_validity(q1, q2, K, 0.02, commit_edge=3)
        assert np.array_equal(f_e, f_o)
        assert np.array_equal(eng.env_get()[0], sc["pos"])                   # objects at rest stay where they are
    # with resting objects the K-th validity point is q2 itself: MotionValidity(K = 1) == isValid(q2)
    v1, _ = eng.motion_validity(q1, q2, 1, 0.02)
    assert np.array_equal(v1, eng.is_valid(q2))
    # back to a plain obstacle table: the environment is switched off
    half = np.float32(0.5) * sc["dim"]
    eng.set_obstacles(sc["pos"] - half, sc["pos"] + half, sc["is_ground"], sc["min_dist_tol"])
    from rpmplv2_b200.engine import RvsError
    with pytest.raises(RvsError):
        eng.env_step(0.01, 1)
    eng.update_obstacle_aabbs(sc["pos"] - half, sc["pos"] + half)            # allowed again
    eng.close()


# ------------------------------------------------------------------- random_directions (Environment.cpp:289-333): libc rand() replayed
def test_oracle_libc_rand_replay_is_the_c_librarys_rand(ob):
    """oracle/rvs_oracle_env.hpp LibcRand against the C library this interpreter runs on: srand(seed), then the same outputs."""
    import ctypes
    libc = ctypes.CDLL(None)
    lib = ob.lib()
    lib.orc_libc_rand.restype = ctypes.c_uint32
    for seed in (0, 1, 7, 1005, 2 ** 31 + 5, 2 ** 32 - 1):
        st = np.zeros(36, np.uint32)
        lib.orc_libc_srand(ctypes.c_uint32(seed), st.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        libc.srand(ctypes.c_uint(seed))
        assert [lib.orc_libc_rand(st.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) for _ in range(3000)] == [libc.rand() for _ in range(3000)]
        assert st[35] == 3000


def _random_oracle(ob, case):
    prm, sc, acc, gi, dt, steps, seed = case
    rb = robots.xarm6()
    o = ob.Oracle(rb)
    o.env_set(env_params("random_directions", ws_center=prm[:3], ws_radius=prm[3], base_radius=prm[4], robot_max_vel=prm[5]), **sc)
    o.env_set_random_directions(acc, gi, seed)
    return o


def test_oracle_random_directions_matches_the_reference_environment_golden(ob):
    """Environment::moveRandomDirections compiled unmodified, run after srand(seed) (tests/golden/make_env_golden.py: envref_update_random):
    positions, velocities and accelerations of every object bit for bit along the whole run, and the stream ends where libc's does."""
    import ctypes
    mod = _golden_cases()
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz")))
    for k, case in enumerate(mod.random_cases()):
        o = _random_oracle(ob, case)
        dt, steps = float(case[4]), case[5]
        keep = list(gold[f"r{k}_keep"])
        for step in range(steps):
            o.env_step(dt, 1)
            if step in keep:
                i = keep.index(step)
                (pos, vel), (acc, _) = o.env_get(), o.env_get_random_directions()
                assert np.array_equal(pos, gold[f"r{k}_pos"][i]) and np.array_equal(vel, gold[f"r{k}_vel"][i]) and np.array_equal(acc, gold[f"r{k}_acc"][i]), (k, step)
        st = o._env["rand"]
        ob.lib().orc_libc_rand.restype = ctypes.c_uint32
        assert st[35] > 0 and ob.lib().orc_libc_rand(st.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) == gold[f"r{k}_next"][0]
        # invariants of the law: speed never above max_vel, every object inside the workspace shell and above the ground
        pos, vel = o.env_get()
        dyn = case[1]["is_dynamic"].astype(bool)
        assert (np.linalg.norm(vel[dyn], axis=1) <= case[1]["max_vel"][dyn] * (1 + 1e-6)).all()
        assert (np.linalg.norm(pos[dyn] - case[0][:3], axis=1) <= case[0][3] * (1 + 1e-6)).all()


@pytest.mark.gpu
def test_gpu_random_directions_matches_the_oracle_and_the_reference_golden(ob):
    """rvs_env_step under RVS_MOTION_RANDOM_DIRECTIONS: the device's replay of the libc stream moves every object exactly as the
    reference's loop does (golden vectors of the compiled Environment.cpp), one update per call and many updates per call."""
    from rpmplv2_b200.engine import Engine
    mod = _golden_cases()
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz")))
    for k, case in enumerate(mod.random_cases()):
        prm, sc, acc, gi, dt, steps, seed = case
        ep = env_params("random_directions", ws_center=prm[:3], ws_radius=prm[3], base_radius=prm[4], robot_max_vel=prm[5])
        for chunk in (1, 10):
            eng = Engine(robots.xarm6())
            eng.env_set(ep, **sc)
            eng.env_set_random_directions(acc, gi, seed)
            keep = list(gold[f"r{k}_keep"])
            for step in range(chunk - 1, steps, chunk):
                eng.env_step(float(dt), chunk)
                if step in keep:
                    i = keep.index(step)
                    (pos, vel), (a_dev, draws) = eng.env_get(), eng.env_get_random_directions()
                    assert np.array_equal(pos, gold[f"r{k}_pos"][i]) and np.array_equal(vel, gold[f"r{k}_vel"][i]) and np.array_equal(a_dev, gold[f"r{k}_acc"][i]), (k, chunk, step)
            o = _random_oracle(ob, case)
            o.env_step(float(dt), (steps // chunk) * chunk)
            assert eng.env_get_random_directions()[1] == o.env_get_random_directions()[1] > 0           # same number of rand() calls consumed
            q = synthetic.configs(robots.xarm6(), 512, seed=0xE5)
            assert np.array_equal(eng.distance(q)[0], o.distance(q)[0])                                  # the resident table followed
            eng.close()


@pytest.mark.gpu
def test_gpu_motion_validity_under_random_directions(ob):
    """MotionValidity::check with obstacles moving by the random law: what-if queries leave objects AND stream untouched; a committed
    motion leaves both as after first_invalid (or K) updates."""
    from rpmplv2_b200.engine import Engine
    rb = robots.xarm6()
    sc = dynamic_scene(rb, 30, 77, speed=0.5, size=0.25)
    ep = env_params("random_directions", ws_center=WS_CENTER, ws_radius=1.0, base_radius=0.15, robot_max_vel=3.0)
    acc = np.zeros((31, 3), np.float32)
    acc[2::2] = [0.3, -0.2, 0.1]
    eng, o = Engine(rb), ob.Oracle(rb)
    for x in (eng, o):
        x.env_set(ep, **sc)
        x.env_set_random_directions(acc, 2, 99)
    q1, q2 = random_motions(rb, 256, 0x77)
    K, dt = 10, 0.01
    v_e, f_e = eng.motion_validity(q1, q2, K, dt)
    v_o, f_o = o.motion_validity(q1, q2, K, dt)
    assert np.array_equal(f_e, f_o) and 0 < v_o.sum() < len(v_o)
    assert np.array_equal(eng.env_get()[0], sc["pos"]) and eng.env_get_random_directions()[1] == 0
    for it in range(40):
        e = it * 37 % 256
        v_e, f_e = eng.motion_validity(q1, q2, K, dt, commit_edge=e)
        v_o, f_o = o.motion_validity(q1, q2, K, dt, commit_edge=e)
        assert np.array_equal(f_e, f_o), it
        assert np.array_equal(eng.env_get()[0], o.env_get()[0]) and np.array_equal(eng.env_get()[1], o.env_get()[1])
        assert eng.env_get_random_directions()[1] == o.env_get_random_directions()[1]
    assert o.env_get_random_directions()[1] > 0
    eng.close()


# ------------------------------------------------------------------- pinned by the reference's own Environment.cpp (compiled with shims)
def _golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_env_golden", os.path.join(ROOT, "tests", "golden", "make_env_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_motion_laws_match_the_reference_environment_golden(ob):
    """oracle/rvs_oracle_env.hpp against vectors produced by the reference's Environment.cpp / Object.cpp / Box.cpp compiled against
    oracle/eigen_shim + oracle/env_shim (tests/golden/make_env_golden.py): positions, velocities, two_tunnels bookkeeping and the AABBs
    after 1 .. 250 updates, bit for bit (circular: the same libm on both sides here)."""
    mod = _golden_cases()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz"))
    for k, (law, prm, sc, dt, steps) in enumerate(mod.cases()):
        o = ob.Oracle(robots.xarm6())
        o.env_set(env_params(law, ws_center=tuple(prm[:3]), ws_radius=float(prm[3]), base_radius=float(prm[4]), robot_max_vel=float(prm[5])), **sc)
        o.env_step(float(dt), steps)
        pos, vel = o.env_get()
        assert np.array_equal(pos, gold[f"c{k}_pos"]), (law, k, np.abs(pos - gold[f"c{k}_pos"]).max())
        assert np.array_equal(vel, gold[f"c{k}_vel"]), (law, k)
        assert np.array_equal(o._env["sign"], gold[f"c{k}_sign"]) and np.array_equal(o._env["path_len"], gold[f"c{k}_plen"])
        half = np.float32(0.5) * sc["dim"]
        assert np.array_equal(pos + (-half), gold[f"c{k}_mn"]) and np.array_equal(pos + half, gold[f"c{k}_mx"])


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libenv_ref.so")), reason="oracle/_ref is built where /root/reference exists")
def test_oracle_motion_laws_match_the_compiled_reference_step_by_step(ob):
    """The same comparison against the compiled reference itself, continuing from state to state (random restarts exercise the
    reflection branches and the two_tunnels reversal with arbitrary bookkeeping)."""
    import ctypes
    mod = _golden_cases()
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libenv_ref.so"))
    rb = robots.xarm6()
    prm = np.array(list(WS_CENTER) + [0.9, 0.25, 1.2], np.float32)
    for law in mod.LAWS:
        sc = dynamic_scene(rb, 14, 77, speed=1.1)
        o = ob.Oracle(rb)
        o.env_set(env_params(law, ws_center=WS_CENTER, ws_radius=0.9, base_radius=0.25, robot_max_vel=1.2), **sc)
        state = None
        for dt, steps in ((0.013, 5), (0.02, 61), (0.007, 200), (0.031, 17)):
            state = mod.ref_update(lib, prm, law, sc, dt, steps, state)
            o.env_step(dt, steps)
            pos, vel = o.env_get()
            assert np.array_equal(pos, state[0]) and np.array_equal(vel, state[1]), law
            assert np.array_equal(o._env["sign"], state[2]) and np.array_equal(o._env["path_len"], state[3]), law


# ---------------------------------------------- pinned by the reference's own MotionValidity.cpp on its own Environment / RealVectorSpace
def _motion_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_motion_golden", os.path.join(ROOT, "tests", "golden", "make_motion_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run_motions(space, mod, law, sc, q1, q2, K, mit, points=None):
    """T committed checks on one environment through a space with the engine's env_set / motion_validity[_points] / env_get methods."""
    space.env_set(env_params(law, **mod.ENV), **sc)
    dt = mod.delta_time(K, mit)
    T = len(q1)
    out = dict(valid=np.zeros(T, np.uint8), passed=np.zeros(T, np.int32), pos=[], vel=[])
    for t in range(T):
        if points is None:
            v, first = space.motion_validity(q1[t:t + 1], q2[t:t + 1], K, dt, commit_edge=0)
        else:
            v, first = space.motion_validity_points(points[t:t + 1], dt, commit_edge=0)
        out["valid"][t] = 1 if v[0] else 0
        out["passed"][t] = K if v[0] else first[0] - 1                   # `path` receives the points that passed before the failing one
        pos, vel = space.env_get()
        out["pos"].append(pos)
        out["vel"].append(vel)
    out["pos"], out["vel"] = np.stack(out["pos"]), np.stack(out["vel"])
    return out


def _compare_motion_run(got, want, tag, pos_tol=0.0):
    assert np.array_equal(got["valid"], want["valid"]) and np.array_equal(got["passed"], want["passed"]), tag
    for k in ("pos", "vel"):
        assert np.array_equal(got[k], want[k]) if pos_tol == 0 else np.abs(got[k] - want[k]).max() <= pos_tol, (tag, k)


def test_oracle_motion_check_matches_the_reference_motionvalidity_golden(ob):
    """oracle (env_update + motion_point + is_valid, commit semantics) against runs of the reference's MotionValidity::check compiled
    unmodified (tests/golden/make_motion_golden.py): verdicts, the number of validity points that passed, and the obstacles'
    positions / velocities after every check — bit for bit."""
    mod = _motion_golden()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "motion_golden.npz"))
    n_valid = n_all = 0
    for k, (robot, M_dyn, law, T, K, mit) in enumerate(mod.CASES):
        rb, sc, q1, q2 = mod.case_inputs(robot, M_dyn, law, T)
        got = _run_motions(ob.Oracle(rb), mod, law, sc, q1, q2, K, mit)
        _compare_motion_run(got, {n: gold[f"c{k}_{n}"] for n in ("valid", "passed", "pos", "vel")}, (robot, law, K))
        n_valid, n_all = n_valid + int(got["valid"].sum()), n_all + T
        # the other overload: explicit trajectory points (MotionValidity.cpp:60-82)
        got = _run_motions(ob.Oracle(rb), mod, law, sc, q1, q2, K, mit, points=mod.trajectory_points(q1, q2, K))
        _compare_motion_run(got, {n: gold[f"c{k}_pts_{n}"] for n in ("valid", "passed", "pos", "vel")}, (robot, law, K, "points"))
    assert 0.2 < n_valid / n_all < 0.8


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")), reason="oracle/_ref is built where /root/reference exists")
def test_oracle_motion_check_matches_the_compiled_reference_on_fresh_runs(ob):
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so"))
    if not hasattr(lib, "rvsref_motion_run"):
        pytest.skip("oracle/_ref/librvs_ref.so predates the MotionValidity harness")
    mod = _motion_golden()
    for robot, M_dyn, law, T, K, mit, seed in [("xarm6", 20, "light_directions", 80, 5, 0.03, 3), ("xarm6", 6, "two_tunnels", 80, 12, 0.1, 4),
                                               ("spatial22", 9, "circular", 50, 10, 0.05, 5), ("planar10", 6, "straight", 50, 3, 0.04, 6)]:
        rb, sc, q1, q2 = mod.case_inputs(robot, M_dyn, law, T, seed=seed)
        want = mod.run_reference(lib, rb, sc, law, q1, q2, K, mit)
        _compare_motion_run(_run_motions(ob.Oracle(rb), mod, law, sc, q1, q2, K, mit), want, (robot, law, K))
        if hasattr(lib, "rvsref_motion_points_run"):
            pts = mod.trajectory_points(q1, q2, K)
            want = mod.run_reference(lib, rb, sc, law, q1, q2, K, mit, points=pts)
            _compare_motion_run(_run_motions(ob.Oracle(rb), mod, law, sc, q1, q2, K, mit, points=pts), want, (robot, law, K, "points"))


@pytest.mark.gpu
def test_gpu_motion_check_matches_the_reference_motionvalidity_golden():
    """rvs_env_set + rvs_motion_validity (commit) + rvs_env_get straight against the compiled reference's runs.  The circular law goes
    through atan2f / sinf / cosf (last-ulp differences between CUDA and glibc): positions within 1e-5 m there, bit for bit elsewhere."""
    from rpmplv2_b200.engine import Engine
    mod = _motion_golden()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "motion_golden.npz"))
    for k, (robot, M_dyn, law, T, K, mit) in enumerate(mod.CASES):
        rb, sc, q1, q2 = mod.case_inputs(robot, M_dyn, law, T)
        eng = Engine(rb)
        got = _run_motions(eng, mod, law, sc, q1, q2, K, mit)
        _compare_motion_run(got, {n: gold[f"c{k}_{n}"] for n in ("valid", "passed", "pos", "vel")}, (robot, law, K), pos_tol=1e-5 if law == "circular" else 0.0)
        got = _run_motions(eng, mod, law, sc, q1, q2, K, mit, points=mod.trajectory_points(q1, q2, K))
        _compare_motion_run(got, {n: gold[f"c{k}_pts_{n}"] for n in ("valid", "passed", "pos", "vel")}, (robot, law, K, "points"),
                            pos_tol=1e-5 if law == "circular" else 0.0)
        # what-if form: many candidate trajectories against one timeline, nothing moves
        eng.env_set(env_params(law, **mod.ENV), **sc)
        pts = mod.trajectory_points(q1, q2, K)
        v_all, f_all = eng.motion_validity_points(pts, mod.delta_time(K, mit))
        v_one = np.array([eng.motion_validity_points(pts[t:t + 1], mod.delta_time(K, mit))[0][0] for t in range(0, len(pts), 7)])
        assert np.array_equal(v_all[::7], v_one) and np.array_equal(eng.env_get()[0], sc["pos"])
        eng.close()
