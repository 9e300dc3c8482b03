"""The CPU oracle against (i) golden vectors produced by the reference's own CollisionAndDistance.cpp, (ii) analytic
known answers, (iii) an independent brute-force distance, (iv) the reference-authored invariant of
apps/test_real_vector_space.cpp:46-57."""
import numpy as np
import pytest

from conftest import same
from rpmplv2_b200 import robots, synthetic


@pytest.mark.parametrize("tag,thr", [("t4", 1e-4), ("t6", 1e-6)])
def test_capsule_box_distance_matches_reference_golden(golden, ob, tag, thr):
    g = golden[f"{tag}_dist"]
    o = ob.capsule_box_distance(golden[f"{tag}_AB"], golden[f"{tag}_r"], golden[f"{tag}_obs"], thr=thr)
    assert same(o[:, 0], g[:, 0]).all()                      # distances: bit exact
    assert (o[:, 1] == g[:, 1]).all()                        # nullptr-ness of the nearest points
    live = (g[:, 1] == 1) & (g[:, 0] > 0)
    assert same(o[live, 2:], g[live, 2:]).all()              # nearest points: bit exact where defined


@pytest.mark.parametrize("tag,thr", [("t4", 1e-4), ("t6", 1e-6)])
def test_collision_predicates_match_reference_golden(golden, ob, tag, thr):
    AB, r, obs = golden[f"{tag}_AB"], golden[f"{tag}_r"], golden[f"{tag}_obs"]
    assert (ob.capsule_box_collision(AB, r, obs, thr=thr) == golden[f"{tag}_coll"]).all()
    assert (ob.capsule_rect_collision(AB, r, obs, golden[f"{tag}_face"], thr=thr) == golden[f"{tag}_rect"]).all()


def test_segment_primitives_match_reference_golden(golden, ob):
    assert same(ob.seg_seg(golden["ss_ABCD"]), golden["ss_out"]).all()
    assert same(ob.seg_point(golden["sp_ABC"]), golden["sp_out"]).all()


def test_hand_checkable_capsule_box_kats(ob):
    """Unit box, r = 0.1."""
    AB = np.array([[-1, .5, .5, -.5, .5, .5], [-1, -1, .5, -1, -1, .6], [-1, -1, -1, -2, -2, -2], [.5, .5, .5, 2, 2, 2]], np.float32)
    out = ob.capsule_box_distance(AB, 0.1, np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (4, 1)))
    np.testing.assert_allclose(out[:3, 0], [0.4, np.sqrt(2) - 0.1, np.sqrt(3) - 0.1], rtol=0, atol=2e-7)
    assert out[3, 0] == 0.0 and out[3, 1] == 0.0             # endpoint inside -> exactly 0, null points (Q11)
    np.testing.assert_allclose(out[0, 2:], [-.5, .5, .5, 0, .5, .5], atol=1e-7)


def test_brute_force_distance_bound(ob):
    """Oracle distance vs dense sampling of the segment against the closed-form point-AABB distance."""
    rng = np.random.default_rng(5)
    K = 400
    c = rng.uniform(-1, 1, (K, 3)); h = rng.uniform(0.05, 0.6, (K, 3)) / 2
    obs = np.concatenate([c - h, c + h], 1).astype(np.float32)
    A = rng.uniform(-1.5, 1.5, (K, 3)); B = A + rng.normal(size=(K, 3)) * 0.4
    AB = np.concatenate([A, B], 1).astype(np.float32)
    d = ob.capsule_box_distance(AB, 0.0, obs)[:, 0].astype(np.float64)
    t = np.linspace(0, 1, 4001)[None, :, None]
    P = AB[:, None, :3] + t * (AB[:, None, 3:] - AB[:, None, :3])
    gap = np.maximum(np.maximum(obs[:, None, :3] - P, P - obs[:, None, 3:]), 0.0)
    brute = np.sqrt((gap ** 2).sum(-1)).min(1)
    free = brute > 1e-3
    assert free.sum() > 100
    # sampling can only over-estimate by <= |AB| / 8000; the oracle must never be larger than the sampled distance
    seglen = np.linalg.norm(AB[:, 3:] - AB[:, :3], axis=1)
    assert (d[free] <= brute[free] + 1e-5).all()
    assert (d[free] >= brute[free] - seglen[free] / 8000 - 1e-5).all()
    assert (d[~free & (brute == 0)] <= 0).all()


def test_forward_kinematics_known_answers(ob):
    x = ob.Oracle(robots.xarm6())
    sk = x.skeleton(np.zeros((1, 6), np.float32))[0]
    expect = np.array([[0, 0, 0], [0, 0, .267], [.0535, 0, .5515], [.131, 0, .459], [.131, 0, .209], [.207, 0, .209], [.207, 0, -.007]])
    np.testing.assert_allclose(sk, expect, atol=5e-6)        # 1.5708 != pi/2 -> ~4e-6
    phi = 0.7
    sk2 = x.skeleton(np.array([[phi, 0, 0, 0, 0, 0]], np.float32))[0]
    rot = np.array([[np.cos(phi), -np.sin(phi), 0], [np.sin(phi), np.cos(phi), 0], [0, 0, 1]])
    np.testing.assert_allclose(sk2, sk @ rot.T, atol=1e-6)   # joint 1 rotates everything about z
    p = ob.Oracle(robots.planar(2))
    q1, q2 = 0.3, -1.1
    sp = p.skeleton(np.array([[q1, q2]], np.float32))[0]
    np.testing.assert_allclose(sp, [[0, 0, 0], [np.cos(q1), np.sin(q1), 0], [np.cos(q1) + np.cos(q1 + q2), np.sin(q1) + np.sin(q1 + q2), 0]], atol=2e-7)
    s = ob.Oracle(robots.spatial(10))
    ss = s.skeleton(np.zeros((1, 10), np.float32))[0]
    np.testing.assert_allclose(ss[:, 2], np.arange(11) * 0.1, atol=1e-7)


def test_fk_matches_scipy(ob):
    from scipy.spatial.transform import Rotation as Rot
    rb = robots.xarm6()
    q = synthetic.configs(rb, 50)
    sk = ob.Oracle(rb).skeleton(q, f64=True)
    for i in range(len(q)):
        R, p, frames = np.eye(3), np.zeros(3), []
        for k in range(6):
            p = p + R @ rb.xyz[k]
            R = R @ Rot.from_euler("xyz", rb.rpy[k]).as_matrix() @ Rot.from_rotvec(rb.axis[k] * float(q[i, k])).as_matrix()
            frames.append((R.copy(), p.copy()))
        g = rb.gripper_length
        expect = [np.zeros(3), frames[1][1], frames[2][1], frames[3][1] - 0.25 * frames[3][0][:, 2], frames[4][1],
                  frames[4][1] + 0.076 * frames[4][0][:, 0], frames[5][1] + g * frames[5][0][:, 2] - 0.3 * g * frames[5][0][:, 2]]
        np.testing.assert_allclose(sk[i], np.array(expect), atol=2e-8)


def test_underestimation_equals_distance_at_the_generating_configuration(ob):
    """apps/test_real_vector_space.cpp:46-57: |d_c - d_c_under(own planes)| <= 1e-3 for collision-free q."""
    rb = robots.xarm6()
    o = ob.Oracle(rb)
    mn, mx, g, tol = synthetic.sweep_obstacles(12, "D2")
    tol[:] = 1000.0                                           # no min_dist_tol clipping in this check
    o.set_obstacles(mn, mx, g, tol)
    q = synthetic.configs(rb, 300)
    d, prof, near = o.distance(q, robots.DIST_FULL)
    free = d > 0
    assert free.sum() > 20
    idx = np.flatnonzero(free)
    du, pu, st = o.underestimation(q[idx], near[idx], np.arange(len(idx), dtype=np.int32))
    assert (st == 0).all()
    assert np.abs(du - d[idx]).max() <= 1e-3
    assert np.abs(pu - prof[idx]).max() <= 1e-3


def test_distance_early_exit_fill_and_ground_rule(ob):
    rb = robots.xarm6()
    o = ob.Oracle(rb)
    mn, mx, g, tol = synthetic.sweep_obstacles(30, "D2")
    o.set_obstacles(mn, mx, g, tol)
    q = synthetic.configs(rb, 500)
    d, prof, _ = o.distance(q, robots.DIST_PROFILE)
    hit = d == 0
    assert hit.any() and (~hit).any()
    for i in np.flatnonzero(hit)[:50]:
        first = np.flatnonzero(prof[i] <= 0)[0]
        assert (prof[i, first + 1:] == 0).all() and (prof[i, :first] > 0).all()
    assert np.allclose(d[~hit], prof[~hit].min(1))
    # links 0 and 1 never see the ground (ground_included = 2): with only the ground present their profile is MAX_DISTANCE
    o.set_obstacles(mn[:1], mx[:1], g[:1], tol[:1])
    _, p2, _ = o.distance(q[:20], robots.DIST_PROFILE)
    assert (p2[:, :2] == 1000.0).all()


# --------------------------------------------------------------------------------------------------------------------------------
# The LOOPS of the hot path, pinned by the reference's own RealVectorSpace.cpp (compiled unmodified against oracle/eigen_shim +
# oracle/env_shim; tests/golden/make_rvs_golden.py -> tests/golden/rvs_golden.npz): isValid(q), isValid(q1,q2), computeDistance (d_c,
# profile with its first-collision fill, nearest points and their nullptr rule, is_real) and computeDistanceUnderestimation (values,
# the reject rule, untouched states).  Skeletons come from the oracle's FK on both sides (KDL itself cannot be pinned here).
# --------------------------------------------------------------------------------------------------------------------------------
def _rvs_golden_module():
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_rvs_golden", os.path.join(ROOT, "tests", "golden", "make_rvs_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, dict(np.load(os.path.join(ROOT, "tests", "golden", "rvs_golden.npz")))


def _check_space_against(space_factory, name, case, gold, exact=True):
    from rpmplv2_b200 import robots
    rb, mn, mx, g, tol, thr, N = case
    mod, _ = _rvs_golden_module()
    q, q1, q2 = mod.inputs(rb, N)
    sp = space_factory(rb, robots.default_params(equality_threshold=thr))
    sp.set_obstacles(mn, mx, g, tol)
    G = lambda k: gold[f"{name}__{k}"]
    assert np.array_equal(sp.is_valid(q).astype(np.uint8), G("valid")), name
    assert np.array_equal(sp.is_valid_edges(q1, q2).astype(np.uint8), G("edge")), name
    d, p, near = sp.distance(q, robots.DIST_FULL)
    free = G("d") > 0
    assert np.array_equal(G("has_near").astype(bool), free) and G("is_real").all()         # nullptr exactly after a collision
    if exact:
        assert np.array_equal(d, G("d")) and np.array_equal(p, G("prof")) and np.array_equal(near[free], G("near")[free]), name
    else:
        assert ((d == 0) == (G("d") == 0)).all() and np.abs(d - G("d")).max() <= 1e-4 and np.abs(p - G("prof")).max() <= 1e-4
        assert np.abs(near[free] - G("near")[free]).max() <= 1e-4
        assert (d != G("d")).mean() < 5e-3                                                   # bit-exact up to rare FK last-ulp effects
    root = int(G("root"))
    du, pu, st = sp.underestimation(q, np.ascontiguousarray(G("near")[root])[None], np.zeros(N, np.int32))
    assert np.array_equal(st, G("un_status")), name
    acc = G("un_status") == 0
    assert not G("un_is_real")[acc].any()                                                    # accepted: is_real_d_c = false
    if exact:
        assert np.array_equal(du, G("un_d")) and np.array_equal(pu[acc], G("un_prof")[acc]), name
    else:
        assert np.abs(du - G("un_d")).max() <= 1e-4 and np.abs(pu[acc] - G("un_prof")[acc]).max() <= 1e-4
    assert 0 < free.sum() < N and 0 < G("un_status").sum()


def test_oracle_loops_match_the_reference_realvectorspace_golden(ob):
    mod, gold = _rvs_golden_module()
    for name, case in mod.cases().items():
        _check_space_against(lambda rb, prm: ob.Oracle(rb, prm), name, case, gold, exact=True)


def test_oracle_loops_match_the_compiled_reference_on_fresh_inputs(ob):
    """The same comparison against oracle/_ref/librvs_ref.so itself on scenes that are not in the golden file (only where the
    reference tree was available to build it)."""
    import os
    from conftest import ROOT
    import scenes
    so = os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref is built where /root/reference exists")
    import ctypes
    mod, _ = _rvs_golden_module()
    lib = ctypes.CDLL(so)
    from rpmplv2_b200 import robots
    fresh = {"xarm6_m120": scenes.xarm6_scene(120, "D2", 0.3) + (1e-4, 150), "planar10": scenes.chain_scene(robots.planar(10), 8) + (1e-4, 150),
             "spatial14": scenes.chain_scene(robots.spatial(14), 20) + (1e-5, 120)}
    for name, case in fresh.items():
        r = mod.run_reference(lib, *case)
        gold = {f"{name}__{k}": v for k, v in r.items()}
        _check_space_against(lambda rb, prm: ob.Oracle(rb, prm), name, case, gold, exact=True)


@pytest.mark.gpu
def test_gpu_engine_matches_the_reference_realvectorspace_golden():
    """The CUDA engine (REF32) straight against the vectors of the compiled reference — not via the oracle."""
    from rpmplv2_b200 import engine
    mod, gold = _rvs_golden_module()
    for name, case in mod.cases().items():
        _check_space_against(lambda rb, prm: engine.Engine(rb, prm), name, case, gold, exact=False)


@pytest.mark.gpu
def test_gpu_full_mode_split_matches_the_reference_realvectorspace_golden():
    """The golden configurations of the compiled reference, repeated until the batch is past the small-batch cross-over, so that FULL goes
    through the three-launch path (csrc/rvs_distance_full.cuh): every replica reproduces the reference's d_c, profile and nearest points
    (same tolerances as the test above) and is bit-identical to the small-batch kernel's answer for the same configuration."""
    from rpmplv2_b200 import engine, robots
    mod, gold = _rvs_golden_module()
    checked = 0
    for name, case in mod.cases().items():
        rb, mn, mx, g, tol, thr, N = case
        if len(mn) < 16:
            continue                                                                         # short tables keep the one-kernel path
        q, _, _ = mod.inputs(rb, N)
        reps = -(-8300 // N)
        eng = engine.Engine(rb, robots.default_params(equality_threshold=thr))
        eng.set_obstacles(mn, mx, g, tol)
        d_s, p_s, n_s = eng.distance(q, robots.DIST_FULL)                                    # one CTA per configuration
        k0 = eng.launch_count()
        d, p, near = eng.distance(np.tile(q, (reps, 1)), robots.DIST_FULL)
        assert eng.launch_count() - k0 >= 3, name
        G = lambda k: gold[f"{name}__{k}"]
        free = G("d") > 0
        for r in range(reps):
            sl = slice(r * N, (r + 1) * N)
            assert np.array_equal(d[sl], d_s) and np.array_equal(p[sl], p_s) and np.array_equal(near[sl][free], n_s[free]), name
        assert ((d[:N] == 0) == (G("d") == 0)).all() and np.abs(d[:N] - G("d")).max() <= 1e-4 and np.abs(p[:N] - G("prof")).max() <= 1e-4
        assert np.abs(near[:N][free] - G("near")[free]).max() <= 1e-4 and (near[:N][free] != G("near")[free]).mean() < 5e-3
        eng.close()
        checked += 1
    assert checked >= 3
