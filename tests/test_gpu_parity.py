"""GPU parity tests: the sm_100a kernels, called through the C ABI (librvs_cuda.so), against the CPU oracle and the
golden vectors of the reference's own CollisionAndDistance.cpp.

Tolerances (BASELINE.json north_star / SURVEY.md §8d):
  REF32  : verdicts identical; distances |d - d_oracle<float>| <= 1e-4 m (expected: bit-exact up to FK last-ulp effects)
  FAST32 : distances <= 1e-4 m; verdicts identical outside the band |clearance| <= 1e-4 m
  F64    : distances |d - d_oracle<double>| <= 1e-9 m
"""
import numpy as np
import pytest

import scenes
from conftest import same
from rpmplv2_b200 import engine, robots, synthetic

pytestmark = pytest.mark.gpu

TOL32, TOL64 = 1e-4, 1e-9


def make(rb, mn, mx, g, tol, prec=robots.PREC_REF32, thr=1e-4):
    eng = engine.Engine(rb, robots.default_params(prec, equality_threshold=thr))
    eng.set_obstacles(mn, mx, g, tol)
    return eng


def make_oracle(ob, rb, mn, mx, g, tol, thr=1e-4):
    o = ob.Oracle(rb, robots.default_params(equality_threshold=thr))
    o.set_obstacles(mn, mx, g, tol)
    return o


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,thr", [("t4", 1e-4), ("t6", 1e-6)])
def test_primitives_match_reference_golden_bit_exact(golden, tag, thr):
    eng = engine.Engine(robots.xarm6(), robots.default_params(robots.PREC_REF32, equality_threshold=thr))
    AB, r, obs, g = golden[f"{tag}_AB"], golden[f"{tag}_r"], golden[f"{tag}_obs"], golden[f"{tag}_dist"]
    out = eng.capsule_box_distance(AB, r, obs)
    assert same(out[:, 0], g[:, 0]).all()
    live = (g[:, 1] == 1) & (g[:, 0] > 0)
    assert same(out[live, 2:], g[live, 2:]).all()
    assert (eng.capsule_box_collision(AB, r, obs) == golden[f"{tag}_coll"]).all()


def test_primitives_fast32_within_tolerance(golden):
    eng = engine.Engine(robots.xarm6(), robots.default_params(robots.PREC_FAST32))
    g = golden["t4_dist"]
    out = eng.capsule_box_distance(golden["t4_AB"], golden["t4_r"], golden["t4_obs"])
    ok = np.isfinite(g[:, 0])
    diff = np.abs(out[ok, 0] - g[ok, 0])
    # the snap (d < EQUALITY_THRESHOLD -> -r) and the exactly-on-a-face cases are discontinuous: allow them in the band
    big = diff > TOL32
    assert big.mean() < 0.01
    coll = eng.capsule_box_collision(golden["t4_AB"], golden["t4_r"], golden["t4_obs"])
    assert (coll != golden["t4_coll"]).mean() < 0.005


@pytest.mark.parametrize("rb", [robots.xarm6(), robots.planar(2), robots.planar(10), robots.spatial(10), robots.spatial(14),
                                robots.spatial(18), robots.spatial(22)], ids=lambda r: r.name)
def test_skeleton_radii_selfcollision(ob, rb):
    q = synthetic.configs(rb, 5000)
    eng, o = engine.Engine(rb), ob.Oracle(rb)
    sk, sk_o = eng.skeleton(q), o.skeleton(q)
    assert np.abs(sk - sk_o).max() <= 2.4e-7          # FK in FP64 on both sides; at most a last-ulp flip after narrowing
    assert (sk != sk_o).mean() < 1e-4
    R, R_o = eng.enclosing_radii(q), o.enclosing_radii(q)
    assert np.abs(R - R_o).max() <= 1e-6
    # capsule pre-filter of xArm6::checkSelfCollision(q): exact unless a capsule distance sits within an FK ulp of zero (the skeletons differ
    # by <= 2.4e-7 m, so a pair whose computeCapsulesDistance is inside +-1e-6 m may flip; everything else must agree bit for bit)
    m_g, m_o = eng.self_collision_capsules(q), o.self_collision_capsules(q)
    assert (m_g != m_o).sum() <= 2
    if rb.kind == robots.ROBOT_XARM6:
        assert (m_o != 0).sum() > 100


@pytest.mark.parametrize("M,dist,r_min,N", [(1, "D2", 0.0, 12000), (10, "D2", 0.0, 12000), (51, "D1", 0.0, 12000), (100, "D2", 0.5, 12000), (600, "D2", 0.9, 12000),
                                             (1300, "D1", 0.8, 12000), (1000, "D2", 0.9, 60000)])
def test_distance_ref32_vs_oracle(ob, M, dist, r_min, N):
    """(1000, D2, 0.9) is EXACTLY the scene bench.py's headline is quoted on (bench.SCENES["headline"]), 60 000 configurations against the oracle."""
    rb, mn, mx, g, tol = scenes.xarm6_scene(M, dist, r_min)
    q = synthetic.configs(rb, N)
    eng, o = make(rb, mn, mx, g, tol), make_oracle(ob, rb, mn, mx, g, tol)
    d_o, p_o, _ = o.distance(q, robots.DIST_PROFILE, threads=8)
    d_min, _, _ = eng.distance(q, robots.DIST_MIN)
    d, p, _ = eng.distance(q, robots.DIST_PROFILE)
    assert (d_min == d).all()
    assert ((d == 0) == (d_o == 0)).all()                              # collision verdict of computeDistance
    assert np.abs(d - d_o).max() <= TOL32 and np.abs(p - p_o).max() <= TOL32
    assert (d != d_o).mean() < 1e-3 and (p != p_o).mean() < 1e-3       # REF32 is bit-exact up to rare FK last-ulp effects
    assert eng.pair_count(reset=True) == 2 * o.pair_count(reset=True)  # two GPU calls vs one oracle call


@pytest.mark.parametrize("prec", [robots.PREC_REF32, robots.PREC_FAST32, robots.PREC_F64])
@pytest.mark.parametrize("case", ["xarm6_1000", "xarm6_2500_two_tiles", "xarm6_dense", "spatial22", "planar2", "tol_and_thr"])
def test_work_queue_kernel_is_bit_identical_to_sequential_kernel(case, prec):
    """The scan -> work-queue -> compacted-phase kernel (rvs_distance_wq.cuh, the default for MIN/PROFILE) against the
    one-thread-per-configuration kernel (rvs_set_culling(0)): d_c, the whole profile (including the early-exit fill after a
    collision) and the reference-order pair count must be identical BIT FOR BIT, whatever order the queues were served in."""
    thr = 1e-4
    if case == "xarm6_1000":
        rb, mn, mx, g, tol = scenes.xarm6_scene(1000, "D2", 0.9)
    elif case == "xarm6_2500_two_tiles":
        rb, mn, mx, g, tol = scenes.xarm6_scene(2500, "D1", 0.6)
    elif case == "xarm6_dense":
        rb, mn, mx, g, tol = scenes.xarm6_scene(300, "D2", 0.0)           # most configurations collide
    elif case == "spatial22":
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(22), 40)
    elif case == "planar2":
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    else:
        rb, mn, mx, g, tol = scenes.xarm6_scene(200, "D2", 0.5)
        tol = np.random.default_rng(5).choice(np.array([0.02, 0.1, 0.4, 1000.0], np.float32), len(tol)).astype(np.float32)
        thr = 0.05                                                         # a large EQUALITY_THRESHOLD: snaps far from contact
    q = synthetic.configs(rb, 40000 if prec != robots.PREC_F64 else 15000)
    eng = make(rb, mn, mx, g, tol, prec, thr)
    for mode in (robots.DIST_PROFILE, robots.DIST_MIN):
        eng.set_culling(True)
        eng.pair_count(reset=True)
        d1, p1, _ = eng.distance(q, mode)
        n1 = eng.pair_count(reset=True)
        eng.set_culling(False)
        d0, p0, _ = eng.distance(q, mode)
        n0 = eng.pair_count(reset=True)
        assert case in ("planar2", "xarm6_dense") or 0.0 < (d0 > 0).mean() < 1.0
        assert (d1.view(np.uint8) == d0.view(np.uint8)).all()
        if mode == robots.DIST_PROFILE:
            assert (p1.view(np.uint8) == p0.view(np.uint8)).all()
        assert n1 == n0


@pytest.mark.parametrize("prec", [robots.PREC_REF32, robots.PREC_FAST32])
@pytest.mark.parametrize("case", ["xarm6_51", "xarm6_1000", "xarm6_dense", "spatial22", "planar2"])
def test_small_batch_kernel_is_bit_identical_to_sequential_kernel(case, prec):
    """rvs_distance* calls of up to 4096 configurations (8192 for FULL) take distance_cta_kernel (one CTA per configuration, rvs_store.cuh): every
    output of every mode — d_c, the profile with its early-exit fill, the nearest points of collision-free configurations and the
    reference-order pair count — must equal the one-thread-per-configuration kernel (rvs_set_culling(0)) bit for bit."""
    if case == "xarm6_51":
        rb, mn, mx, g, tol = scenes.xarm6_scene(51, "D2", 0.45)
    elif case == "xarm6_1000":
        rb, mn, mx, g, tol = scenes.xarm6_scene(1000, "D2", 0.9)
    elif case == "xarm6_dense":
        rb, mn, mx, g, tol = scenes.xarm6_scene(300, "D2", 0.0)
    elif case == "spatial22":
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(22), 40)
    else:
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    eng = make(rb, mn, mx, g, tol, prec, 1e-4)
    for N in (1, 7, 300, 4096):
        q = synthetic.configs(rb, N, seed=0x51 + N)
        for mode in (robots.DIST_MIN, robots.DIST_PROFILE, robots.DIST_FULL):
            eng.set_culling(True)
            eng.pair_count(reset=True)
            d1, p1, n1 = eng.distance(q, mode)
            c1 = eng.pair_count(reset=True)
            eng.set_culling(False)
            d0, p0, n0 = eng.distance(q, mode)
            c0 = eng.pair_count(reset=True)
            # REF32 (no FMA contraction): bit for bit. FAST32 lets the compiler contract differently at the two inlining sites.
            same = (lambda x, y: (x.view(np.uint8) == y.view(np.uint8)).all()) if prec == robots.PREC_REF32 else \
                   (lambda x, y: np.abs(x - y).max(initial=0.0) <= 1e-5)
            assert same(d1, d0) and ((d1 == 0) == (d0 == 0)).all()
            assert c1 == c0 or prec != robots.PREC_REF32
            if mode >= robots.DIST_PROFILE:
                assert same(p1, p0)
            if mode == robots.DIST_FULL:
                free = d0 > 0
                if prec == robots.PREC_REF32:
                    assert same(n1[free], n0[free])
                elif free.any():     # FAST32: among equidistant candidates a differently contracted product may pick another point
                    assert (np.abs(n1[free] - n0[free]) <= 1e-5).mean() > 0.98
    eng.close()


@pytest.mark.parametrize("prec", [robots.PREC_REF32, robots.PREC_FAST32, robots.PREC_F64])
@pytest.mark.parametrize("case", ["xarm6_16_sorted", "xarm6_51", "xarm6_300", "spatial10", "planar2"])
def test_full_mode_of_a_large_batch_equals_sequential_kernel(case, prec):
    """FULL beyond the small-batch cross-over is three launches (rvs_distance_full.cuh): PROFILE through the work-queue kernel, the list of
    collision-free configurations, their nearest points. d_c, the profile with its early-exit fill, the reference-order pair count and the
    nearest points of every collision-free configuration must equal distance_kernel<FULL> (rvs_set_culling(0)) bit for bit; the block of a
    configuration in collision is unspecified (include/rvs_cuda.h; the device entry point leaves it as the caller passed it, which
    test_full_mode_at_scale_nearest_points_reproduce_the_profile checks)."""
    if case == "xarm6_16_sorted":
        (rb, mn, mx, g, tol), N = scenes.xarm6_scene(16, "D2", 0.45), 70000        # >= 65536: coherent-warp permutation in front
    elif case == "xarm6_51":
        (rb, mn, mx, g, tol), N = scenes.xarm6_scene(51, "D1", 0.3), 20000
    elif case == "xarm6_300":
        (rb, mn, mx, g, tol), N = scenes.xarm6_scene(300, "D2", 0.8), 3000         # long table: the cross-over is 1024, two obstacle tiles
    elif case == "spatial10":
        (rb, mn, mx, g, tol), N = scenes.chain_scene(robots.spatial(10), 20), 20000
    else:
        (rb, mn, mx, g, tol), N = scenes.chain_scene(robots.planar(2), 17), 20000          # two links: 48-byte rows, odd pair count
    eng = make(rb, mn, mx, g, tol, prec, 1e-4)
    q = synthetic.configs(rb, N, seed=0xF011)
    eng.set_culling(True)
    eng.pair_count(reset=True)
    k0 = eng.launch_count()
    d1, p1, n1 = eng.distance(q, robots.DIST_FULL)
    assert eng.launch_count() - k0 >= 3
    c1 = eng.pair_count(reset=True)
    eng.set_culling(False)
    d0, p0, n0 = eng.distance(q, robots.DIST_FULL)
    c0 = eng.pair_count(reset=True)
    eq = (lambda x, y: (x.view(np.uint8) == y.view(np.uint8)).all()) if prec != robots.PREC_FAST32 else \
         (lambda x, y: np.abs(x - y).max(initial=0.0) <= 1e-5)
    free = d0 > 0
    assert 0.02 < free.mean() < 0.999
    assert eq(d1, d0) and ((d1 > 0) == free).all() and eq(p1, p0)
    assert c1 == c0 or prec == robots.PREC_FAST32
    if prec != robots.PREC_FAST32:
        assert eq(n1[free], n0[free])
    elif free.any():
        assert (np.abs(n1[free] - n0[free]) <= 1e-5).mean() > 0.98
    eng.close()


def test_full_mode_large_batch_one_range_over_two_obstacle_tiles(monkeypatch):
    """RVS_FULL_PARTS_TARGET=1: every block of configurations walks the whole table itself — 300 obstacles = two shared-memory tiles, the
    rings drained in between — instead of one short range per block; same bytes as the default split and as distance_kernel<FULL>."""
    rb, mn, mx, g, tol = scenes.xarm6_scene(300, "D2", 0.8)
    eng = make(rb, mn, mx, g, tol)
    q = synthetic.configs(rb, 2500, seed=0x7117)
    d1, p1, n1 = eng.distance(q, robots.DIST_FULL)
    monkeypatch.setenv("RVS_FULL_PARTS_TARGET", "1")
    k0 = eng.launch_count()
    d2, p2, n2 = eng.distance(q, robots.DIST_FULL)
    assert eng.launch_count() - k0 >= 3
    monkeypatch.delenv("RVS_FULL_PARTS_TARGET")
    eng.set_culling(False)
    d0, p0, n0 = eng.distance(q, robots.DIST_FULL)
    free = d0 > 0
    assert 0.3 < free.mean() < 0.9
    for d, p, n in ((d1, p1, n1), (d2, p2, n2)):
        assert same(d, d0).all() and same(p, p0).all() and same(n[free], n0[free]).all()
    eng.close()


def test_full_mode_large_batch_with_a_misaligned_result_block():
    """A caller's nearest-point block that is only 4-byte aligned takes scalar stores instead of 8-byte ones: same bytes."""
    import torch
    rb, mn, mx, g, tol = scenes.xarm6_scene(20, "D2", 0.45)
    eng = make(rb, mn, mx, g, tol)
    N, M, L = 18000, 20, 6
    d_q = torch.from_numpy(synthetic.configs(rb, N, seed=0xA11)).cuda()
    d_dc, d_p = torch.empty(N, dtype=torch.float32, device="cuda"), torch.empty((N, L), dtype=torch.float32, device="cuda")
    out = []
    for off in (0, 1):
        buf = torch.zeros(N * M * L * 6 + 4, dtype=torch.float32, device="cuda")
        eng.distance_dev(d_q.data_ptr(), N, robots.DIST_FULL, d_dc.data_ptr(), d_p.data_ptr(), buf.data_ptr() + 4 * off)
        torch.cuda.synchronize()
        out.append(buf[off:off + N * M * L * 6].cpu().numpy())
    assert (out[0].view(np.uint8) == out[1].view(np.uint8)).all() and (out[0] != 0).any()
    eng.close()


@pytest.mark.parametrize("prec", [robots.PREC_REF32, robots.PREC_FAST32, robots.PREC_F64])
@pytest.mark.parametrize("case", ["xarm6_1000", "xarm6_2500_dense", "spatial22", "planar2"])
def test_work_queue_validity_kernels_equal_sequential_kernels(case, prec):
    """isValid(q) / isValid(q1,q2) through the scan + queue kernel (rvs_valid_wq.cuh) vs the one-thread-per-configuration kernels."""
    if case == "xarm6_1000":
        rb, mn, mx, g, tol = scenes.xarm6_scene(1000, "D2", 0.7)
    elif case == "xarm6_2500_dense":
        rb, mn, mx, g, tol = scenes.xarm6_scene(2500, "D1", 0.3)
    elif case == "spatial22":
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(22), 40)
    else:
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    eng = make(rb, mn, mx, g, tol, prec)
    q = synthetic.configs(rb, 30000)
    q1, q2 = synthetic.edges(rb, 5000, 0.1)
    eng.set_culling(True)
    v1, e1 = eng.is_valid(q), eng.is_valid_edges(q1, q2)
    eng.set_culling(False)
    v0, e0 = eng.is_valid(q), eng.is_valid_edges(q1, q2)
    assert (v1 == v0).all() and (e1 == e0).all()
    assert case == "xarm6_2500_dense" or 0.02 < v0.mean() < 0.999
    # planner-sized batches against a long table are split over slices of the (sorted) table: same verdicts
    eng.set_culling(True)
    for n_small in (1, 33, 2000):
        assert (eng.is_valid(q[:n_small]) == v0[:n_small]).all()
        assert (eng.is_valid_edges(q1[:n_small], q2[:n_small]) == e0[:n_small]).all()


def test_distance_full_nearest_points(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(51, "D1", 0.3)
    q = synthetic.configs(rb, 3000)
    eng, o = make(rb, mn, mx, g, tol), make_oracle(ob, rb, mn, mx, g, tol)
    d, p, near = eng.distance(q, robots.DIST_FULL)
    d_o, p_o, near_o = o.distance(q, robots.DIST_FULL)
    free = d_o > 0
    assert free.sum() > 300 and ((d > 0) == free).all()
    assert np.abs(near[free] - near_o[free]).max() <= TOL32
    assert (near[free] != near_o[free]).mean() < 1e-3
    # ground placeholder planes for the first ground_included links (RealVectorSpace.cpp:336-341)
    assert (near[free][:, 0, :2, :3] == 0).all() and (near[free][:, 0, :2, 5] == -1000.0).all()


def test_is_valid_and_edges_vs_oracle(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(40, "D2", 0.3)
    q = synthetic.configs(rb, 20000)
    eng, o = make(rb, mn, mx, g, tol), make_oracle(ob, rb, mn, mx, g, tol)
    v, v_o = eng.is_valid(q), o.is_valid(q)
    assert 0.05 < v_o.mean() < 0.95
    assert (v != v_o).sum() == 0
    q1, q2 = synthetic.edges(rb, 6000, 0.1)
    e, e_o = eng.is_valid_edges(q1, q2), o.is_valid_edges(q1, q2)
    assert 0.05 < e_o.mean() < 0.95
    assert (e != e_o).sum() == 0
    # an edge is valid iff all its interpolated configurations are: q2 itself is the first check (k = N)
    assert (e <= eng.is_valid(q2)).all()


def test_underestimation_vs_oracle_and_reference_invariant(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(30, "D2", 0.4)
    tol[:] = 1000.0
    eng, o = make(rb, mn, mx, g, tol), make_oracle(ob, rb, mn, mx, g, tol)
    roots = synthetic.configs(rb, 400)
    d, p, near = eng.distance(roots, robots.DIST_FULL)
    idx = np.flatnonzero(d > 0)[:64]
    assert len(idx) >= 16
    planes = near[idx]
    rng = np.random.default_rng(3)
    N = 8000
    pick = rng.integers(0, len(idx), N).astype(np.int32)
    q = (roots[idx][pick] + rng.normal(0, 0.15, (N, rb.num_dofs))).astype(np.float32)
    du, pu, st = eng.underestimation(q, planes, pick)
    du_o, pu_o, st_o = o.underestimation(q, planes, pick)
    assert (st == st_o).all() and 0.02 < st_o.mean() < 0.98
    assert np.abs(du - du_o).max() <= TOL32 and np.abs(pu - pu_o).max() <= TOL32
    # apps/test_real_vector_space.cpp:46-57 — at the generating configuration the planes reproduce the real distance
    own, own_p, own_s = eng.underestimation(roots[idx], planes, np.arange(len(idx), dtype=np.int32))
    assert (own_s == 0).all() and np.abs(own - d[idx]).max() <= 1e-3


def test_f64_variant_vs_oracle_double(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(60, "D2", 0.5)
    q = synthetic.configs(rb, 8000)
    eng, o = make(rb, mn, mx, g, tol, robots.PREC_F64), make_oracle(ob, rb, mn, mx, g, tol)
    d, p, _ = eng.distance(q, robots.DIST_PROFILE)
    d_o, p_o, _ = o.distance(q, robots.DIST_PROFILE, f64=True)
    assert d.dtype == np.float64
    assert ((d == 0) == (d_o == 0)).all()
    assert np.abs(d - d_o).max() <= TOL64 and np.abs(p - p_o).max() <= TOL64
    assert (eng.is_valid(q) == o.is_valid(q, f64=True)).all()


def test_fast32_variant_within_stated_tolerance(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(60, "D2", 0.5)
    q = synthetic.configs(rb, 20000)
    eng, o = make(rb, mn, mx, g, tol, robots.PREC_FAST32), make_oracle(ob, rb, mn, mx, g, tol)
    d, p, _ = eng.distance(q, robots.DIST_PROFILE)
    d_o, p_o, _ = o.distance(q, robots.DIST_PROFILE)
    d64, _, _ = o.distance(q, robots.DIST_MIN, f64=True)
    band = np.abs(d64) <= 1e-4                                          # |true clearance| within the epsilon band
    flips = (d == 0) != (d_o == 0)
    assert not (flips & ~band).any()
    both = (d > 0) & (d_o > 0)
    assert np.abs(d[both] - d_o[both]).max() <= TOL32
    v, v_o = eng.is_valid(q), o.is_valid(q)
    assert (v != v_o).mean() < 2e-4


@pytest.mark.parametrize("case", ["planar2_scenario_test", "xarm6_scenario1", "planar10", "spatial10", "spatial22"])
def test_other_robots_and_shipped_scenes(ob, case):
    if case == "planar2_scenario_test":
        rb, mn, mx, g, tol = scenes.planar2_scenario_test()
    elif case == "xarm6_scenario1":
        rb, mn, mx, g, tol = scenes.xarm6_scenario1()
    elif case == "planar10":
        rb, mn, mx, g, tol = scenes.chain_scene(robots.planar(10), 4)
    elif case == "spatial10":
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(10), 6)
    else:
        rb, mn, mx, g, tol = scenes.chain_scene(robots.spatial(22), 6)
    q = synthetic.configs(rb, 10000)
    eng, o = make(rb, mn, mx, g, tol), make_oracle(ob, rb, mn, mx, g, tol)
    d, p, near = eng.distance(q, robots.DIST_FULL)
    d_o, p_o, near_o = o.distance(q, robots.DIST_FULL)
    assert 0.02 < (d_o > 0).mean() < 0.999
    assert ((d == 0) == (d_o == 0)).all()
    assert np.abs(d - d_o).max() <= TOL32 and np.abs(p - p_o).max() <= TOL32
    free = d_o > 0
    assert np.abs(near[free] - near_o[free]).max() <= TOL32
    assert (eng.is_valid(q) != o.is_valid(q)).sum() == 0
    q1, q2 = synthetic.edges(rb, 3000, 0.1)
    assert (eng.is_valid_edges(q1, q2) != o.is_valid_edges(q1, q2)).sum() == 0


def test_edge_cases_and_errors(ob):
    rb, mn, mx, g, tol = scenes.xarm6_scene(5, "D2")
    eng = engine.Engine(rb)
    q = synthetic.configs(rb, 130)
    d, p, near = eng.distance(q, robots.DIST_FULL)                     # M = 0: no obstacles at all
    assert (d == 1000.0).all() and (p == 1000.0).all() and near.shape == (130, 0, 6, 6)
    assert eng.is_valid(q).all() and eng.is_valid_edges(q, q[::-1].copy()).all()
    eng.set_obstacles(mn, mx, g, tol)
    assert eng.distance(q[:0], robots.DIST_PROFILE)[0].shape == (0,)   # N = 0
    assert eng.is_valid(q[:0]).shape == (0,)
    one = eng.distance(q[:1], robots.DIST_PROFILE)                    # N = 1 == batch slice
    whole = eng.distance(q, robots.DIST_PROFILE)
    assert one[0][0] == whole[0][0] and (one[1][0] == whole[1][0]).all()
    # min_dist_tol: distances above the tolerance become MAX_DISTANCE (RealVectorSpace.cpp:367-368)
    o = make_oracle(ob, rb, mn, mx, g, np.full(5, 0.1, np.float32))
    eng.set_obstacles(mn, mx, g, np.full(5, 0.1, np.float32))
    dd, pp, _ = eng.distance(q, robots.DIST_PROFILE)
    d_o, p_o, _ = o.distance(q, robots.DIST_PROFILE)
    assert (pp == 1000.0).any() and np.abs(pp - p_o).max() <= TOL32
    # obstacle motion: update_obstacle_aabbs == set_obstacles with the moved boxes
    shift = np.array([0.07, -0.03, 0.02], np.float32)
    eng.update_obstacle_aabbs(mn + shift, mx + shift)
    o.set_obstacles(mn + shift, mx + shift, g, np.full(5, 0.1, np.float32))
    assert np.abs(eng.distance(q, robots.DIST_MIN)[0] - o.distance(q, robots.DIST_MIN)[0]).max() <= TOL32
    with pytest.raises(engine.RvsError):
        eng.update_obstacle_aabbs(mn[:2], mx[:2])                      # M differs from the resident table
    with pytest.raises(engine.RvsError):
        eng.distance(q, 7)                                             # bad mode
    with pytest.raises(engine.RvsError):
        engine.Engine(rb, robots.default_params(robots.PREC_F64)).skeleton(q)
    with pytest.raises(engine.RvsError):
        engine.Engine(rb, device=10 ** 6)
    bad = robots.xarm6()
    bad.axis[2] = 0.0
    with pytest.raises(engine.RvsError):
        engine.Engine(bad)


def test_device_pointer_api_and_generator(ob):
    torch = pytest.importorskip("torch")
    rb, mn, mx, g, tol = scenes.xarm6_scene(64, "D2", 0.6)
    eng = make(rb, mn, mx, g, tol)
    N = 50000
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
        d_dc = torch.empty(N, dtype=torch.float32, device="cuda")
        d_prof = torch.empty((N, 6), dtype=torch.float32, device="cuda")
        d_val = torch.empty(N, dtype=torch.uint8, device="cuda")
        eng.generate_configs_dev(synthetic.SEED_CONFIGS, 1000, N, d_q.data_ptr())
        eng.distance_dev(d_q.data_ptr(), N, robots.DIST_PROFILE, d_dc.data_ptr(), d_prof.data_ptr())
        eng.is_valid_dev(d_q.data_ptr(), N, d_val.data_ptr())
    stream.synchronize()
    q = synthetic.configs(rb, N, first=1000)
    assert (d_q.cpu().numpy() == q).all()                              # device generator == numpy generator, bit for bit
    o = make_oracle(ob, rb, mn, mx, g, tol)
    d_o, p_o, _ = o.distance(q, robots.DIST_PROFILE, threads=8)
    assert np.abs(d_dc.cpu().numpy() - d_o).max() <= TOL32
    assert np.abs(d_prof.cpu().numpy() - p_o).max() <= TOL32
    assert (d_val.cpu().numpy() == o.is_valid(q, threads=8)).all()


@pytest.mark.parametrize("robot,M,dist,r_min", [("xarm6", 1000, "D2", 0.9), ("xarm6", 300, "D1", 0.0), ("spatial10", 120, "D2", 0.3), ("planar2", 40, "D2", 0.0)])
def test_large_batch_switches_do_not_change_any_result(ob, robot, M, dist, r_min):
    """The two result-neutral devices of the large-batch path — thresholds bootstrapped from the coarse distance field
    (rvs_set_distance_field) and the coherent-warp ordering of the batch (rvs_set_config_sort) — against the same kernels with both
    switched off, bit for bit, and against the oracle on a sample: distances (MIN / PROFILE), isValid and edge validity."""
    rb = {"xarm6": robots.xarm6, "spatial10": lambda: robots.spatial(10), "planar2": lambda: robots.planar(2)}[robot]()
    if robot == "xarm6":
        mn, mx, g, tol = synthetic.sweep_obstacles(M, dist, r_min=r_min)
    else:
        reach = float(np.abs(rb.xyz).sum())
        mn, mx, g, tol = synthetic.sweep_obstacles(M, dist, r_min=r_min * reach, ws_center=(0.0, 0.0, 0.0), ws_radius=1.2 * reach)
        if rb.ground_included == 0:
            g = np.zeros_like(g)
    tol = tol.copy()
    tol[1::7] = 0.08                                                       # some obstacles with a min_dist_tol: they must stay out of the field
    N = 200_000
    q = synthetic.configs(rb, N, seed=0xC0FE)
    q1, q2 = synthetic.edges(rb, 70_000, 0.3, seed=0xC0FF)
    out = {}
    for df, srt in ((1, 1), (0, 0), (1, 0), (0, 1)):
        eng = make(rb, mn, mx, g, tol)
        eng.set_distance_field(bool(df)); eng.set_config_sort(bool(srt))
        out[(df, srt)] = (eng.distance(q, robots.DIST_MIN)[0], *eng.distance(q, robots.DIST_PROFILE)[:2], eng.is_valid(q), eng.is_valid_edges(q1, q2))
        if df:
            out[(df, srt, "second")] = eng.second_pass_warps()
        eng.close()
    base = out[(0, 0)]
    for key in ((1, 1), (1, 0), (0, 1)):
        for a, b in zip(out[key], base):
            assert same(a, b).all(), key
    assert out[(1, 1, "second")] < 0.01 * N / 32                            # the second pass is the exception
    o = make_oracle(ob, rb, mn, mx, g, tol)
    sample = slice(0, 4000)
    d_o, p_o, _ = o.distance(q[sample], robots.DIST_PROFILE)
    assert np.abs(base[1][sample] - d_o).max() <= TOL32 and np.abs(base[2][sample] - p_o).max() <= TOL32
    assert (base[3][sample] != o.is_valid(q[sample])).sum() == 0
    assert 0 < base[3].mean() < 1 or robot != "xarm6"
    # the other two arithmetic variants go through the same switches (F64: its own kernel instantiation, minima kept as doubles)
    for prec in (robots.PREC_FAST32, robots.PREC_F64):
        res = {}
        for df, srt in ((1, 1), (0, 0)):
            eng = make(rb, mn, mx, g, tol, prec)
            eng.set_distance_field(bool(df)); eng.set_config_sort(bool(srt))
            res[(df, srt)] = (*eng.distance(q[:120_000], robots.DIST_PROFILE)[:2], eng.is_valid(q[:120_000]))
            eng.close()
        for a, b in zip(res[(1, 1)], res[(0, 0)]):
            assert same(a, b).all(), prec


@pytest.mark.parametrize("eps,length", [(0.1, 0.3), (1.0, 3.0), (0.05, 6.0), (0.1, 0.0)])
def test_warp_per_edge_self_collision_equals_the_thread_per_edge_kernel(ob, eps, length):
    """xArm6::checkSelfCollision(q1, q2&) for planner-sized batches (one warp per edge: cooperative FK in the bubble iterations, the
    marching points tested 32 at a time) against the one-thread-per-edge kernel that larger batches take, bit for bit — verdicts and
    truncated configurations — and both against the oracle."""
    rb = robots.xarm6()
    eng = engine.Engine(rb)
    eng.set_self_collision(True, eps)
    E = 3000
    q1, q2 = synthetic.edges(rb, 9000, max(length, 1e-3), seed=0x5E1F)
    if length == 0.0:
        q2 = q1.copy()                                                 # zero-length motions: the bubble steps are inf / NaN as in the reference
    hit_w, out_w = eng.self_collision_edges(q1[:E], q2[:E])            # <= 8192 edges: warp kernel
    hit_t, out_t = eng.self_collision_edges(q1, q2)                    # 9000 edges: thread kernel
    assert np.array_equal(hit_w, hit_t[:E]) and same(out_w, out_t[:E]).all()
    # rvs_check_edges_detailed = the two separate calls in one
    mn, mx, g, tol = synthetic.sweep_obstacles(40, "D2")
    eng.set_obstacles(mn, mx, g, tol)
    ok_d, hit_d, out_d = eng.check_edges_detailed(q1[:E], q2[:E])
    assert np.array_equal(ok_d, eng.is_valid_edges(q1[:E], q2[:E]).astype(bool)) and np.array_equal(hit_d, hit_w.astype(bool)) and same(out_d, out_w).all()
    assert np.array_equal(eng.check_edges(q1[:E], q2[:E]).astype(bool), ok_d & ~hit_d)
    one_h, one_q = eng.self_collision_edges(q1[7:8], q2[7:8])
    assert one_h[0] == hit_t[7] and same(one_q[0], out_t[7]).all()
    o = ob.Oracle(rb)
    o.set_self_collision(True, eps)
    hit_o, out_o = o.self_collision_edges(q1[:E], q2[:E])
    assert (hit_w != hit_o).mean() <= 0.005                            # FK last-ulp band, as in tests/test_robot_golden.py
    agree = hit_w == hit_o
    assert (np.abs(out_w[agree] - out_o[agree]).max(axis=1) <= 1e-5).mean() >= 0.995
    if length >= 3.0:
        assert 0 < hit_w.sum() < E
    eng.close()


def test_full_size_properties():
    """BASELINE sweep size (2^20 configurations x 1000 boxes): size-independent properties instead of an oracle run."""
    torch = pytest.importorskip("torch")
    rb, mn, mx, g, tol = scenes.xarm6_scene(1000, "D2", 0.9)
    eng = make(rb, mn, mx, g, tol)
    N = 1 << 20
    d_q = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    out = {k: torch.empty(N, dtype=torch.float32, device="cuda") for k in ("min", "prof", "perm", "again")}
    prof = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    prof2 = torch.empty((N, 6), dtype=torch.float32, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_MIN, out["min"].data_ptr())
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_PROFILE, out["prof"].data_ptr(), prof.data_ptr())
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_MIN, out["again"].data_ptr())
    perm = np.random.default_rng(0).permutation(1000)
    eng.set_obstacles(mn[perm], mx[perm], g[perm], tol[perm])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_PROFILE, out["perm"].data_ptr(), prof2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out["min"], out["prof"]) and torch.equal(out["min"], out["again"])      # modes agree, idempotent
    free = out["min"] > 0
    assert 0.4 < free.float().mean().item() < 0.9
    assert torch.equal(prof[free].min(dim=1).values, out["min"][free])                        # d_c = min of the profile
    hit = ~free
    first = (prof[hit] <= 0).float().argmax(dim=1)
    cols = torch.arange(6, device="cuda")[None, :]
    assert (prof[hit][cols > first[:, None]] == 0).all()                                       # early-exit fill
    # the minimum over obstacles does not depend on their order (collision-free configurations)
    assert torch.equal(free, out["perm"] > 0)
    assert torch.equal(prof[free], prof2[free])
    # verdicts: a configuration whose distance query collides is (outside measure-zero cases) invalid, and vice versa
    val = torch.empty(N, dtype=torch.uint8, device="cuda")
    eng.is_valid_dev(d_q.data_ptr(), N, val.data_ptr())
    torch.cuda.synchronize()
    assert ((val == 1) != free).float().mean().item() < 1e-3


def test_full_mode_at_scale_nearest_points_reproduce_the_profile(ob):
    """300 000 configurations x 51 boxes with nearest points (2.2 GB of results, the three-launch path with a sorted PROFILE pass, rings
    and two obstacle ranges per block): for every collision-free configuration the profile is recovered from the nearest points alone —
    min over the obstacles of |P_robot - P_obstacle| - radius, with the min_dist_tol rule — and a sample of 2000 configurations equals
    the CPU oracle's points."""
    torch = pytest.importorskip("torch")
    rb, mn, mx, g, tol = scenes.xarm6_scene(51, "D1", 0.3)
    eng = make(rb, mn, mx, g, tol)
    N, M, L = 300000, 51, 6
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    d_q = torch.empty((N, L), dtype=torch.float32, device="cuda")
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 0, N, d_q.data_ptr())
    d_dc, d_p = torch.empty(N, dtype=torch.float32, device="cuda"), torch.empty((N, L), dtype=torch.float32, device="cuda")
    near = torch.zeros((N, M, L, 6), dtype=torch.float32, device="cuda")
    k0 = eng.launch_count()
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_FULL, d_dc.data_ptr(), d_p.data_ptr(), near.data_ptr())
    torch.cuda.synchronize()
    assert eng.launch_count() - k0 >= 3
    free = d_dc > 0
    assert 0.3 < free.float().mean().item() < 0.5
    assert (near[~free] == 0).all()                                                              # blocks of colliding configurations: untouched
    radius = torch.tensor(rb.radius[:L], dtype=torch.float32, device="cuda")
    t_tol = torch.tensor(tol, dtype=torch.float32, device="cuda")
    worst = 0.0
    idx = torch.nonzero(free).flatten()
    for chunk in torch.split(idx, 20000):
        nb = near[chunk]
        d = (nb[..., :3] - nb[..., 3:]).norm(dim=-1) - radius[None, None, :]                     # [n, M, L]
        clear = ((d - t_tol[None, :, None]).abs() > 1e-4).all(dim=2).all(dim=1)                  # (not within rounding of a min_dist_tol)
        d = torch.where(d > t_tol[None, :, None], torch.full_like(d, 1000.0), d)                 # min_dist_tol (:367-368)
        worst = max(worst, (d.min(dim=1).values - d_p[chunk])[clear].abs().max().item())
    assert worst <= 1e-5
    # a sample against the oracle
    o = make_oracle(ob, rb, mn, mx, g, tol)
    sel = torch.arange(0, N, 150, device="cuda")
    q_s = d_q[sel].cpu().numpy()
    d_o, p_o, n_o = o.distance(q_s, robots.DIST_FULL)
    f_o = d_o > 0
    assert ((d_dc[sel].cpu().numpy() > 0) == f_o).all()
    n_g = near[sel].cpu().numpy()
    assert np.abs(n_g[f_o] - n_o[f_o]).max() <= TOL32 and (n_g[f_o] != n_o[f_o]).mean() < 1e-3
    eng.close()


def test_headline_step_of_1e7_configurations_sampled_against_the_oracle(ob):
    """bench.py's step as it is timed — 10^7 device-generated configurations against the headline scene in ONE launch of the work-queue
    kernel — with 40 000 configurations sampled out of it (stride 250) and re-computed by the CPU oracle: bit-identical up to rare
    FK last-ulp effects, plus the size-independent properties on the whole batch (d_c == min(profile) for free configurations, profile
    tail zero-filled after the colliding link, d_c == 0 exactly when the profile holds a non-positive entry)."""
    import torch
    import bench
    rb, mn, mx, g, tol = bench.scene("headline")
    N, n, L = 10_000_000, rb.num_dofs, rb.num_links
    eng = make(rb, mn, mx, g, tol)
    dev = torch.device("cuda", 0)
    d_q = torch.empty((N, n), dtype=torch.float32, device=dev)
    d_dc, d_prof = torch.empty(N, dtype=torch.float32, device=dev), torch.empty((N, L), dtype=torch.float32, device=dev)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.generate_configs_dev(synthetic.SEED_CONFIGS, 123 * N, N, d_q.data_ptr())
    eng.distance_dev(d_q.data_ptr(), N, robots.DIST_PROFILE, d_dc.data_ptr(), d_prof.data_ptr())
    torch.cuda.synchronize()
    free = d_dc > 0
    assert 0.5 < float(free.float().mean()) < 0.8
    assert bool((d_prof[free].min(dim=1).values == d_dc[free]).all())                 # d_c = min over links (RealVectorSpace.cpp:370-384)
    hit = ~free
    first_bad = (d_prof[hit] <= 0).float().argmax(dim=1)
    assert bool(((d_prof[hit] <= 0).any(dim=1)).all())                                # a collision leaves a non-positive entry ...
    tail = torch.arange(L, device=dev)[None, :] > first_bad[:, None]
    assert bool((d_prof[hit][tail] == 0).all())                                       # ... and zeros behind it (:371-378)
    idx = torch.arange(0, N, 250, device=dev)
    q_s, d_s, p_s = d_q[idx].cpu().numpy(), d_dc[idx].cpu().numpy(), d_prof[idx].cpu().numpy()
    i = (np.uint64(123 * N) + np.arange(0, N, 250, dtype=np.uint64))[:, None]
    u = synthetic.unit24(synthetic.SEED_CONFIGS, i * np.uint64(32) + np.arange(n, dtype=np.uint64)[None, :])
    assert np.array_equal(q_s, (rb.limits[:, 0][None, :] + (rb.limits[:, 1] - rb.limits[:, 0])[None, :] * u).astype(np.float32))   # device generator == host generator
    o = make_oracle(ob, rb, mn, mx, g, tol)
    d_o, p_o, _ = o.distance(q_s, robots.DIST_PROFILE, threads=8)
    assert ((d_s == 0) == (d_o == 0)).all()
    assert np.abs(d_s - d_o).max() <= TOL32 and np.abs(p_s - p_o).max() <= TOL32
    assert (d_s != d_o).mean() < 1e-3 and (p_s != p_o).mean() < 1e-3
    eng.close()
