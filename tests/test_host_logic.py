"""CPU-side tests: host logic, C-ABI loading, the device headers compiled for the host, reference data parsing."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE_DATA, ROOT, same
from rpmplv2_b200 import engine, robots, synthetic


def test_abi_library_loads_and_exports_every_declared_symbol():
    lib = engine.load_library()
    header = open(os.path.join(ROOT, "include", "rvs_cuda.h")).read()
    declared = set(re.findall(r"\b(rvs_[a-z0-9_]+)\s*\(", header)) - {"rvs_status"}
    assert declared == set(engine.SYMBOLS), declared ^ set(engine.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.rvs_abi_version() == 1


def test_struct_layout_matches_header():
    assert ctypes.sizeof(robots.RvsRobotDesc) == 16 + 8 * 3 * 33 * 2 + 8 * 3 * 32 + 4 * 32 * 3
    assert ctypes.sizeof(robots.RvsParams) == 16 and ctypes.sizeof(robots.RvsSpineParams) == 16


@pytest.mark.skipif(engine.load_library().rvs_device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(engine.RvsError) as err:
        engine.Engine(robots.xarm6())
    assert err.value.code == 4 and "no CPU fallback" in str(err.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rpmplv2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle<", "").replace("oracle/", "ORC/") or f == "synthetic.py" or "import" not in text or "oracle_binding" not in text, f
                assert "oracle_binding" not in text and "librvs_oracle" not in text, f


def test_synthetic_generators_match_the_oracle_generator(ob):
    for rb in (robots.xarm6(), robots.spatial(22)):
        o = ob.Oracle(rb)
        assert (synthetic.configs(rb, 777, first=123456) == o.generate_configs(synthetic.SEED_CONFIGS, 123456, 777)).all()
    mn, mx, g, tol = synthetic.sweep_obstacles(200, "D2", r_min=0.9)
    assert g[0] == 1 and g[1:].sum() == 0 and tol[0] == np.float32(0.05)
    assert (mx > mn).all() and np.allclose(mn[0], [-0.75, -0.75, -0.1]) and np.allclose(mx[0], [0.75, 0.75, 0.0])
    assert synthetic.pair_queries(robots.xarm6(), g) == 6 * 200 - 2


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference data only exists in the build container")
def test_builtin_robot_tables_equal_the_reference_urdfs():
    cases = [("xarm6", "xarm6/xarm6.urdf", robots.xarm6()), ("planar_2dof", "planar_2dof/planar_2dof.urdf", robots.planar(2)),
             ("planar_10dof", "planar_10dof/planar_10dof.urdf", robots.planar(10))]
    cases += [(f"spatial_{n}dof", f"spatial_{n}dof/spatial_{n}dof.urdf", robots.spatial(n)) for n in (10, 14, 18, 22)]
    for rtype, rel, builtin in cases:
        parsed = robots.from_urdf(os.path.join(REFERENCE_DATA, rel), rtype)
        assert parsed.num_dofs == builtin.num_dofs
        for field in ("xyz", "rpy", "axis", "limits"):
            assert np.array_equal(getattr(parsed, field), getattr(builtin, field)), (rtype, field)


def test_device_geometry_headers_compiled_for_host_match_golden(golden, host_sim):
    """rvs_geom.cuh built with g++ (tests/host_sim) reproduces the reference's golden vectors bit for bit."""
    f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
    for tag, thr in (("t4", 1e-4), ("t6", 1e-6)):
        AB, r, obs, g = (np.ascontiguousarray(golden[f"{tag}_{k}"]) for k in ("AB", "r", "obs", "dist"))
        K = AB.shape[0]
        for pts in (0, 1):
            out = np.empty((K, 8), np.float32)
            host_sim.hs_capsule_box_distance_f32(ctypes.c_int64(K), AB.ctypes.data_as(f32p), r.ctypes.data_as(f32p), obs.ctypes.data_as(f32p),
                                                 ctypes.c_float(thr), ctypes.c_float(1000.0), ctypes.c_int(pts), out.ctypes.data_as(f32p))
            assert same(out[:, 0], g[:, 0]).all()
            if pts:
                live = (g[:, 1] == 1) & (g[:, 0] > 0)
                assert same(out[live, 2:], g[live, 2:]).all()
        hit = np.empty(K, np.uint8)
        host_sim.hs_capsule_box_collision_f32(ctypes.c_int64(K), AB.ctypes.data_as(f32p), r.ctypes.data_as(f32p), obs.ctypes.data_as(f32p),
                                              ctypes.c_float(thr), ctypes.c_float(1000.0), hit.ctypes.data_as(u8p))
        assert (hit == golden[f"{tag}_coll"]).all()


def test_deferred_edge_evaluation_of_the_full_mode_kernel_matches_golden(golden, host_sim):
    """nearest_points_kernel (rvs_distance_full.cuh) parks an undecided pair as a plan word + candidate distance and evaluates its box edges
    later; the candidate points are rebuilt from PairPlan::cand. The same steps on the host (tests/host_sim) give the reference's golden
    distances and nearest points bit for bit, and a fair share of the golden pairs does go through the parked path."""
    f32p = ctypes.POINTER(ctypes.c_float)
    for tag, thr in (("t4", 1e-4), ("t6", 1e-6)):
        AB, r, obs, g = (np.ascontiguousarray(golden[f"{tag}_{k}"]) for k in ("AB", "r", "obs", "dist"))
        K = AB.shape[0]
        out, seq = np.empty((K, 8), np.float32), np.empty((K, 8), np.float32)
        host_sim.hs_capsule_box_points_deferred_f32(ctypes.c_int64(K), AB.ctypes.data_as(f32p), r.ctypes.data_as(f32p), obs.ctypes.data_as(f32p),
                                                    ctypes.c_float(thr), ctypes.c_float(1000.0), out.ctypes.data_as(f32p))
        host_sim.hs_capsule_box_distance_f32(ctypes.c_int64(K), AB.ctypes.data_as(f32p), r.ctypes.data_as(f32p), obs.ctypes.data_as(f32p),
                                             ctypes.c_float(thr), ctypes.c_float(1000.0), ctypes.c_int(1), seq.ctypes.data_as(f32p))
        assert same(out[:, 0], g[:, 0]).all()
        live = (g[:, 1] == 1) & (g[:, 0] > 0)
        assert same(out[live, 2:], g[live, 2:]).all() and same(out[live, 2:], seq[live, 2:]).all()
        parked = out[:, 1] == 0
        assert parked.mean() > 0.1 and (parked & live).sum() > 100


def test_device_fk_header_compiled_for_host_matches_oracle(ob, host_sim):
    f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    for rb in (robots.xarm6(), robots.planar(10), robots.spatial(22)):
        q = synthetic.configs(rb, 3000)
        o = ob.Oracle(rb)
        s32, s64 = np.empty((3000, rb.num_dofs + 1, 3), np.float32), np.empty((3000, rb.num_dofs + 1, 3), np.float64)
        desc = rb.to_desc()
        assert host_sim.hs_skeleton(ctypes.byref(desc), q.ctypes.data_as(f32p), ctypes.c_int64(3000), ctypes.c_int(1),
                                    s32.ctypes.data_as(f32p), s64.ctypes.data_as(f64p)) == 0
        assert (s32 == o.skeleton(q)).all()
        assert np.abs(s64 - o.skeleton(q, f64=True)).max() < 1e-15
