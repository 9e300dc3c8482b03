"""Rows A8 / A10 / A11 against tests/golden/robot_golden.npz — outputs of the REFERENCE'S OWN robot classes (src/robots/*.cpp compiled
unmodified against the KDL / kdl_parser / urdfdom / FCL / glog work-alikes of oracle/env_shim, built from the reference's own URDFs;
tests/golden/make_robot_golden.py).  CPU tests pin the oracle's restatements (computeSkeleton, computeEnclosingRadii,
computeCapsulesDistance, both checkSelfCollision overloads) bit for bit; GPU tests check the CUDA engine through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT
from rpmplv2_b200 import robots, synthetic

GOLD = os.path.join(ROOT, "tests", "golden", "robot_golden.npz")


def _mod():
    spec = importlib.util.spec_from_file_location("make_robot_golden", os.path.join(ROOT, "tests", "golden", "make_robot_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    try:
        spec.loader.exec_module(mod)
    except OSError:                               # refharness is only importable next to oracle/_ref; the generators' inputs are all we need
        raise
    return mod


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def mod():
    return _mod()


def test_reference_classes_report_the_tabulated_chains(gold, mod):
    """limits (narrowed to float like `float lower = ...->limits->lower`, xArm6.cpp:30-35) and radii as the reference's classes hold them."""
    for name, rb in mod.robots_under_test().items():
        assert np.array_equal(gold[f"{name}__limits"], rb.limits), name
        assert np.array_equal(gold[f"{name}__radii"], rb.radius), name


def test_reference_skeleton_known_answers(gold):
    """SURVEY.md §8.2 hand-derived known answers, checked on the REFERENCE's output (q = 0 is row 0 of every case)."""
    sk = gold["xarm6__skeleton"][0]
    want = np.array([(0, 0, 0), (0, 0, 0.267), (0.0535, 0, 0.5515), (0.131, 0, 0.459), (0.131, 0, 0.209), (0.207, 0, 0.209), (0.207, 0, -0.007)])
    assert np.abs(sk - want).max() < 5e-6                      # rpy = 1.5708 is not pi/2: ~4e-6
    sk = gold["planar_2dof__skeleton"][0]
    assert np.array_equal(sk, np.array([(0, 0, 0), (1, 0, 0), (2, 0, 0)], np.float32))
    sk = gold["spatial_10dof__skeleton"][0]
    assert np.abs(sk[:, 2] - np.concatenate([[0], 0.1 * np.arange(1, 11)])).max() < 1e-6 and np.abs(sk[:, :2]).max() == 0
    # frames: rotation blocks are orthonormal, translation of frame k = skeleton point k + 1 for the plain chains
    fr = gold["spatial_22dof__frames"]
    Rm = fr[..., :9].reshape(fr.shape[0], fr.shape[1], 3, 3)
    assert np.abs(Rm @ np.swapaxes(Rm, -1, -2) - np.eye(3)).max() < 1e-12
    assert np.array_equal(fr[:, :, 9:].astype(np.float32), gold["spatial_22dof__skeleton"][:32, :22])


def test_oracle_robot_rows_equal_the_reference_classes(ob, gold, mod):
    for name, rb in mod.robots_under_test().items():
        o = ob.Oracle(rb, robots.default_params())
        q = mod.configs(rb)
        assert np.array_equal(o.skeleton(q), gold[f"{name}__skeleton"]), name
        assert np.array_equal(o.enclosing_radii(q).reshape(gold[f"{name}__radii_matrix"].shape), gold[f"{name}__radii_matrix"]), name
        if rb.kind != robots.ROBOT_XARM6:
            o.set_self_collision(True)
            q1, q2 = mod.edge_inputs(rb, 0.5, 0x55)
            hit, out = o.self_collision_edges(q1, q2)
            assert not hit.any() and np.array_equal(out, q2)
            continue
        q_many = synthetic.configs(rb, 4000, seed=0xC1)
        mask = o.self_collision_capsules(q_many)
        bits = np.stack([(mask >> k) & 1 for k in range(4)], 1).astype(bool)
        assert np.array_equal(bits, gold[f"{name}__caps_d"] < 0), name          # pair order {0-5, 1-5, 0-4, 1-4}
        assert np.array_equal(mask != 0, gold[f"{name}__self_hit"] != 0), name
        for k, (eps, length, seed) in enumerate(mod.EDGE_SETS):
            q1, q2 = mod.edge_inputs(rb, length, seed)
            o.set_self_collision(True, eps)
            hit, out = o.self_collision_edges(q1, q2)
            assert np.array_equal(hit, gold[f"{name}__edge{k}_hit"]), (name, k)
            assert np.array_equal(out, gold[f"{name}__edge{k}_q2"]), (name, k)
        o.set_self_collision(False)
        hit, out = o.self_collision_edges(q1, q2)
        assert not hit.any() and np.array_equal(out, q2)


def test_oracle_robot_rows_equal_the_compiled_reference_on_fresh_inputs(ob):
    """The same comparison straight against oracle/_ref/librvs_ref.so on inputs outside the golden file (where /root/reference exists)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librvs_ref.so")) or not os.path.isdir("/root/reference/data"):
        pytest.skip("oracle/_ref and the reference's URDFs exist in the build container only")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refharness as rh
    lib = rh.load()
    for rb in (robots.xarm6(capsules_radius=(0.05, 0.1, 0.08, 0.07, 0.06, 0.04), gripper_length=0.1), robots.planar(10), robots.spatial(18)):
        h, used = rh.create(lib, rb)
        assert used
        o = ob.Oracle(rb, robots.default_params())
        q = synthetic.configs(rb, 3000, seed=0xF00)
        assert np.array_equal(o.skeleton(q), rh.skeleton(lib, h, q))
        assert np.array_equal(o.enclosing_radii(q).reshape(len(q), rb.num_dofs, -1), rh.enclosing_radii(lib, h, q))
        if rb.kind == robots.ROBOT_XARM6:
            d, hit = rh.self_collision(lib, h, q)
            assert np.array_equal(o.self_collision_capsules(q) != 0, hit != 0)
            for eps, length, seed in ((0.1, 0.8, 7), (0.05, 0.2, 8)):
                q1, q2 = synthetic.edges(rb, 3000, length, seed=seed)
                o.set_self_collision(True, eps)
                want_hit, want_q2 = rh.self_collision_edges(lib, h, q1, q2, eps)
                got_hit, got_q2 = o.self_collision_edges(q1, q2)
                assert np.array_equal(got_hit, want_hit) and np.array_equal(got_q2, want_q2)
                assert want_hit.sum() > 100
        lib.rvsref_destroy(h)


@pytest.mark.gpu
def test_gpu_robot_rows_vs_the_reference_classes(gold, mod):
    """CUDA engine through the C ABI against the reference classes' own outputs.  FK runs in FP64 on both sides but in a different
    operation order (device: R . Rot(axis, q) with the structural zeros of x / y / z axes removed; KDL: Rot2(R . axis, q) . R), so a
    skeleton coordinate may differ in the last float ulp: 2.4e-7 m; enclosing radii 1e-6; capsule verdicts are exact outside a band of
    1e-6 m around zero distance; an edge verdict / truncation point may only differ when some marching point's capsule distance, or
    a bubble step's comparison with 1, sits inside the same band."""
    from rpmplv2_b200 import engine
    for name, rb in mod.robots_under_test().items():
        eng = engine.Engine(rb, robots.default_params())
        q = mod.configs(rb)
        assert np.abs(eng.skeleton(q) - gold[f"{name}__skeleton"]).max() <= 2.4e-7, name
        assert np.abs(eng.enclosing_radii(q).reshape(gold[f"{name}__radii_matrix"].shape) - gold[f"{name}__radii_matrix"]).max() <= 1e-6, name
        if rb.kind == robots.ROBOT_XARM6:
            q_many = synthetic.configs(rb, 4000, seed=0xC1)
            mask = eng.self_collision_capsules(q_many)
            bits = np.stack([(mask >> k) & 1 for k in range(4)], 1).astype(bool)
            d = gold[f"{name}__caps_d"]
            outside = np.abs(d) > 1e-6
            assert np.array_equal(bits[outside], (d < 0)[outside]), name
            assert (bits != (d < 0)).sum() <= 2, name
            for k, (eps, length, seed) in enumerate(mod.EDGE_SETS):
                q1, q2 = mod.edge_inputs(rb, length, seed)
                eng.set_self_collision(True, eps)
                hit, out = eng.self_collision_edges(q1, q2)
                want_hit, want_q2 = gold[f"{name}__edge{k}_hit"], gold[f"{name}__edge{k}_q2"]
                same = hit == want_hit
                assert same.mean() >= 0.995, (name, k, int((~same).sum()))
                close = np.abs(out[same] - want_q2[same]).max(axis=1) <= 1e-5
                assert close.mean() >= 0.995, (name, k)
            eng.set_self_collision(False)
            hit, out = eng.self_collision_edges(q1, q2)
            assert not hit.any() and np.array_equal(out, q2)
        else:
            eng.set_self_collision(True)
            q1, q2 = mod.edge_inputs(rb, 0.5, 0x55)
            hit, out = eng.self_collision_edges(q1, q2)
            assert not hit.any() and np.array_equal(out, q2)
        eng.close()
