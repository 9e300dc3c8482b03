"""Scenario / fixture I/O: rpmplv2_b200/scenario.py against fixtures of the reference's schema, against the
reference's own data/ files when they are present (build container only), and — on the GPU — the engine a scenario builds against
the oracle fed with the same arrays."""
import glob
import os

import numpy as np
import pytest

import scenes
from conftest import REFERENCE_DATA, ROOT
from rpmplv2_b200 import robots, scenario, synthetic

FIX = os.path.join(ROOT, "tests", "golden", "scenario_fixture.yaml")
RFIX = os.path.join(ROOT, "tests", "golden", "random_scenarios_fixture.yaml")


def test_fixture_loads_with_the_reference_rules():
    sc = scenario.load(FIX)
    rb = sc.robot
    assert rb.name == "xarm6" and rb.num_dofs == 6 and rb.ground_included == 2 and rb.gripper_length == pytest.approx(0.1)
    assert np.array_equal(rb.radius, np.full(6, 0.06, np.float32))                      # one value -> every link (Scenario.cpp:40-41)
    assert sc.space == "RealVectorSpaceCUDA" and sc.motion_type == "light_directions"
    assert sc.labels == ["ground", "", "dynamic_obstacle", ""] and list(sc.is_ground) == [1, 0, 0, 0] and list(sc.is_dynamic) == [0, 0, 1, 0]
    assert list(sc.rotated) == [0, 0, 1, 0]                                             # identity quaternion stays tight
    assert sc.min_dist_tol[0] == np.float32(0.04) and sc.min_dist_tol[1] == np.float32(1000.0)
    mn, mx = sc.aabbs()
    assert np.allclose(mn[1], [0.35, 0.3, 0.0]) and np.allclose(mx[1], [0.65, 0.5, 0.6])
    rad = np.sqrt(0.05 ** 2 + 0.1 ** 2 + 0.15 ** 2)                                     # bounding sphere of the rotated box
    assert np.allclose(mx[2] - mn[2], 2 * rad, atol=1e-6) and np.allclose((mx[2] + mn[2]) / 2, [-0.45, 0.35, 0.55], atol=1e-6)
    assert sc.ws_center == (0.0, 0.0, 0.267) and sc.ws_radius == pytest.approx(1.2)
    assert sc.base_radius() == pytest.approx(0.06 + np.sqrt(3 * 0.05 ** 2), rel=1e-6)   # max(r0, r1) + |obs_dim|
    ep = sc.env_params()
    assert ep.motion_type == 4 and ep.robot_max_vel == pytest.approx(2.5) and ep.ws_radius == pytest.approx(1.2)
    assert np.allclose(sc.q_start, [0, 0, 0, 3.1415, 1.5708, 0])


def test_ground_is_dropped_without_ground_included_and_other_spaces_are_refused(tmp_path):
    text = open(FIX).read()
    p = tmp_path / "a.yaml"
    p.write_text(text.replace("ground_included: 2", "ground_included: 0"))
    sc = scenario.load(str(p))
    assert sc.labels == ["", "dynamic_obstacle", ""] and sc.robot.ground_included == 0      # Environment.cpp:33-34
    p.write_text(text.replace('space: "RealVectorSpaceCUDA"', 'space: "RealVectorSpaceFCL"'))
    with pytest.raises(ValueError):
        scenario.load(str(p))
    p.write_text(text.replace("capsules_radius: [0.06]", "capsules_radius: [0.06, 0.05]"))
    with pytest.raises(ValueError):
        scenario.load(str(p))                                                               # "Number of capsules is not correct!"
    p.write_text(text.replace('type: "xarm6"', 'type: "ur5"'))
    with pytest.raises(Exception):
        scenario.load(str(p))                                                               # "Robot type is not correct!"


def test_random_scenarios_file_and_run_application():
    rs = scenario.load_random_scenarios(RFIX)
    assert sorted(rs) == [0, 2] and len(rs[0]) == 2 and len(rs[2]) == 1
    assert rs[2][0]["pos"].shape == (2, 3) and np.allclose(rs[2][0]["vel"][1], [0.1, 0.4, 0.2])
    sc = scenario.load(FIX)
    run = scenario.apply_random_run(sc, rs[2][0])
    assert run.num_objects == sc.num_objects + 2 and run.labels[-2:] == ["dynamic_obstacle"] * 2
    assert np.allclose(run.dim[-1], [0.05, 0.05, 0.05]) and np.allclose(run.max_vel[-2:], 0.8)
    assert np.allclose(run.q_start, rs[2][0]["start"]) and sc.num_objects == 4                  # the base scenario is untouched


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference data only exists in the build container")
def test_reference_scenarios_load_unchanged():
    root = os.path.dirname(REFERENCE_DATA)
    files = sorted(glob.glob(os.path.join(REFERENCE_DATA, "*", "*", "scenario*.yaml")))
    assert len(files) >= 10
    loaded = 0
    for f in files:
        rel = f[len(root):]
        if "robot:" not in open(f).read():
            continue                                    # two recorded test logs share the name pattern but are not scenarios
        try:
            sc = scenario.load(rel, root)
        except ValueError as e:
            assert "RealVectorSpaceFCL" in str(e), (rel, e)
            continue
        loaded += 1
        assert len(sc.q_start) == sc.robot.num_dofs == len(sc.q_goal)
        assert sc.dim.shape == (sc.num_objects, 3) and np.all(sc.dim > 0)
    assert loaded >= 8
    # the scenes the tests and benches rebuild by hand are exactly what the files say
    s1 = scenario.load("/data/xarm6/scenario1/scenario1.yaml", root)
    rb, mn, mx, g, tol = scenes.xarm6_scenario1()
    a, b = s1.aabbs()
    assert np.array_equal(a, mn) and np.array_equal(b, mx) and np.array_equal(s1.is_ground, g) and np.array_equal(s1.min_dist_tol, tol)
    assert np.array_equal(s1.robot.radius, rb.radius) and s1.robot.gripper_length == rb.gripper_length and s1.motion_type == "circular"
    assert np.array_equal(s1.robot.xyz, rb.xyz) and np.array_equal(s1.robot.axis, rb.axis)       # parsed from the reference URDF
    p2 = scenario.load("/data/planar_2dof/scenario_test/scenario_test.yaml", root)
    _, mn, mx, g, tol = scenes.planar2_scenario_test()
    a, b = p2.aabbs()
    assert np.array_equal(a, mn) and np.array_equal(b, mx)
    rs = scenario.load_random_scenarios(os.path.join(REFERENCE_DATA, "xarm6", "scenario_random_obstacles", "random_scenarios.yaml"), only=10)
    assert list(rs) == [10] and rs[10][0]["pos"].shape == (10, 3) and len(rs[10][0]["start"]) == 6 and len(rs[10]) >= 100


@pytest.mark.gpu
def test_gpu_engine_built_from_a_scenario_matches_the_oracle(ob):
    sc = scenario.load(FIX)
    rs = scenario.load_random_scenarios(RFIX)
    run = scenario.apply_random_run(sc, rs[2][0])
    q = np.vstack([run.q_start, run.q_goal, synthetic.configs(run.robot, 500, seed=0x5C)])
    # static table
    eng = run.make_engine(dynamic=False)
    o = ob.Oracle(run.robot, robots.default_params())
    mn, mx = run.aabbs()
    o.set_obstacles(mn, mx, run.is_ground, run.min_dist_tol)
    assert np.array_equal(eng.is_valid(q), o.is_valid(q))
    assert np.array_equal(eng.distance(q)[0], o.distance(q)[0])
    # dynamic environment (the default for a scenario with a motion type and dynamic obstacles)
    eng2 = run.make_engine()
    o.env_set(run.env_params(), run.dim, run.pos, vel=run.vel, max_vel=run.max_vel, is_dynamic=run.is_dynamic, is_ground=run.is_ground,
              min_dist_tol=run.min_dist_tol, rotated=run.rotated)
    eng2.env_step(0.01, 40)
    o.env_step(0.01, 40)
    assert np.array_equal(eng2.env_get()[0], o.env_get()[0])
    assert not np.array_equal(eng2.env_get()[0], run.pos)
    assert np.array_equal(eng2.distance(q)[0], o.distance(q)[0])
    v_e, f_e = eng2.motion_validity(q[:-1], q[1:], 10, 0.005)
    v_o, f_o = o.motion_validity(q[:-1], q[1:], 10, 0.005)
    assert np.array_equal(f_e, f_o)


def _parse_planner_log(text):
    """What the reference's visualizer does with a planner log (visualizer/log_parser.py:33-67), re-expressed: take everything after
    `Path:`, strip the `q: ` / `parent q: ` tags, stop at the dashes, keep the first configuration of every line."""
    import re
    block = re.findall(r"Path:\n(.*\n)+", text.replace("NONE", "None"), flags=re.DOTALL)[0]
    out = []
    for line in block.split("\n"):
        line = line.replace("parent q: ", "").replace("q: ", "")
        if not line:
            continue
        if line[0] == "-":
            break
        first = line.split("; ")[0].replace("(", "").replace(")", "")
        out.append([float(x) for x in first.split(" ") if x])
    return out


def test_planner_log_is_readable_by_the_visualizer_parser(tmp_path):
    from rpmplv2_b200 import planning
    rng = np.random.default_rng(9)
    path = rng.uniform(-3.1, 3.1, (7, 6)).astype(np.float32)
    res = dict(success=1, iterations=123, states=456, planning_time=0.03125, path=path)
    f = tmp_path / "plannerData.log"
    planning.write_planner_log(str(f), res, planning.RGBT_CONNECT, 6)
    text = f.read_text()
    assert text.startswith("Space Type:      RealVectorSpaceCUDA type\nDimensionality:  6\nPlanner type:    RGBT-Connect\nPlanner info:\n\t Succesfull:           yes\n")
    assert text.endswith("-" * 25 + "\n") and "parent q: NONE" in text.split("Path:\n")[1].split("\n")[0]
    got = np.asarray(_parse_planner_log(text), np.float32)
    assert got.shape == (7, 6) and np.allclose(got, path, rtol=2e-6, atol=1e-6)          # 6 significant digits, like Eigen's stream precision
    vis = os.path.join(os.path.dirname(REFERENCE_DATA), "visualizer", "log_parser.py")
    if os.path.isfile(vis):                                                              # the real parser, when the reference tree is here
        import importlib.util
        try:
            spec = importlib.util.spec_from_file_location("ref_log_parser", vis)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except ImportError:
            mod = None                                                                   # (it imports sre_parse, gone in recent Pythons)
        if mod is not None:
            ref = np.asarray(mod.LogParser(str(f)).get_path(), np.float32)
            assert ref.shape == (7, 6) and np.allclose(ref, path, rtol=2e-6, atol=1e-6)
    planning.write_planner_log(str(f), dict(success=0, iterations=5, states=2, planning_time=1.5, path=None), planning.RRT_CONNECT, 6, append=True)
    assert f.read_text().count("Planner type:") == 2 and "Succesfull:           no" in f.read_text()
