"""Pattern trees of generalized burs over the batched interface (rpmplv2_b200/pattern_tree.py) against
the reference's own PatternTree::generateLocalTree, compiled unmodified (tests/golden/make_tree_golden.py -> tree_golden.npz), and
the dataset.csv text of apps/generate_dataset.cpp.  CPU: the oracle-backed space; GPU: the CUDA engine (two calls per layer)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from rpmplv2_b200 import engine, pattern_tree, robots


def _golden():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    spec = importlib.util.spec_from_file_location("make_tree_golden", os.path.join(ROOT, "tests", "golden", "make_tree_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, dict(np.load(os.path.join(ROOT, "tests", "golden", "tree_golden.npz")))


def _scene(mod, name):
    fn, qs, qg = mod.mp.SCENARIOS[name]
    rb, mn, mx, g, tol = fn()
    return rb, (mn, mx, g, tol), np.asarray(qs, np.float32), np.asarray(qg, np.float32)


def _compare(tree, gold, key, layers, exact):
    G = lambda k: gold[f"{key}_{k}"]
    assert len(tree["coord"]) == len(G("coord")), key
    # a node without clearance returns ITSELF from extendGenSpine (Trapped) and is re-inserted into the reference's tree as its own
    # child: base::Tree::upgradeTree then overwrites its idx and parent, so its reported parent is a later copy of itself.  Compare
    # proper parents exactly and require the self-parented ones to be exactly such copies.
    proper = G("parent") < np.arange(len(G("parent")))
    assert np.array_equal(tree["parent"][proper], G("parent")[proper]), key
    for k in np.flatnonzero(~proper):
        assert np.array_equal(G("coord")[G("parent")[k]], G("coord")[k]) and G("d_c")[k] <= 0, (key, k)
    assert np.array_equal(tree["coord"], G("coord")) if exact else np.abs(tree["coord"] - G("coord")).max() <= 1e-5, key
    inner = tree["layer"] < layers                                   # nodes that were expanded: the reference evaluated their REAL distance
    assert G("is_real")[inner].all() and not np.isnan(tree["d_c"][inner]).any()
    if exact:
        assert np.array_equal(tree["d_c"][inner], G("d_c")[inner]) and np.array_equal(tree["profile"][inner], G("profile")[inner]), key
    else:
        assert np.abs(tree["d_c"][inner] - G("d_c")[inner]).max() <= 1e-4 and np.abs(tree["profile"][inner] - G("profile")[inner]).max() <= 1e-4, key


def _run(space_of, exact):
    mod, gold = _golden()
    for k, (name, layers, prm) in enumerate(mod.CASES):
        rb, scene, qs, qg = _scene(mod, name)
        space = space_of(rb, scene)
        for tag, q in (("start", qs), ("goal", qg)):
            tree = pattern_tree.generate_local_tree(space, rb, q, layers, robots.RvsSpineParams(*prm), delta=mod.DELTA)
            _compare(tree, gold, f"c{k}_{tag}", layers, exact)
            assert (tree["layer"][tree["parent"][1:]] + 1 == tree["layer"][1:]).all()


def test_pattern_tree_over_the_oracle_space_matches_the_reference_pattern_tree(ob):
    from oracle.oracle_space import OracleSpace

    def space_of(rb, scene):
        o = ob.Oracle(rb, robots.default_params())
        o.set_obstacles(*scene)
        return OracleSpace(o)
    _run(space_of, exact=True)


@pytest.mark.gpu
def test_pattern_tree_over_the_cuda_engine_matches_the_reference_pattern_tree():
    def space_of(rb, scene):
        eng = engine.Engine(rb)
        eng.set_obstacles(*scene)
        return eng
    _run(space_of, exact=False)


def test_dataset_header_and_row_follow_generate_dataset(ob):
    """apps/generate_dataset.cpp:58-91 (header) and :145-189 (row): column counts, order, and the numbers' text (operator<< of a float
    prints 6 significant digits)."""
    from oracle.oracle_space import OracleSpace
    mod, _ = _golden()
    rb, scene, qs, qg = _scene(mod, "planar2")
    o = ob.Oracle(rb, robots.default_params())
    o.set_obstacles(*scene)
    n, layers = 2, 3
    head = pattern_tree.dataset_header(n, layers).rstrip("\n").split(",")
    total, inner = pattern_tree.num_nodes(n, layers), pattern_tree.num_nodes(n, layers - 1)
    assert (total, inner) == (53, 17)                                 # PatternTree::getNumNodes: 1 + 4 + 4*3 + 4*9
    assert len(head) == n + (total - 1) + n + inner * n + n
    assert head[:3] == ["root_node(1)", "root_node(2)", "spine_length1"] and head[n + total - 1] == "goal_node(1)"
    assert head[n + total - 1 + n] == "dc0_link1" and head[-1] == "next_vector(2)"
    tree = pattern_tree.generate_local_tree(OracleSpace(o), rb, qs, layers)
    q_next = qs + np.array([0.1, 0.05], np.float32)
    row = pattern_tree.dataset_row(tree, qg, q_next, layers).rstrip("\n").split(",")
    assert len(row) == len(head)
    vals = np.array([float(v) for v in row])
    assert np.allclose(vals[:n], qs) and np.allclose(vals[n + total - 1:n + total - 1 + n], qg, atol=1e-5)
    lengths = np.linalg.norm(tree["coord"][1:total] - tree["coord"][tree["parent"][1:total]], axis=1)
    assert np.allclose(vals[n:n + total - 1], lengths, rtol=1e-5, atol=1e-6)
    assert np.allclose(vals[n + total - 1 + n:n + total - 1 + n + n], tree["profile"][0], rtol=1e-5)     # the root's distance profile
    assert np.allclose(vals[-n:], [0.1, 0.05], rtol=1e-4)
    assert row[0] == "-2" and row[1] == "-2.5"                        # "%g": no trailing zeros, like std::ostream
