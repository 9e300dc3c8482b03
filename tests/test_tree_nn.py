"""Device-resident planner trees: rvs_tree_nearest against an exhaustive scan in nanoflann's metric
(base::Tree::getNearestState, src/state_spaces/Tree.cpp:35-50) and rvs_tree_radius against RRTx::findNeighbors' loop
(src/planners/rrtx/RRTx.cpp:424-452), both restated with numpy in the arithmetic order of the reference (bit-exact bar)."""
import numpy as np
import pytest

from rpmplv2_b200 import engine, robots, synthetic

pytestmark = pytest.mark.gpu


def dist2_sequential(q, pts):
    """nanoflann L2_Simple_Adaptor::evalMetric: sum of squared differences in index order, float32."""
    s = np.zeros(len(pts), np.float32)
    for d in range(pts.shape[1]):
        diff = (q[d] - pts[:, d]).astype(np.float32)
        s = (s + diff * diff).astype(np.float32)
    return s


def norm_eigen(q, pts):
    """(q1 - q2).norm() in Eigen's linear-vectorised order (oracle/rvs_oracle.hpp coord_norm), float32, vectorised over pts."""
    n = pts.shape[1]
    sq = ((q[None, :] - pts).astype(np.float32) ** 2).astype(np.float32)
    aligned, aligned2 = (n // 4) * 4, (n // 8) * 8
    if aligned:
        p0 = sq[:, 0:4].copy()
        if aligned > 4:
            p1 = sq[:, 4:8].copy()
            for k in range(8, aligned2, 8):
                p0 = (p0 + sq[:, k:k + 4]).astype(np.float32)
                p1 = (p1 + sq[:, k + 4:k + 8]).astype(np.float32)
            p0 = (p0 + p1).astype(np.float32)
            if aligned > aligned2:
                p0 = (p0 + sq[:, aligned2:aligned2 + 4]).astype(np.float32)
        res = ((p0[:, 0] + p0[:, 2]).astype(np.float32) + (p0[:, 1] + p0[:, 3]).astype(np.float32)).astype(np.float32)
        for i in range(aligned, n):
            res = (res + sq[:, i]).astype(np.float32)
    else:
        res = sq[:, 0].copy()
        for i in range(1, n):
            res = (res + sq[:, i]).astype(np.float32)
    return np.sqrt(res).astype(np.float32)


@pytest.mark.parametrize("rb", [robots.xarm6(), robots.planar(2), robots.spatial(22)], ids=lambda r: r.name)
def test_tree_nearest_is_the_exhaustive_minimum(rb):
    eng = engine.Engine(rb)
    t0, t1, t2 = eng.tree_create(), eng.tree_create(), eng.tree_create()
    pts0 = synthetic.configs(rb, 5000, seed=11)
    pts1 = synthetic.configs(rb, 37, seed=12)
    eng.tree_append(t0, pts0[:1000])                      # grows in pieces, across a reallocation
    eng.tree_append(t0, pts0[1000:])
    eng.tree_append(t1, pts1)
    assert eng.tree_size(t0) == 5000 and eng.tree_size(t1) == 37 and eng.tree_size(t2) == 0
    qs = synthetic.configs(rb, 300, seed=13)
    qs[7] = pts0[4321]                                    # an exact member: distance 0
    trees = np.array([t0, t1, t2] * 100, np.int32)
    idx, d2 = eng.tree_nearest(qs, trees)
    for k in range(300):
        pts = (pts0, pts1, pts0[:0])[k % 3]
        if len(pts) == 0:
            assert idx[k] == -1
            continue
        s = dist2_sequential(qs[k], pts)
        assert idx[k] == int(np.argmin(s)) and d2[k] == s.min()          # argmin = first minimum = lowest index on ties
    # duplicates: the lowest index wins
    eng.tree_clear(t1)
    eng.tree_append(t1, np.repeat(pts1[:3], 4, axis=0))
    idx, d2 = eng.tree_nearest(pts1[:3], np.full(3, t1, np.int32))
    assert list(idx) == [0, 4, 8] and (d2 == 0).all()
    # default tree 0
    idx0, _ = eng.tree_nearest(qs[:5])
    assert all(idx0[k] == int(np.argmin(dist2_sequential(qs[k], pts0))) for k in range(5))
    eng.close()


@pytest.mark.parametrize("rb", [robots.xarm6(), robots.spatial(10)], ids=lambda r: r.name)
def test_tree_radius_lists_the_neighbours_in_tree_order(rb):
    eng = engine.Engine(rb)
    t = eng.tree_create()
    rng = np.random.default_rng(5)
    centre = synthetic.configs(rb, 1, seed=3)[0]
    pts = (centre[None, :] + rng.uniform(-0.6, 0.6, (7000, eng.n))).astype(np.float32)
    eng.tree_append(t, pts)
    for radius, exclude in [(0.5, -1), (0.9, 123), (3.0, 6999), (1e-3, -1)]:
        q = pts[exclude] if exclude >= 0 else centre
        d = norm_eigen(q, pts)
        want = np.nonzero((d <= np.float32(radius)) & (np.arange(len(pts)) != exclude))[0]
        idx, dist = eng.tree_radius(t, q, radius, exclude, capacity=64)   # small capacity: the wrapper re-queries with the count
        assert (idx == want).all() and (dist == d[want]).all()
    eng.close()


def test_tree_errors():
    rb = robots.xarm6()
    eng = engine.Engine(rb)
    with pytest.raises(engine.RvsError):
        eng.tree_nearest(synthetic.configs(rb, 2))        # no tree yet
    t = eng.tree_create()
    with pytest.raises(engine.RvsError):
        eng.tree_append(t + 1, synthetic.configs(rb, 2))
    with pytest.raises(engine.RvsError):
        eng.tree_nearest(synthetic.configs(rb, 2), np.array([0, 5], np.int32))
    # a released id is handed out again, empty; using it in between is an error
    eng.tree_append(t, synthetic.configs(rb, 10))
    eng.tree_release(t)
    with pytest.raises(engine.RvsError):
        eng.tree_append(t, synthetic.configs(rb, 2))
    assert eng.tree_create() == t and eng.tree_size(t) == 0
    eng.close()
