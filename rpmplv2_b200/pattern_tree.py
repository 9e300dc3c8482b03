"""Pattern trees of generalized burs and the dataset rows built from them (SURVEY.md §8 rows f-1 "PatternTree layers" and f-3
"dataset CSV header/rows") over the batched interface.

planning::rbt::PatternTree::generateLocalTree (/root/reference/src/planners/rbt/PatternTree.cpp:62-91) grows, from a root, one
generalized-bur spine towards root +- DELTA along every coordinate axis (generateGBur, :20-60), then does the same from every new
node, num_layers times; a node with a parent skips the axis direction that points back at the parent.  The reference makes one
computeDistance and up to 2n extendGenSpine calls PER NODE; here a whole LAYER is two engine calls whatever its width: the real
distances of all its nodes into State slots (rvs_store_distance), then all their spines (rvs_store_extend_spines).
apps/generate_dataset.cpp:58-91,140-191 turns such trees into dataset.csv rows; `dataset_header` / `dataset_row` write the same text.

`space` is an engine (rpmplv2_b200.engine.Engine) or anything with store_reserve / store_distance / store_extend_spines.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .robots import Robot, RvsSpineParams

DELTA = np.float32(3.14159)                     # RBTConnectConfig::DELTA (src/configurations/RBTConnectConfig.cpp)


def num_nodes(n: int, num_layers: int) -> int:
    """PatternTree::getNumNodes (PatternTree.cpp:8-18)."""
    total = 1 + 2 * n
    for i in range(1, num_layers):
        total += 2 * n * (2 * n - 1) ** i
    return total


def prune_edge(q1, q2, lo, hi, eq_thr):
    """RealVectorSpace::pruneEdge (RealVectorSpace.cpp:118-163) in float32, same expression order."""
    f = np.float32
    q1, q2 = np.asarray(q1, f), np.asarray(q2, f)
    for k in range(len(q1)):
        if q2[k] > hi[k]:
            bound = hi[k]
        elif q2[k] < lo[k]:
            bound = lo[k]
        else:
            continue
        t = f((bound - q1[k]) / (q2[k] - q1[k]))
        cand = (q1 + t * (q2 - q1)).astype(f)
        if not ((cand < lo - f(eq_thr)) | (cand > hi + f(eq_thr))).any():
            return cand
    return q2.copy()


def _axis_targets(roots, parents, has_parent, lo, hi, delta, eq_thr):
    """The q_e of generateGBur for a whole layer (PatternTree.cpp:34-55), vectorised: for every root, axis order, sign -1 then +1, the
    direction that points back at the parent skipped, every target pruned to the joint limits.  roots / parents: [F, n].
    Returns (root_index [S], targets [S, n])."""
    f = np.float32
    F, n = roots.shape
    signs = np.tile(np.array([-1, 1], f), n)                                  # [2n]
    axes = np.repeat(np.arange(n), 2)                                         # [2n]
    q_e = np.repeat(roots[:, None, :], 2 * n, axis=1)                         # [F, 2n, n]
    q_e[:, np.arange(2 * n), axes] = (roots[:, axes] + signs[None, :] * f(delta)).astype(f)
    # the skip rule: |cos(q_e - root, parent - root) - 1| < threshold (NaN, i.e. a zero-length spine, never skips)
    a_i = (q_e[:, np.arange(2 * n), axes] - roots[:, axes]).astype(f)         # the only non-zero component of q_e - root
    b = (parents - roots).astype(f)                                           # [F, n]
    norm_a = np.sqrt((a_i * a_i).astype(f))
    norm_b = np.sqrt((b * b).sum(axis=1, dtype=f)).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        cos_phi = ((a_i * b[:, axes]).astype(f) / (norm_a * norm_b[:, None]).astype(f)).astype(f)
        skip = has_parent[:, None] & (np.abs(cos_phi - f(1)) < f(eq_thr))
    # pruneEdge: q_e leaves the limits along its own axis only, unless the root itself is outside (then the generic routine decides)
    qi = q_e[:, np.arange(2 * n), axes]
    over, under = qi > hi[axes][None, :], qi < lo[axes][None, :]
    bound = np.where(over, hi[axes][None, :], lo[axes][None, :]).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((bound - roots[:, axes]) / (qi - roots[:, axes])).astype(f)
        cand_i = (roots[:, axes] + t * a_i).astype(f)
    inside = (roots >= lo - f(eq_thr)) & (roots <= hi + f(eq_thr))            # the untouched coordinates of a candidate are the root's
    root_out = ((roots > hi) | (roots < lo))                                  # coordinates pruneEdge would also list
    ok_i = ~((cand_i < lo[axes][None, :] - f(eq_thr)) | (cand_i > hi[axes][None, :] + f(eq_thr)))
    others_inside = np.stack([np.delete(inside, i, axis=1).all(axis=1) for i in range(n)], axis=1)[:, axes] if n > 1 else np.ones((F, 2 * n), bool)
    pruned = (over | under) & ok_i & others_inside
    q_e[:, np.arange(2 * n), axes] = np.where(pruned, cand_i, qi)
    generic = root_out.any(axis=1)
    for fi in np.flatnonzero(generic):                                        # rare: a root outside its limits
        for j in range(2 * n):
            raw = roots[fi].copy()
            raw[axes[j]] = f(roots[fi, axes[j]] + signs[j] * f(delta))
            q_e[fi, j] = prune_edge(roots[fi], raw, lo, hi, eq_thr)
    keep = ~skip
    return np.nonzero(keep)[0].astype(np.int32), q_e[keep]


def generate_local_tree(space, robot: Robot, q_root, num_layers: int, spine_params: Optional[RvsSpineParams] = None, delta: float = DELTA,
                        equality_threshold: float = 1e-4):
    """PatternTree::generateLocalTree.  Returns a dict: coord [T, n] (node 0 = root, then layer by layer in the reference's order),
    parent [T] (-1 for the root), layer [T], d_c [T] / profile [T, n] (the REAL distance of every node that was expanded, NaN for the
    leaves of the last layer, which the reference never evaluates), status [T] (spine status that produced the node)."""
    sp = spine_params or RvsSpineParams(5, 5, 1, 0.03)              # RGBTConnectConfig / RBTConnectConfig defaults
    f = np.float32
    n = robot.num_dofs
    lo, hi = np.asarray(robot.limits[:, 0], f), np.asarray(robot.limits[:, 1], f)
    coord, parent, layer, status = [np.asarray(q_root, f).reshape(1, n).copy()], [np.array([-1], np.int32)], [np.zeros(1, np.int32)], [np.zeros(1, np.uint8)]
    d_c, profile = [np.full(1, np.nan, f)], [np.full((1, n), np.nan, f)]
    first, count = 0, 1                                               # the frontier = nodes [first, first + count) of the flat arrays
    all_coord = coord[0]
    all_parent = parent[0]
    for depth in range(num_layers):
        if count == 0:
            break
        # (1) the layer's real distances, nearest points kept on the device
        q_layer = all_coord[first:first + count]
        space.store_reserve(count)
        d, p = space.store_distance(q_layer, np.arange(count, dtype=np.int32))
        dl, pl = np.concatenate(d_c), np.concatenate(profile)
        dl[first:first + count], pl[first:first + count] = d, p
        d_c, profile = [dl], [pl]
        # (2) every spine of the layer
        par = all_parent[first:first + count]
        has_parent = par >= 0
        roots, targets = _axis_targets(q_layer, all_coord[np.maximum(par, 0)], has_parent, lo, hi, delta, equality_threshold)
        if len(roots) == 0:
            break
        q_lay, n_lay, _, st = space.store_extend_spines(roots, targets, sp, gen2=False)
        last = np.maximum(n_lay, 1) - 1
        q_new = np.where((n_lay > 0)[:, None], q_lay[np.arange(len(roots)), last], q_layer[roots]).astype(f)    # Trapped: the root itself
        S = len(roots)
        coord.append(q_new)
        parent.append((first + roots).astype(np.int32))
        layer.append(np.full(S, depth + 1, np.int32))
        status.append(st.astype(np.uint8))
        d_c.append(np.full(S, np.nan, f))
        profile.append(np.full((S, n), np.nan, f))
        all_coord, all_parent = np.concatenate(coord), np.concatenate(parent)
        coord, parent = [all_coord], [all_parent]
        first, count = first + count, S
    return dict(coord=all_coord, parent=all_parent, layer=np.concatenate(layer), d_c=np.concatenate(d_c), profile=np.concatenate(profile),
                status=np.concatenate(status))


# ------------------------------------------------------------------------------------------------- dataset.csv (generate_dataset.cpp)
def _num(x) -> str:
    """operator<<(std::ostream &, float) with the default precision (6 significant digits, %g)."""
    return "%g" % float(x)


def dataset_header(n: int, num_layers: int) -> str:
    """apps/generate_dataset.cpp:58-91."""
    cols = [f"root_node({j + 1})" for j in range(n)]
    cols += [f"spine_length{i}" for i in range(1, num_nodes(n, num_layers))]
    cols += [f"goal_node({j + 1})" for j in range(n)]
    for i in range(num_nodes(n, num_layers - 1)):
        cols += [f"dc{i}_link{j + 1}" for j in range(n)]
    cols += [f"next_vector({j + 1})" for j in range(n)]
    return ",".join(cols) + "\n"


def dataset_row(tree: dict, q_goal, q_next, num_layers: int) -> str:
    """One row of dataset.csv for the pattern tree of a path node (apps/generate_dataset.cpp:145-189): root, the length of every spine,
    the goal, the distance profile of every node but the last layer's, and the step to the next path node."""
    f = np.float32
    coord, parent = tree["coord"], tree["parent"]
    n = coord.shape[1]
    total, inner = num_nodes(n, num_layers), num_nodes(n, num_layers - 1)
    if len(coord) < total:
        raise ValueError(f"the pattern tree has {len(coord)} nodes, the dataset row needs {total} (the reference indexes past the tree in that case)")
    # a tree can hold MORE nodes than getNumNodes() (a zero-length spine leaves no direction to skip): the reference writes the first `total`
    vals = [_num(v) for v in coord[0]]
    diff = (coord[1:total] - coord[parent[1:total]]).astype(f)
    vals += [_num(np.sqrt((dv * dv).sum(dtype=f))) for dv in diff]
    vals += [_num(v) for v in np.asarray(q_goal, f)]
    for i in range(inner):
        vals += [_num(v) for v in tree["profile"][i]]
    vals += [_num(v) for v in (np.asarray(q_next, f) - coord[0]).astype(f)]
    return ",".join(vals) + "\n"
