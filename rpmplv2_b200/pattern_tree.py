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


def _axis_targets(root, parent, lo, hi, delta, eq_thr):
    """The q_e of generateGBur for one root (PatternTree.cpp:34-55): axis order, sign -1 then +1, the direction back to the parent skipped."""
    f = np.float32
    out = []
    for i in range(len(root)):
        for sign in (-1, 1):
            q_e = root.copy()
            q_e[i] = f(root[i] + f(sign) * f(delta))
            if parent is not None:
                a, b = (q_e - root).astype(f), (parent - root).astype(f)
                den = f(np.sqrt((a * a).sum(dtype=f)) * np.sqrt((b * b).sum(dtype=f)))
                with np.errstate(divide="ignore", invalid="ignore"):
                    cos_phi = f((a * b).sum(dtype=f) / den)
                if abs(cos_phi - f(1)) < f(eq_thr):
                    continue
            out.append(prune_edge(root, q_e, lo, hi, eq_thr))
    return out


def generate_local_tree(space, robot: Robot, q_root, num_layers: int, spine_params: Optional[RvsSpineParams] = None, delta: float = DELTA,
                        equality_threshold: float = 1e-4):
    """PatternTree::generateLocalTree.  Returns a dict: coord [T, n] (node 0 = root, then layer by layer in the reference's order),
    parent [T] (-1 for the root), layer [T], d_c [T] / profile [T, n] (the REAL distance of every node that was expanded, NaN for the
    leaves of the last layer, which the reference never evaluates), status [T] (spine status that produced the node)."""
    sp = spine_params or RvsSpineParams(5, 5, 1, 0.03)              # RGBTConnectConfig / RBTConnectConfig defaults
    f = np.float32
    n = robot.num_dofs
    lo, hi = np.asarray(robot.limits[:, 0], f), np.asarray(robot.limits[:, 1], f)
    coord, parent, layer, status = [np.asarray(q_root, f).copy()], [-1], [0], [0]
    d_c, profile = [np.nan], [np.full(n, np.nan, f)]
    frontier = [0]
    space.store_reserve(max(1, (2 * n) ** max(num_layers - 1, 0) + 1))
    for depth in range(num_layers):
        if not frontier:
            break
        # (1) the layer's real distances, nearest points kept on the device
        q_layer = np.stack([coord[k] for k in frontier])
        slots = np.arange(len(frontier), dtype=np.int32)
        space.store_reserve(len(frontier))
        d, p = space.store_distance(q_layer, slots)
        # (2) every spine of the layer
        roots, targets = [], []
        for s, k in enumerate(frontier):
            d_c[k], profile[k] = float(d[s]), p[s].copy()
            for q_e in _axis_targets(coord[k], coord[parent[k]] if parent[k] >= 0 else None, lo, hi, delta, equality_threshold):
                roots.append(s)
                targets.append(q_e)
        if not targets:
            break
        q_lay, n_lay, _, st = space.store_extend_spines(np.asarray(roots, np.int32), np.stack(targets), sp, gen2=False)
        nxt = []
        for j, s in enumerate(roots):
            k = frontier[s]
            q_new = q_lay[j, max(int(n_lay[j]), 1) - 1] if n_lay[j] > 0 else coord[k]      # Trapped: extendGenSpine returns the root itself
            coord.append(np.asarray(q_new, f).copy())
            parent.append(k)
            layer.append(depth + 1)
            status.append(int(st[j]))
            d_c.append(np.nan)
            profile.append(np.full(n, np.nan, f))
            nxt.append(len(coord) - 1)
        frontier = nxt
    return dict(coord=np.stack(coord), parent=np.asarray(parent, np.int32), layer=np.asarray(layer, np.int32), d_c=np.asarray(d_c, f),
                profile=np.stack(profile), status=np.asarray(status, np.uint8))


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
