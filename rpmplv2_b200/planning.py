"""ctypes binding of librvs_planning.so (include/rvs_planning.h): RRT-Connect / RBT-Connect / RGBT-Connect over the batched
state-space queries of the CUDA engine, P planner runs in lockstep. No CPU fallback: the library links librvs_cuda.so."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from .engine import Engine, RvsError, c_f32p
from .robots import RvsSpineParams

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librvs_planning.so")

RRT_CONNECT, RBT_CONNECT, RGBT_CONNECT = 0, 1, 2
KIND_NAMES = {RRT_CONNECT: "RRT-Connect", RBT_CONNECT: "RBT-Connect", RGBT_CONNECT: "RGBT-Connect"}


class PlanConfig(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("max_num_iter", ctypes.c_int64), ("max_num_states", ctypes.c_int64),
                ("max_planning_time", ctypes.c_float), ("eps_step", ctypes.c_float), ("max_extension_steps", ctypes.c_int32),
                ("delta", ctypes.c_float), ("num_spines", ctypes.c_int32), ("spine", RvsSpineParams)]


class PlanResult(ctypes.Structure):
    _fields_ = [("success", ctypes.c_int32), ("invalid_ends", ctypes.c_int32), ("iterations", ctypes.c_int64), ("states", ctypes.c_int64),
                ("planning_time", ctypes.c_float), ("path_len", ctypes.c_int32), ("rounds", ctypes.c_int64), ("num_valid", ctypes.c_int64),
                ("num_edges", ctypes.c_int64), ("num_distances", ctypes.c_int64), ("num_spines", ctypes.c_int64)]


class RrtxConfig(ctypes.Structure):
    _fields_ = [("eps_step", ctypes.c_float), ("r_rewire", ctypes.c_float), ("max_neighbors", ctypes.c_int32),
                ("one_call_per_iteration", ctypes.c_int32), ("device_index", ctypes.c_int32)]


class RrtxStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in ("states", "iterations", "trapped", "backend_calls", "edges_checked", "edges_reference", "reparented", "rewired", "reseated")]


c_i32p = ctypes.POINTER(ctypes.c_int32)

SYMBOLS = {
    "rvsplan_default_config": (None, [ctypes.c_int32, ctypes.POINTER(PlanConfig)]),
    "rvsplan_solve_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PlanConfig), c_f32p, c_f32p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                           ctypes.POINTER(PlanResult), c_f32p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]),
    "rvsplan_solve_batch_trees": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PlanConfig), c_f32p, c_f32p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                                 ctypes.POINTER(PlanResult), c_f32p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), c_f32p, c_i32p, c_i32p, ctypes.c_int32]),
    "rvsplan_last_error": (ctypes.c_char_p, []),
    "rvsplan_rrtx_default_config": (None, [ctypes.POINTER(RrtxConfig)]),
    "rvsplan_rrtx_grow": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(RrtxConfig), c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, c_f32p, c_i32p, c_f32p,
                                         c_i32p, ctypes.c_int32, c_i32p, c_i32p, c_i32p, ctypes.POINTER(RrtxStats)]),
}

_lib = None


def load_library(path: Optional[str] = None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def default_config(kind: int = RGBT_CONNECT, **overrides) -> PlanConfig:
    cfg = PlanConfig()
    load_library().rvsplan_default_config(kind, ctypes.byref(cfg))
    for k, v in overrides.items():
        if k in ("num_layers", "num_iter_spine", "use_expanded_bubble", "d_crit"):
            setattr(cfg.spine, k, v)
        else:
            setattr(cfg, k, v)
    return cfg


def unpack_results(results, paths, n):
    out = []
    for p, r in enumerate(results):
        d = {f: getattr(r, f) for f, _ in PlanResult._fields_}
        d["path"] = None if paths is None else paths[p, :min(r.path_len, paths.shape[1])].copy()
        out.append(d)
    return out


def attach_trees(results, tree_coord, tree_parent, tree_sizes):
    """results[p]['trees'] = [(coord [k, n], parent [k]) for the start tree and the goal tree]."""
    for p, r in enumerate(results):
        r["trees"] = []
        for t in range(2):
            k = min(int(tree_sizes[2 * p + t]), tree_coord.shape[1])
            r["trees"].append((tree_coord[2 * p + t, :k].copy(), tree_parent[2 * p + t, :k].copy()))
            r.setdefault("tree_sizes", []).append(int(tree_sizes[2 * p + t]))


def solve_batch(eng: Engine, cfg: PlanConfig, q_start, q_goal, seeds=None, max_path: int = 512, trees: int = 0):
    """P planner runs on the engine's resident scene. q_start / q_goal: [P, n] (or [n], broadcast over the seeds).
    Returns (list of result dicts incl. 'path' [len, n], pool stats dict).  trees = T > 0: every result also carries both trees
    (result['trees'], up to T states each) as RBTConnect::outputPlannerData dumps them."""
    lib = load_library()
    n = eng.n
    qs, qg = np.atleast_2d(np.asarray(q_start, np.float32)), np.atleast_2d(np.asarray(q_goal, np.float32))
    P = len(seeds) if seeds is not None else qs.shape[0]
    qs = np.ascontiguousarray(np.broadcast_to(qs, (P, n)))
    qg = np.ascontiguousarray(np.broadcast_to(qg, (P, n)))
    sd = np.ascontiguousarray(np.arange(P) if seeds is None else seeds, np.uint64)
    results = (PlanResult * P)()
    paths = np.zeros((P, max_path, n), np.float32)
    stats = (ctypes.c_int64 * 3)()
    head = (eng._h, ctypes.byref(cfg), qs.ctypes.data_as(c_f32p), qg.ctypes.data_as(c_f32p), sd.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), P, results,
            paths.ctypes.data_as(c_f32p), max_path, stats)
    if trees > 0:
        tc, tp, ts = np.zeros((2 * P, trees, n), np.float32), np.zeros((2 * P, trees), np.int32), np.zeros(2 * P, np.int32)
        rc = lib.rvsplan_solve_batch_trees(*head, tc.ctypes.data_as(c_f32p), tp.ctypes.data_as(c_i32p), ts.ctypes.data_as(c_i32p), trees)
    else:
        rc = lib.rvsplan_solve_batch(*head)
    if rc != 0:
        raise RvsError(rc, (lib.rvsplan_last_error() or b"").decode())
    out = unpack_results(results, paths, n)
    if trees > 0:
        attach_trees(out, tc, tp, ts)
    return out, {"rounds": int(stats[0]), "backend_calls": int(stats[1]), "kernel_launches": int(stats[2])}


def rrtx_config(**overrides) -> RrtxConfig:
    """configuration_rrtx.yaml's EPS_STEP / R_REWIRE / MAX_NEIGHBORS plus the two execution switches of rvsplan_rrtx_grow."""
    cfg = RrtxConfig()
    load_library().rvsplan_rrtx_default_config(ctypes.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def rrtx_call(fn, handle, cfg: RrtxConfig, n: int, q_root, samples, path_start: int = -1, max_states: Optional[int] = None):
    """Marshalling of rvsplan_rrtx_grow; `fn` / `handle` are parameters so that the tests can drive their CPU checker, which exports a
    function of the same signature, through the same code."""
    q_root = np.ascontiguousarray(q_root, np.float32).reshape(n)
    samples = np.ascontiguousarray(samples, np.float32).reshape(-1, n)
    S = samples.shape[0]
    cap = max_states or S + 1
    coord, parent, cost = np.zeros((cap, n), np.float32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    new_index, path = np.zeros(max(S, 1), np.int32), np.zeros(cap, np.int32)
    path_len, path_invalid, stats = ctypes.c_int32(0), ctypes.c_int32(0), RrtxStats()
    rc = fn(handle, ctypes.byref(cfg), q_root.ctypes.data_as(c_f32p), samples.ctypes.data_as(c_f32p), S, cap, coord.ctypes.data_as(c_f32p),
            parent.ctypes.data_as(c_i32p), cost.ctypes.data_as(c_f32p), new_index.ctypes.data_as(c_i32p), path_start, path.ctypes.data_as(c_i32p),
            ctypes.byref(path_len), ctypes.byref(path_invalid), ctypes.byref(stats))
    if rc != 0:
        return rc, None
    k = int(stats.states)
    return 0, dict(coord=coord[:k].copy(), parent=parent[:k].copy(), cost=cost[:k].copy(), new_index=new_index[:S].copy(), path=path[:path_len.value].copy(),
                   path_invalid=bool(path_invalid.value), stats={f: int(getattr(stats, f)) for f, _ in RrtxStats._fields_})


def rrtx_grow(eng: Engine, cfg: RrtxConfig, q_root, samples, path_start: int = -1, max_states: Optional[int] = None) -> dict:
    """RRTx's tree growth over the CUDA engine (include/rvs_planning.h: rvsplan_rrtx_grow): the tree rooted at q_root that RRTx::solve's
    loop builds from the random states `samples`; every candidate edge of an iteration is one batched engine call."""
    lib = load_library()
    rc, out = rrtx_call(lib.rvsplan_rrtx_grow, eng._h, cfg, eng.n, q_root, samples, path_start, max_states)
    if rc != 0:
        raise RvsError(rc, (lib.rvsplan_last_error() or b"").decode())
    return out


def _eigen_row(v) -> str:
    """Eigen's default IOFormat for `vec.transpose()`: 6 significant digits (StreamPrecision), coefficients separated by one space,
    every coefficient right-aligned to the widest one."""
    cells = [f"{float(x):.6g}" for x in v]
    w = max(len(c) for c in cells)
    return " ".join(c.rjust(w) for c in cells)


def write_planner_log(filename: str, result: dict, kind: int, num_dimensions: int, space_type: str = "RealVectorSpaceCUDA type",
                      append: bool = False) -> None:
    """The planner `.log` the reference's visualizer reads (RBTConnect::outputPlannerData, /root/reference/src/planners/rbt/RBTConnect.cpp:203-253;
    /root/reference/visualizer/log_parser.py:33-67 parses the `Path:` block): header, planner info, the path as
    `q: (...); parent q: (...)` lines (State's operator<<, src/state_spaces/State.cpp:58-65), the 25-dash terminator. When the result
    carries trees (solve_batch(..., trees=T)) the two dumps the reference writes between the info and the path are written too:
    `Tree: <name>` and one state line per node in tree order (base::Tree's operator<<, src/state_spaces/Tree.cpp:155-162; the trees are
    named "q_start" / "q_goal" by RRTConnect's constructor, RRTConnect.cpp:21-22)."""
    path = result.get("path")
    with open(filename, "a" if append else "w") as f:
        f.write(f"Space Type:      {space_type}\n")
        f.write(f"Dimensionality:  {num_dimensions}\n")
        f.write(f"Planner type:    {KIND_NAMES[kind]}\n")
        f.write("Planner info:\n")
        f.write(f"\t Succesfull:           {'yes' if result['success'] else 'no'}\n")
        f.write(f"\t Number of iterations: {int(result['iterations'])}\n")
        f.write(f"\t Number of states:     {int(result['states'])}\n")
        f.write(f"\t Planning time [s]:    {float(result['planning_time']):.6g}\n")
        for name, (coord, parent) in zip(("q_start", "q_goal"), result.get("trees") or []):
            f.write(f"Tree: {name}\n")
            for k in range(len(parent)):
                par = "NONE" if parent[k] < 0 else f"({_eigen_row(coord[parent[k]])})"
                f.write(f"q: ({_eigen_row(coord[k])}); parent q: {par}\n")
        if path is not None and len(path) > 0:
            f.write("Path:\n")
            # every state is printed with ITS parent (State's operator<<): the previous path state on the start-tree side, the next one
            # on the goal-tree side (RRTConnect::computePath walks the goal tree towards its root).  Without trees: the previous state.
            goal_side = {}
            if result.get("trees"):
                coord, par = result["trees"][1]
                goal_side = {coord[k].tobytes(): (None if par[k] < 0 else coord[par[k]]) for k in range(len(par))}
            on_goal_side = False
            for i, q in enumerate(path):
                key = np.ascontiguousarray(q, np.float32).tobytes()
                if not on_goal_side and key in goal_side and (i + 1 == len(path) or (goal_side[key] is not None and np.array_equal(goal_side[key], path[i + 1]))):
                    on_goal_side = True
                if on_goal_side:
                    parent = "NONE" if goal_side[key] is None else f"({_eigen_row(goal_side[key])})"
                else:
                    parent = "NONE" if i == 0 else f"({_eigen_row(path[i - 1])})"
                f.write(f"q: ({_eigen_row(q)}); parent q: {parent}\n")
        f.write("-" * 25 + "\n")
