"""DRGBT's per-iteration geometry over the batched interface ("a DRGBT generateGBur driver").

One DRGBT iteration (/root/reference/src/planners/drbt/DRGBT.cpp:85-170) spends its Task-1 budget on three kinds of state-space calls:
    ss->computeDistance(q_current, true)                              DRGBT.cpp:101      (the environment moved: new distance + planes)
    extendGenSpine(q_current, q_horizon) for every horizon state      DRGBT.cpp:272-320 -> computeReachedState :475-480
    ss->computeDistance(q_reached, true) for every reached state      DRGBT.cpp:482
and then moves the robot and the obstacles together: MotionValidity::check(q_previous, q_current) (DRGBT.cpp:148, MotionValidity.cpp:30-51).
The reference makes 1 + H * (NUM_LAYERS underestimations) + H distance calls and K environment updates, one at a time.  Here that is
THREE engine calls whatever the horizon size H (root distance into a State slot, H generalized-bur spines from that cached root, H real
distances) plus ONE call for the motion check.  The planner's control logic (horizon bookkeeping, weights, replanning) stays on the
host: `horizon_weights` restates the weight heuristic of computeNextState (DRGBT.cpp:499-561) because the benches report it with the
geometry; horizon generation / modification / replanning are the reference's own code and are not rebuilt here.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .robots import DIST_PROFILE, RvsSpineParams

STATUS_GOOD, STATUS_BAD, STATUS_CRITICAL, STATUS_GOAL = 0, 1, 2, 3        # planning::drbt::HorizonState::Status
D_CRIT = 0.05                                                              # DRGBTConfig::D_CRIT
TRESHOLD_WEIGHT = 0.5                                                      # DRGBTConfig::TRESHOLD_WEIGHT


def default_spine_params() -> RvsSpineParams:
    """RGBTConnectConfig::NUM_LAYERS 5, RBTConnectConfig::NUM_ITER_SPINE 1, USE_EXPANDED_BUBBLE true, D_CRIT 0.05."""
    return RvsSpineParams(5, 1, 1, 0.05)


def generate_gbur(space, q_current, horizon, spine_params: Optional[RvsSpineParams] = None, root_slot: int = 0):
    """DRGBT::generateGBur without its state modification: computeReachedState for every horizon state.

    `space` is an engine (rpmplv2_b200.engine.Engine) or anything with the same store_distance / store_extend_spines / distance
    methods.  Returns a dict: d_c_current, profile_current, q_reached [H, n], is_reached [H] (spine status == Reached), d_c [H] (real
    distance of every reached state, as computeDistance(q_reached, true) gives it), spine_status [H], status [H] (HorizonState::Status
    after setDistance: Critical below D_CRIT, else Good)."""
    sp = spine_params or default_spine_params()
    q_current = np.ascontiguousarray(q_current, np.float32).reshape(1, -1)
    horizon = np.ascontiguousarray(horizon, np.float32).reshape(-1, q_current.shape[1])
    H = horizon.shape[0]
    space.store_reserve(max(root_slot + 1, 1))
    d_cur, prof_cur = space.store_distance(q_current, np.array([root_slot], np.int32))              # DRGBT.cpp:101
    if not d_cur[0] > 0:
        # extendSpine returns {Trapped, q} when the root has no clearance (RBTConnect.cpp:88-90): every spine stays at q_current
        return dict(d_c_current=float(d_cur[0]), profile_current=prof_cur[0], q_reached=np.repeat(q_current, H, axis=0),
                    is_reached=np.zeros(H, bool), d_c=np.zeros(H, np.float32), spine_status=np.full(H, 2, np.uint8),
                    status=np.full(H, STATUS_CRITICAL, np.int32))
    q_layers, n_layers, _, st = space.store_extend_spines(np.full(H, root_slot, np.int32), horizon, sp, gen2=False)      # extendGenSpine
    last = np.maximum(n_layers, 1) - 1
    q_reached = np.where((n_layers > 0)[:, None], q_layers[np.arange(H), last], q_current)
    d_c = space.distance(q_reached, DIST_PROFILE)[0]                                                # computeDistance(q_reached, true)
    status = np.where(d_c < np.float32(D_CRIT), STATUS_CRITICAL, STATUS_GOOD).astype(np.int32)       # HorizonState::setDistance
    return dict(d_c_current=float(d_cur[0]), profile_current=prof_cur[0], q_reached=q_reached.astype(np.float32), is_reached=st == 3,
                d_c=d_c, spine_status=st, status=status)


def horizon_weights(q_reached, d_c, d_c_previous, index, status, q_goal, d_max_mean: float, num_iterations: int,
                    equality_threshold: float = 1e-4):
    """The weight heuristic of DRGBT::computeNextState (DRGBT.cpp:499-561) in float32: returns (weights [H], idx_best or -1,
    new d_max_mean).  index[i] == -1 marks states that do not belong to the predefined path."""
    q_reached = np.asarray(q_reached, np.float32)
    f = np.float32
    dist_to_goal = np.sqrt(((q_reached - np.asarray(q_goal, np.float32)) ** 2).sum(axis=1, dtype=np.float32)).astype(np.float32)
    i_min = int(np.argmin(dist_to_goal))
    d_goal_min = dist_to_goal[i_min]
    if d_goal_min < f(equality_threshold):
        d_goal_min = f(equality_threshold)
        dist_to_goal[i_min] = d_goal_min
    weights_dist = (d_goal_min / dist_to_goal).astype(np.float32)
    d_c = np.asarray(d_c, np.float32)
    d_max_mean = f((f(num_iterations) * f(d_max_mean) + d_c.max()) / f(num_iterations + 1))
    mean = f(weights_dist.astype(np.float64).sum() / len(d_c))                         # std::accumulate(..., 0.0): double accumulator
    w = d_c / d_max_mean + (d_c - np.asarray(d_c_previous, np.float32)) / d_max_mean + weights_dist - mean
    w = np.clip(w, 0, 1).astype(np.float32)
    w = np.where(np.asarray(index) == -1, w * f(TRESHOLD_WEIGHT), w).astype(np.float32)
    w = np.where(np.asarray(status) == STATUS_CRITICAL, f(0), w)                        # critical states keep weight 0 and are skipped
    best = int(np.argmax(w)) if (w > 0).any() else -1
    return w, best, float(d_max_mean)
