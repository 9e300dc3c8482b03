"""Scenario / fixture I/O of RPMPLv2: the reference's own `data/**/scenario*.yaml` and
`random_scenarios.yaml` files load unchanged and become an engine (robot chain + resident obstacle table, or a dynamic environment).

Restates, for the fields the hot path consumes:
  scenario::Scenario::Scenario          /root/reference/src/scenario/Scenario.cpp:3-133   (robot type / urdf / capsules_radius /
                                        gripper_length / ground_included / q_start / q_goal / space)
  env::Environment::Environment         /root/reference/src/environments/Environment.cpp:13-109  (boxes: label, dim, pos, rot,
                                        min_dist_tol; "ground" dropped when ground_included == 0; WS_center / WS_radius; motion type)
  random_scenarios.yaml consumers       /root/reference/apps/test_drgbt_random_obstacles.cpp:124-151 (scenario_<n>/run_<k>: start, goal,
                                        object_<j>: pos, vel -> "dynamic_obstacle" boxes of random_obstacles.dim)
  apps' base radius / robot max vel     apps/test_drgbt_random_obstacles.cpp:115-116
`space` may be "RealVectorSpace" (what the files ship with) or "RealVectorSpaceCUDA" (this engine): both select this engine here;
"RealVectorSpaceFCL" (mesh-based) is out of scope and raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import yaml

from . import robots
from .engine import MOTION_TYPES, Engine, RvsEnvParams, env_params
from .robots import Robot

SUPPORTED_SPACES = ("RealVectorSpace", "RealVectorSpaceCUDA")


@dataclass
class Scenario:
    robot: Robot
    space: str
    q_start: np.ndarray
    q_goal: np.ndarray
    dim: np.ndarray                      # [M, 3] box sides
    pos: np.ndarray                      # [M, 3]
    rotated: np.ndarray                  # [M] uint8: non-identity `rot` (FCL then reports the bounding-sphere cube)
    labels: List[str]
    min_dist_tol: np.ndarray             # [M]
    vel: np.ndarray                      # [M, 3]
    max_vel: np.ndarray                  # [M]
    ws_center: tuple = (0.0, 0.0, 0.0)
    ws_radius: float = 0.0
    motion_type: Optional[str] = None    # testing.motion_type
    robot_max_vel: Optional[np.ndarray] = None
    random_obstacles: Dict = field(default_factory=dict)
    testing: Dict = field(default_factory=dict)

    @property
    def num_objects(self) -> int:
        return len(self.labels)

    @property
    def is_ground(self) -> np.ndarray:
        return np.array([1 if lab == "ground" else 0 for lab in self.labels], np.uint8)

    @property
    def is_dynamic(self) -> np.ndarray:
        return np.array([1 if lab == "dynamic_obstacle" else 0 for lab in self.labels], np.uint8)

    def aabbs(self):
        """(aabb_min, aabb_max) as FCL's computeAABB reports them (Box.cpp:3-8): pos -/+ dim/2, or the bounding-sphere cube."""
        half = (np.float32(0.5) * self.dim).astype(np.float32)
        rad = np.sqrt(((half[:, 0] * half[:, 0] + half[:, 1] * half[:, 1]) + half[:, 2] * half[:, 2]).astype(np.float32)).astype(np.float32)
        half = np.where(self.rotated.astype(bool)[:, None], rad[:, None], half).astype(np.float32)
        return (self.pos + (-half)).astype(np.float32), (self.pos + half).astype(np.float32)

    def base_radius(self) -> float:
        """env->setBaseRadius(max(r_0, r_1) + |obs_dim|) (apps/test_drgbt_random_obstacles.cpp:115)."""
        d = np.asarray(self.random_obstacles.get("dim", [0.0, 0.0, 0.0]), np.float32)
        r = self.robot.radius
        return float(max(r[0], r[1] if len(r) > 1 else r[0]) + np.sqrt((d * d).sum(dtype=np.float32)))

    def env_params(self) -> RvsEnvParams:
        if self.motion_type is None:
            raise ValueError("the scenario has no testing.motion_type")
        vmax = float(self.robot_max_vel[0]) if self.robot_max_vel is not None and len(self.robot_max_vel) else 1.0     # "only velocity of the first joint matters" (:116)
        return env_params(self.motion_type, ws_center=self.ws_center, ws_radius=self.ws_radius, base_radius=self.base_radius(), robot_max_vel=vmax)

    def add_objects(self, dim, pos, vel=None, max_vel=0.0, label="dynamic_obstacle", min_dist_tol=None):
        """Environment::addObject for K boxes (identity rotation) of side `dim`."""
        pos = np.asarray(pos, np.float32).reshape(-1, 3)
        K = len(pos)
        self.dim = np.vstack([self.dim, np.tile(np.asarray(dim, np.float32).reshape(1, 3), (K, 1))]).astype(np.float32)
        self.pos = np.vstack([self.pos, pos]).astype(np.float32)
        self.vel = np.vstack([self.vel, np.zeros((K, 3), np.float32) if vel is None else np.asarray(vel, np.float32).reshape(K, 3)]).astype(np.float32)
        self.max_vel = np.concatenate([self.max_vel, np.full(K, max_vel, np.float32)])
        self.rotated = np.concatenate([self.rotated, np.zeros(K, np.uint8)])
        self.labels += [label] * K
        tol = robots.default_params().max_distance if min_dist_tol is None else min_dist_tol
        self.min_dist_tol = np.concatenate([self.min_dist_tol, np.full(K, tol, np.float32)])

    def make_engine(self, device: int = 0, params=None, dynamic: Optional[bool] = None) -> Engine:
        """The engine for this scenario. dynamic=True (default when the scenario names a motion type and holds dynamic obstacles)
        hands the objects to rvs_env_set, otherwise the AABBs go to rvs_set_obstacles."""
        eng = Engine(self.robot, params or robots.default_params(), device=device)
        if dynamic is None:
            dynamic = self.motion_type is not None and bool(self.is_dynamic.any())
        if dynamic:
            eng.env_set(self.env_params(), self.dim, self.pos, vel=self.vel, max_vel=self.max_vel, is_dynamic=self.is_dynamic,
                        is_ground=self.is_ground, min_dist_tol=self.min_dist_tol, rotated=self.rotated)
        else:
            mn, mx = self.aabbs()
            eng.set_obstacles(mn, mx, self.is_ground, self.min_dist_tol)
        return eng


def _load_yaml(path: str):
    # the reference's files separate values and comments with TABs, which yaml-cpp accepts and PyYAML rejects
    with open(path) as f:
        return yaml.load(f.read().replace("\t", "    "), Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))


def _floats(node, n=None, what=""):
    a = np.asarray([float(x) for x in node], np.float32)
    if n is not None and len(a) != n:
        raise ValueError(f"{what}: expected {n} values, got {len(a)}")
    return a


def load(config_file_path: str, root_path: str = "", urdf: str = "auto") -> Scenario:
    """scenario::Scenario(config_file_path, root_path). urdf = "auto": parse root_path + robot.urdf when that file exists, else use the
    built-in chain table of that robot type (tests/test_host_logic.py checks the tables against the reference's URDFs);
    "file" / "builtin" force one of the two."""
    node = _load_yaml(root_path + config_file_path)
    rn = node["robot"]
    rtype, n = str(rn["type"]), int(rn["num_DOFs"])
    space = str(rn.get("space", "RealVectorSpace"))
    if space not in SUPPORTED_SPACES:
        raise ValueError(f"state space {space!r} is not served by this engine (supported: {SUPPORTED_SPACES})")
    urdf_path = root_path + str(rn.get("urdf", ""))
    use_file = urdf == "file" or (urdf == "auto" and os.path.isfile(urdf_path))
    rb = robots.from_urdf(urdf_path, rtype) if use_file else robots.by_type(rtype)        # unknown type: raises like Scenario.cpp:34
    if rb.num_dofs != n:
        raise ValueError("num_DOFs does not match the robot chain")
    if "ground_included" in rn:
        rb.ground_included = int(rn["ground_included"])
    elif rtype == "xarm6":
        rb.ground_included = 0                                                              # Environment.cpp:66-70
    if "gripper_length" in rn:
        rb.gripper_length = float(rn["gripper_length"])
    if "capsules_radius" in rn:                                                             # Scenario.cpp:36-51
        cr = _floats(rn["capsules_radius"])
        if len(cr) == 1:
            cr = np.full(n, cr[0], np.float32)
        elif len(cr) != n:
            raise ValueError("Number of capsules is not correct!")
        rb.radius = cr.astype(np.float32)
    q_start, q_goal = _floats(rn["q_start"], n, "q_start"), _floats(rn["q_goal"], n, "q_goal")          # Scenario.cpp:114-127

    dims, poss, rots, labels, tols = [], [], [], [], []
    for item in node.get("environment") or []:                                              # Environment.cpp:22-62
        if "box" not in item:
            raise ValueError("Object type is wrong!")                                       # spheres are a TODO stub in the reference
        b = item["box"]
        label = str(b.get("label", ""))
        if label == "ground" and rb.ground_included == 0:
            continue
        dims.append(_floats(b["dim"], 3, "dim"))
        poss.append(_floats(b["pos"], 3, "pos"))
        r = b.get("rot")
        rots.append(0 if r is None or (float(r[0]) == 0.0 and float(r[1]) == 0.0 and float(r[2]) == 0.0) else 1)
        labels.append(label)
        tols.append(float(b["min_dist_tol"]) if "min_dist_tol" in b else robots.default_params().max_distance)
    M = len(labels)
    testing = dict(node.get("testing") or {})
    sc = Scenario(robot=rb, space=space, q_start=q_start, q_goal=q_goal,
                  dim=np.asarray(dims, np.float32).reshape(M, 3), pos=np.asarray(poss, np.float32).reshape(M, 3),
                  rotated=np.asarray(rots, np.uint8), labels=labels, min_dist_tol=np.asarray(tols, np.float32),
                  vel=np.zeros((M, 3), np.float32), max_vel=np.zeros(M, np.float32),
                  ws_center=tuple(float(x) for x in rn.get("WS_center", (0.0, 0.0, 0.0))), ws_radius=float(rn.get("WS_radius", 0.0)),
                  motion_type=str(testing["motion_type"]) if "motion_type" in testing else None,
                  robot_max_vel=_floats(rn["max_vel"]) if "max_vel" in rn else None,
                  random_obstacles=dict(node.get("random_obstacles") or {}), testing=testing)
    if sc.motion_type is not None and sc.motion_type not in MOTION_TYPES:
        sc.motion_type = "light_directions"                                                 # Environment.cpp:98-102
    return sc


def load_random_scenarios(path: str, only: Optional[int] = None) -> dict:
    """random_scenarios.yaml -> {num_obstacles: [ {start, goal, pos[K,3], vel[K,3]} per run ]} (written by
    apps/generate_random_scenarios.cpp, read by apps/test_drgbt_random_obstacles.cpp:124-151). The shipped files are ~20 MB:
    `only` = k parses just the block `scenario_<k>`."""
    if only is None:
        node = _load_yaml(path)
    else:
        lines, keep = [], False
        with open(path) as f:
            for line in f:
                if line.startswith("scenario_"):
                    keep = line.split(":")[0].strip() == f"scenario_{only}"
                if keep:
                    lines.append(line.replace("\t", "    "))
        node = yaml.load("".join(lines), Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader)) or {}
    out = {}
    for skey, runs in node.items():
        k = int(str(skey).split("_")[1])
        rows = []
        for r in range(len(runs)):
            run = runs[f"run_{r}"]
            pos = np.asarray([run[f"object_{j}"]["pos"] for j in range(k)], np.float32).reshape(k, 3)
            vel = np.asarray([run[f"object_{j}"]["vel"] for j in range(k)], np.float32).reshape(k, 3)
            rows.append(dict(start=_floats(run["start"]), goal=_floats(run["goal"]), pos=pos, vel=vel))
        out[k] = rows
    return out


def apply_random_run(sc: Scenario, run: dict) -> Scenario:
    """One run of a random scenario: start / goal and the dynamic obstacles of that run (test_drgbt_random_obstacles.cpp:128-151)."""
    import copy
    s = copy.deepcopy(sc)
    s.q_start, s.q_goal = run["start"].copy(), run["goal"].copy()
    ro = sc.random_obstacles
    s.add_objects(ro.get("dim", [0.01, 0.01, 0.01]), run["pos"], run["vel"], max_vel=float(ro.get("max_vel", 0.0)), label="dynamic_obstacle")
    return s
