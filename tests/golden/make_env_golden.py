#!/usr/bin/env python
"""tests/golden/make_env_golden.py — golden vectors for the obstacle motion laws, produced by the REFERENCE's own Environment.cpp /
Object.cpp / Box.cpp compiled here against oracle/eigen_shim + oracle/env_shim (oracle/Makefile target `ref` -> oracle/_ref/libenv_ref.so).
Run in the build container (needs /root/reference); writes tests/golden/env_golden.npz (committed, a few KB)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from rpmplv2_b200 import robots  # noqa: E402

LAWS = {"straight": 0, "circular": 1, "two_tunnels": 2, "light_directions": 4}
f32p, i32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)


def ref_update(lib, prm, law, sc, dt, steps, state=None):
    M = len(sc["max_vel"])
    pos, vel = (sc["pos"].copy(), sc["vel"].copy()) if state is None else (state[0].copy(), state[1].copy())
    sign = np.ones(M, np.int32) if state is None else state[2].copy()
    plen = np.zeros(M, np.float32) if state is None else state[3].copy()
    mn, mx = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    P = lambda a, t: a.ctypes.data_as(t)
    rc = lib.envref_update(P(prm, f32p), LAWS[law], M, P(sc["dim"], f32p), P(pos, f32p), P(vel, f32p), P(sc["max_vel"], f32p), P(sign, i32p), P(plen, f32p),
                           P(sc["is_dynamic"], u8p), ctypes.c_float(dt), steps, P(mn, f32p), P(mx, f32p))
    assert rc == 0
    return pos, vel, sign, plen, mn, mx


def cases():
    rb = robots.xarm6()
    prm = np.array(list(scenes.WS_CENTER) + [1.0, 0.3, 1.5], np.float32)
    out = []
    for law in LAWS:
        for seed, speed, dt, steps in ((11, 0.8, 0.01, 1), (12, 0.8, 0.004, 37), (13, 1.4, 0.02, 250), (14, 0.3, 0.05, 120)):
            out.append((law, prm, scenes.dynamic_scene(rb, 9, seed, speed=speed), np.float32(dt), steps))
    return out


# ---- random_directions (Environment.cpp:289-333): the whole run after srand(seed), every object after every update -------------------------
def random_cases():
    rb = robots.xarm6()
    out = []
    for seed, M, speed, acc_hi, ground_included, dt, steps in ((5, 9, 0.4, 1.0, 2, 0.05, 300), (6, 40, 0.6, 2.0, 2, 0.02, 200), (7, 9, 0.3, 0.0, 0, 0.05, 400),
                                                                (8, 70, 0.5, 0.5, 2, 0.01, 150)):
        prm = np.array(list(scenes.WS_CENTER) + [1.0, 0.15, 3.0], np.float32)
        sc = scenes.dynamic_scene(rb, M, seed, speed=speed, size=0.05)
        rng = np.random.default_rng(100 + seed)
        a = rng.normal(size=(M + 1, 3))
        a *= rng.uniform(0.2, 1.0, (M + 1, 1)) * acc_hi / np.linalg.norm(a, axis=1, keepdims=True)
        a[0] = 0
        a[1::4] = 0                                                        # like the shipped apps: addObject(object, vel) leaves the acceleration at zero
        out.append((prm, sc, a.astype(np.float32), ground_included, np.float32(dt), steps, 1000 + seed))
    return out


def ref_update_random(lib, prm, sc, acc, ground_included, dt, steps, seed):
    M = len(sc["max_vel"])
    pt, vt, at = [np.zeros((steps, M, 3), np.float32) for _ in range(3)]
    nxt = ctypes.c_int32()
    P = lambda a, t: np.ascontiguousarray(a).ctypes.data_as(t)
    rc = lib.envref_update_random(P(prm, f32p), int(ground_included), 3, M, P(sc["dim"], f32p), P(sc["pos"], f32p), P(sc["vel"], f32p), P(acc, f32p),
                                  P(sc["max_vel"], f32p), P(sc["is_dynamic"], u8p), ctypes.c_float(dt), steps, ctypes.c_uint32(seed), P(pt, f32p), P(vt, f32p),
                                  P(at, f32p), ctypes.byref(nxt))
    assert rc == 0
    return pt, vt, at, nxt.value


if __name__ == "__main__":
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libenv_ref.so"))
    data = {}
    for k, (law, prm, sc, dt, steps) in enumerate(cases()):
        pos, vel, sign, plen, mn, mx = ref_update(lib, prm, law, sc, float(dt), steps)
        data[f"c{k}_pos"], data[f"c{k}_vel"], data[f"c{k}_sign"], data[f"c{k}_plen"], data[f"c{k}_mn"], data[f"c{k}_mx"] = pos, vel, sign, plen, mn, mx
    for k, (prm, sc, acc, gi, dt, steps, seed) in enumerate(random_cases()):
        pt, vt, at, nxt = ref_update_random(lib, prm, sc, acc, gi, float(dt), steps, seed)
        moved = int((np.abs(np.diff(np.linalg.norm(vt, axis=2), axis=0)) > 1e-9).any(axis=0).sum())
        print("random_directions case", k, "objects", len(acc), "steps", steps, "objects whose speed changed", moved, "next rand()", nxt)
        # every 10th update (and the last) is kept: a divergence never heals, so this pins the whole run
        keep = sorted(set(range(9, steps, 10)) | {steps - 1})
        data[f"r{k}_pos"], data[f"r{k}_vel"], data[f"r{k}_acc"], data[f"r{k}_keep"], data[f"r{k}_next"] = pt[keep], vt[keep], at[keep], np.array(keep), np.array([nxt])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "env_golden.npz"), **data)
    print("wrote", len(cases()), "cases")
