"""Generates tests/golden/cad_golden.npz from the REFERENCE's own CollisionAndDistance.cpp.

Run in the build container only (needs /root/reference):   make -C oracle ref && python tests/golden/make_golden.py
The reference source is compiled where it lies against oracle/eigen_shim (see oracle/Makefile); its outputs on the
seeded inputs below are the golden vectors the oracle, the host-compiled device headers and the GPU kernels are pinned to.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

f32p, u8p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32)


def P(a, t=f32p):
    return a.ctypes.data_as(t)


def gen_cases(rng, K, box_scale, snap):
    c = rng.uniform(-1, 1, (K, 3))
    h = rng.uniform(box_scale[0], box_scale[1], (K, 3)) / 2
    obs = np.concatenate([c - h, c + h], 1).astype(np.float32)
    A = rng.uniform(-1.5, 1.5, (K, 3))
    L = rng.uniform(0.01, 1.0, (K, 1))
    d = rng.normal(size=(K, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    B = A + L * d
    if snap:       # axis-aligned segments, endpoints on face planes: the measure-zero branches
        B = np.where(rng.random((K, 3)) < 0.3, A, B)
        A = np.where(rng.random((K, 3)) < 0.15, obs[:, :3], A)
        B = np.where(rng.random((K, 3)) < 0.15, obs[:, 3:], B)
        same = np.all(A == B, axis=1)
        B[same, 0] += 0.25            # no zero-length capsules (the reference returns NaN there)
    AB = np.concatenate([A, B], 1).astype(np.float32)
    r = rng.choice([0.0, 0.025, 0.047, 0.05, 0.09, 0.11, 0.12], K).astype(np.float32)
    return AB, r, obs


def main():
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcad_ref.so"))
    rng = np.random.default_rng(20261017)
    out = {}
    for thr_name, thr in (("t4", 1e-4), ("t6", 1e-6)):
        ref.ref_set_config(ctypes.c_float(thr), ctypes.c_float(1000.0))
        parts = [gen_cases(rng, 700, sb, snap) for snap in (False, True) for sb in ((0.01, 0.011), (0.05, 0.3), (0.3, 2.0))]
        AB, r, obs = (np.concatenate([p[k] for p in parts]) for k in range(3))
        K = AB.shape[0]
        dist = np.empty((K, 8), np.float32)
        ref.ref_capsule_box_distance_f32(ctypes.c_int64(K), P(AB), P(r), P(obs), P(dist))
        coll = np.empty(K, np.uint8)
        ref.ref_capsule_box_collision_f32(ctypes.c_int64(K), P(AB), P(r), P(obs), P(coll, u8p))
        face = rng.integers(0, 6, K).astype(np.int32)
        rect = np.empty(K, np.uint8)
        ref.ref_capsule_rect_collision_f32(ctypes.c_int64(K), P(AB), P(r), P(obs), P(face, i32p), P(rect, u8p))
        out.update({f"{thr_name}_AB": AB, f"{thr_name}_r": r, f"{thr_name}_obs": obs, f"{thr_name}_dist": dist,
                    f"{thr_name}_coll": coll, f"{thr_name}_face": face, f"{thr_name}_rect": rect})
    ref.ref_set_config(ctypes.c_float(1e-4), ctypes.c_float(1000.0))
    K = 3000
    ABCD = rng.uniform(-1, 1, (K, 12)).astype(np.float32)
    m = rng.random(K) < 0.3     # parallel segments: the 0/0 path (Q8)
    ABCD[m, 9:12] = ABCD[m, 6:9] + (ABCD[m, 3:6] - ABCD[m, 0:3]) * rng.uniform(-2, 2, (m.sum(), 1)).astype(np.float32)
    m = rng.random(K) < 0.1     # shared endpoint
    ABCD[m, 6:9] = ABCD[m, 0:3]
    ss = np.empty((K, 8), np.float32)
    ref.ref_seg_seg_f32(ctypes.c_int64(K), P(ABCD), P(ss))
    ABC = rng.uniform(-1, 1, (1000, 9)).astype(np.float32)
    sp = np.empty((1000, 8), np.float32)
    ref.ref_seg_point_f32(ctypes.c_int64(1000), P(ABC), P(sp))
    out.update(dict(ss_ABCD=ABCD, ss_out=ss, sp_ABC=ABC, sp_out=sp))
    np.savez_compressed(os.path.join(HERE, "cad_golden.npz"), **out)
    print("wrote cad_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
