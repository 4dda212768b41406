"""Randomised check of the culled tile queue (cull_setup) and the tile hints against a context with both switched off:
random orbit poses (far, close, inside the scene box, looking away), random frame sizes and band partitions, three scenes.
usage: python tools/cull_fuzz.py [poses=300] -> prints mismatches (must be 0)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from conftest import load_scene, mesh_dict

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(7)
a, b = rtb200.Context(0), rtb200.Context(0)
b.set_option("gate_cull", 0)
b.set_option("tile_hints", 0)
scenes = []
m = rtb200.Mesh().terrain(96, 100.0).icosphere(3, 22.0, (15.0, 28.0, -20.0)).finish(diffuse=(0.6, 0.7, 0.8))
scenes.append((m.arrays(), rtb200.FlatBVH.build(m)))
m2 = rtb200.Mesh().icosphere(4, 50.0).finish(diffuse=(0.8, 0.3, 0.2))
scenes.append((m2.arrays(), rtb200.FlatBVH.build(m2)))
bad = checked = culled = 0
for s, (A, bvh) in enumerate(scenes):
    for c in (a, b):
        c.upload_scene(A, bvh.nodes, bvh.tri_indices)
    for p in range(n_poses):
        w, h = int(rng.integers(5, 80)) * 8, int(rng.integers(5, 60)) * 4
        if p % 7 == 0:
            w += int(rng.integers(1, 8))  # ragged width
        pose = dict(d_radius=float(rng.uniform(-199.0, 600.0)), d_alpha=float(rng.uniform(-3.2, 3.2)), d_beta=float(rng.uniform(-0.75, 0.75)))
        lo, hi = A["aabb_min"].copy(), A["aabb_max"].copy()
        if p % 5 == 0:  # a scene box that is not the mesh's own (the gate uses the Params block's box)
            lo = lo + rng.uniform(0, 20, 3).astype(np.float32)
            hi = hi - rng.uniform(0, 20, 3).astype(np.float32)
        params, _ = rtb200.camera_params(w, h, lo, hi, light_pos=(-150.0, 25.0, 3.0), **pose)
        if p % 11 == 0:  # look past the scene: shift the image plane sideways
            params = params.copy()
            params[8:11] += params[0:3] * float(rng.uniform(-3, 3))
        for c in (a, b):
            c.set_params(params)
        for _ in range(2):  # second launch runs with the first one's hints
            f1, f2 = a.render_frame(w, h), b.render_frame(w, h)
            v1, v2 = a.primary_shadow(w, h), b.primary_shadow(w, h)
            checked += 2
            if not np.array_equal(f1, f2) or not np.array_equal(v1, v2):
                bad += 1
                print("MISMATCH scene", s, "pose", p, pose, (w, h), int((f1 != f2).sum()), int((v1 != v2).sum()), flush=True)
print(f"cull_fuzz: {checked} frame pairs over {len(scenes)} scenes x {n_poses} poses, mismatches: {bad}")
sys.exit(1 if bad else 0)
