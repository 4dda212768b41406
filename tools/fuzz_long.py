"""Extended fuzz of the CUDA path against the CPU oracle: the generator of tests/test_gpu_parity.py::test_random_triangle_soup_fuzz
over many seeds and scales (incl. scales that push coordinates towards the edges of the hoisted-division window), both
schedulers, closest and any hit, plus frames in both frame modes. usage: python tools/fuzz_long.py [seeds=40]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
from oracle import oracle_py as O

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = rtb200.Context(0)
scales = [1.0, 1e-2, 1e3, 37.0, 0.3, 1e-6, 1e6, 1e12, 1e-12, 7e17, 3e-20]
bad = total = 0
for seed in range(100, 100 + nseeds):
    rng = np.random.default_rng(seed)
    scale = scales[seed % len(scales)]
    nt = int(rng.integers(50, 2500))
    v = (rng.normal(size=(nt, 3, 3)) * 10).astype(np.float32)
    v[:, 1] = v[:, 0] + (rng.normal(size=(nt, 3)) * 1.5).astype(np.float32)
    v[:, 2] = v[:, 0] + (rng.normal(size=(nt, 3)) * 1.5).astype(np.float32)
    q = max(1, nt // 40)
    v[:q, 2] = v[:q, 1]
    v[q:2 * q, 1] = v[q:2 * q, 0]
    v[q:2 * q, 2] = v[q:2 * q, 0]
    v[2 * q:5 * q, :, 1] = np.round(v[2 * q:5 * q, :1, 1])
    v[5 * q:6 * q, :, 0] = 0.0
    v *= np.float32(scale)
    verts = np.concatenate([v.reshape(-1, 3), np.ones((nt * 3, 1), dtype=np.float32)], axis=1)
    m = rtb200.Mesh().set(verts, np.arange(nt * 3, dtype=np.int32)).finish()
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    info = ctx.scene_info()
    sc = O.OracleScene(A, b.nodes, b.tri_indices)
    n = 20000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = (rng.normal(size=(n, 3)) * 12 * scale).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[:, 3] = rtb200.T_INIT
    k = n // 8
    rays[:k, 4:7] = 0
    rays[:k, 4 + (np.arange(k) % 3)] = np.where(np.arange(k) % 2, 1.0, -1.0)
    rays[k:2 * k, 5] = 0.0
    rays[2 * k:3 * k, 4] *= np.float32(1e-7)   # direction components below the window
    rays[3 * k:4 * k, 0:3] = np.round(rays[3 * k:4 * k, 0:3] / np.float32(scale)) * np.float32(scale)
    rays[4 * k:5 * k, 3] = (rng.random(k) * 20 * scale + 1e-3 * scale).astype(np.float32)
    rays[5 * k:6 * k, 0:3] = 0.0              # origins at the coordinate origin
    for mode in (rtb200.CLOSEST, rtb200.ANY):
        want, _ = sc.trace(mode, rays)
        for sched in (0, 1):
            ctx.set_option("scheduler", sched)
            got = ctx.trace(mode, rays)
            same = (got["idx"] == want["idx"]) & (got["t"].view(np.uint32) == want["t"].view(np.uint32)) & \
                   (got["u"].view(np.uint32) == want["u"].view(np.uint32)) & (got["v"].view(np.uint32) == want["v"].view(np.uint32))
            total += n
            if not same.all():
                bad += int((~same).sum())
                i = int(np.flatnonzero(~same)[0])
                print(f"MISMATCH seed {seed} scale {scale} mode {mode} sched {sched}: ray {i} {rays[i]} got {got[i]} want {want[i]}")
    ctx.set_option("scheduler", -1)
    # frames in both modes against the oracle
    w, h = 64, 48
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], d_radius=float(np.linalg.norm(A["aabb_max"] - A["aabb_min"])) - 200.0)
    ctx.set_params(params)
    ref, _ = sc.render_frame(params, w, h)
    for fm in (0, 4):  # directly stored and row-assembled frames
        ctx.set_option("store_group", fm)
        img = ctx.render_frame(w, h)
        df = np.abs(img.view(np.uint8).astype(int) - ref.view(np.uint8).astype(int)).max()
        if df > 1:
            bad += 1
            print(f"FRAME seed {seed} scale {scale} mode {fm}: max diff {df}")
    ctx.set_option("store_group", -1)
    print(f"seed {seed} scale {scale:g} tris {nt} hoisted={info['hoisted_division']} hit {float((want['idx'] >= 0).mean()):.3f} ok", flush=True)
print(f"fuzz done: {total} ray comparisons, {bad} mismatches")
sys.exit(1 if bad else 0)
