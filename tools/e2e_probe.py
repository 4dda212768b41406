"""Scratch: e2e timing of rt_primary (host buffer) with zero-copy on/off, pinned vs pageable."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish(); A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0); ctx.upload_scene(A, bvh.nodes, bvh.tri_indices); ctx.set_params(params)
pinned = torch.empty((w * h, 4), dtype=torch.float32).pin_memory()
pageable = np.empty(w * h, dtype=rtb200.HIT_DTYPE)
d = torch.zeros((w * h, 4), device="cuda"); ctx.set_stream(None)
ctx.primary_device(w, h, d); ctx.synchronize(); ref = d.cpu()
for name, buf, zc in (("pinned zero-copy", pinned, 1), ("pinned staged", pinned, 0), ("pageable", pageable, 1)):
    ctx.set_option("zero_copy", zc)
    for _ in range(3): ctx.primary(w, h, buf)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.set_params(params); ctx.primary(w, h, buf)
    dt = (time.perf_counter() - t0) / 20
    got = buf if isinstance(buf, torch.Tensor) else torch.from_numpy(buf.view(np.float32).reshape(-1, 4))
    print(f"{name:18s} {dt*1e3:.3f} ms/step -> {752541/dt/1e6:.0f} Mrays/s  identical={torch.equal(got.view(torch.int32), ref.view(torch.int32))}")
