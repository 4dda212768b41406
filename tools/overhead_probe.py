"""Scratch: fixed overhead of the primary pass vs frame size (C1 sphere scene), plus an all-gated frame."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
mesh = rtb200.Mesh().icosphere(6, 50.0).finish(); A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
ctx = rtb200.Context(0); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
def timeit(fn, iters=20):
    with torch.cuda.stream(stream):
        for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(); fn(); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3
for (w, h) in [(64, 64), (320, 240), (640, 480), (1280, 960), (1920, 1080), (3840, 2160)]:
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
    ctx.set_params(params)
    d = torch.zeros((w * h, 4), device="cuda")
    t = timeit(lambda: ctx.primary_device(w, h, d))
    nh = int((d.view(torch.int32)[:, 0] >= 0).sum())
    # all-gated variant: move the scene box far away in the params (aabb only gates, geometry unchanged)
    p2 = params.copy(); p2[24:27] += 1e4; p2[28:31] += 1e4
    ctx.set_params(p2)
    tg = timeit(lambda: ctx.primary_device(w, h, d))
    print(f"{w}x{h}: {t:.1f} us ({nh} hits) | all pixels gated out: {tg:.1f} us  ({w*h/32:.0f} tiles)")
# empty-kernel launch overhead reference: 1-ray trace
rays = torch.zeros((32, 8), device="cuda"); rays[:, 3] = 1e9; rays[:, 6] = 1.0; hits = torch.zeros((32, 4), device="cuda")
print("32-ray rt_trace_device:", timeit(lambda: ctx.trace_device(0, 32, rays, hits)), "us")
