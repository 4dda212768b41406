"""Scratch: full-frame renderers on the C2 scene: wavefront vs megakernel (bitwise), timings, reference OpenCL kernel."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
from oracle import oracle_py as O
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7)); A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
ctx = rtb200.Context(0); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
def timeit(fn, iters=10):
    with torch.cuda.stream(stream):
        for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(); fn(); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for light in ((-23.0, 200.0, 3.0), (-150.0, 25.0, 3.0)):
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=light)
    ctx.set_params(params)
    img0 = torch.zeros((h, w), dtype=torch.int32, device="cuda"); img1 = torch.zeros_like(img0)
    ctx.set_option("frame_mode", 0); ctx.render_frame_device(w, h, img0); t0 = timeit(lambda: ctx.render_frame_device(w, h, img0))
    res = []
    for lanes, late in ((0, 1), (1, 1), (1, 2), (1, 4)):
        ctx.set_option("frame_mode", 1); ctx.set_option("wf_lanes", lanes); ctx.set_option("wf_late_div", late)
        img1.zero_(); ctx.render_frame_device(w, h, img1); torch.cuda.synchronize()
        t1 = timeit(lambda: ctx.render_frame_device(w, h, img1))
        res.append(f"lanes={lanes},late_div={late}: {t1:.3f} ms identical={torch.equal(img0, img1)}")
    line = f"light {light}: megakernel {t0:.3f} ms | wavefront " + " | ".join(res)
    try:
        ref = O.RefCLScene(A, bvh.nodes, bvh.tri_indices)
        ms = [ref.render_frame(params, w, h)[1] for _ in range(5)]
        line += f" | reference OpenCL kernel on {ref.device_name()}: {np.median(ms):.3f} ms"
    except Exception as e:
        line += f" | reference OpenCL unavailable ({e})"
    print(line, flush=True)
