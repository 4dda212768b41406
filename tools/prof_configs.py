"""ncu / timing target for the BASELINE configurations that had no counters in round 1 (VERDICT r1 task 8).
usage: python tools/prof_configs.py {c1|c2|c4} {primary|shadow|fused|diffuse|frame} [scheduler=-1] [launches=6] [key=value options ...]
Prints the median device time of the launches (CUDA events); under ncu use `-k regex:<kernel> -s <launches-2> -c 1`.
The C4 scene's flat BVH is cached in gpurun_out/ (13 s build) so that the plain run ncu requires first pays it once."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200

scene, what = sys.argv[1], sys.argv[2]
sched = int(sys.argv[3]) if len(sys.argv) > 3 else -1
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 6
opts = dict(kv.split("=") for kv in sys.argv[5:])
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
t0 = time.time()
d_radius = 0.0
if scene == "c1":
    w, h = 640, 480
    src = rtb200.Mesh().icosphere(6, 50.0).finish(diffuse=(0.8, 0.3, 0.2))
    path = "/tmp/c1.dae"
    src.write_dae(path)
    mesh = rtb200.Mesh().load_dae(path)
elif scene == "c2":
    w, h = 1920, 1080
    mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
else:
    w, h = 1920, 1080
    mesh = rtb200.Mesh().sphere_field(11, 120.0, 6, 50.0).finish(diffuse=(0.7, 0.7, 0.7))
    d_radius = 1320.0
A = mesh.arrays()
cache = f"/tmp/rtb200_prof_{scene}.fbvh"  # on the GPU box, NOT under gpurun_out/ (64 MiB limit for what travels back)
if os.path.exists(cache):
    bvh = rtb200.FlatBVH.load(cache)
else:
    bvh = rtb200.FlatBVH.build(mesh)
    bvh.save(cache)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], d_radius=d_radius)
ctx = rtb200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
if "top_pairs" in opts:
    ctx.set_option("top_pairs", int(opts.pop("top_pairs")))
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
ctx.set_params(params)
print(f"scene {scene}: {A['indices'].size // 3} triangles, ready in {time.time() - t0:.1f} s", flush=True)
n = w * h
with torch.cuda.stream(stream):
    d_hits = torch.zeros((n, 4), device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    d_sh = torch.zeros((n, 4), device="cuda")
    d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ctx.set_option("scheduler", 0)
    ctx.primary_device(w, h, d_hits, d_rays)
    if what == "diffuse":
        d_dr = torch.zeros((n * 4, 8), device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
        torch.cuda.synchronize()
        nd = int(d_cnt.item())
        d_dh = torch.zeros((nd, 4), device="cuda")
torch.cuda.synchronize()
ctx.set_option("scheduler", sched)
for k, v in opts.items():
    ctx.set_option(k, int(v))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def go():
    if what == "primary":
        ctx.primary_device(w, h, d_hits)
    elif what == "shadow":
        ctx.shadow_device(n, d_rays, d_hits, d_sh)
    elif what == "fused":
        ctx.primary_shadow_device(w, h, None, None, d_vis)
    elif what == "diffuse":
        ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
    else:
        ctx.render_frame_device(w, h, d_vis)


ts = []
for _ in range(launches):
    flush.fill_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        go()
        e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"{scene} {what} scheduler={sched} {opts}: median {np.median(ts[2:]):.4f} ms  all {[round(t, 4) for t in ts]}", flush=True)
