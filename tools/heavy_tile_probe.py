"""How long does ONE heavy tile take on its own? (config 5 tail analysis, see tools/timeline_probe.py)
The 32 primary rays of the 8x4 tile at (x0, y0) of the 3840x2160 bench frame, traced as one warp through rt_trace_device
(batch scheduler), cold (L2 flushed) and warm (repeated), next to a typical tile. With the tools build the n8 timelines are
taken without the L2 flush as well."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

W, H = 3840, 2160
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
mesh = rtb200.Mesh().terrain(707, 100.0).finish()
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(W, H, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
ctx.set_params(params)
ctx.set_option("scheduler", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
d_hits = torch.zeros((W * H, 4), device="cuda")
d_rays = torch.zeros((W * H, 8), device="cuda")
with torch.cuda.stream(stream):
    ctx.primary_device(W, H, d_hits, d_rays)
torch.cuda.synchronize()
rays = d_rays.view(H, W, 8)


def tile(x0, y0):
    return rays[y0:y0 + 4, x0:x0 + 8].reshape(32, 8).contiguous()


def time_once(r, cold):
    out = torch.zeros((r.shape[0], 4), device="cuda")
    if cold:
        flush.fill_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        ctx.trace_device(rtb200.CLOSEST, r.shape[0], r, out)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


res = {}
empty = torch.zeros((32, 8), device="cuda")
empty[:, 3] = 4294967296.0
empty[:, 4:7] = torch.tensor([0.0, 1.0, 0.0])  # straight up: misses at the root
for name, r in (("launch_floor_32_rays_missing_the_root", empty), ("heavy_tile_2976_816", tile(2976, 816)), ("typical_tile_1000_1000", tile(1000, 1000)),
                ("typical_tile_1920_1080", tile(1920, 1080))):
    for _ in range(3):
        time_once(r, False)
    warm = [time_once(r, False) for _ in range(9)]
    cold = [time_once(r, True) for _ in range(9)]
    res[name] = {"warm_us": float(np.median(warm)), "cold_us": float(np.median(cold))}
    print(name, res[name], flush=True)
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/heavy_tile_probe.json", "w"), indent=1)
