"""Option "smem_carveout": how much of the SM's L1/shared array the traversal kernels leave to L1 (the driver's default for
these kernels is a 32 KB shared-memory configuration, ncu launch__shared_mem_config_size). Bench scene, L2 flushed before each
timed launch, settings interleaved over rounds.   usage: python tools/carveout_probe.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
ctx = rtb200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
w, h = 1920, 1080
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ctx.set_params(params)
n = w * h
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
d_hits = torch.zeros((n, 4), device="cuda")
d_rays = torch.zeros((n, 8), device="cuda")
d_sh = torch.zeros((n, 4), device="cuda")
d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
ctx.primary_device(w, h, d_hits, d_rays)
fns = {"primary": lambda: ctx.primary_device(w, h, d_hits),
       "fused": lambda: ctx.primary_shadow_device(w, h, None, None, d_vis),
       "shadow": lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh),
       "frame": lambda: ctx.render_frame_device(w, h, d_img)}


def timeit(fn, steps=15, warm=4):
    for _ in range(warm):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


res = {}
for rnd in range(4):
    for carve in (-1, 0, 6, 12, 25, 50):
        ctx.set_option("smem_carveout", carve)
        for name, fn in fns.items():
            res.setdefault((name, carve), []).append(timeit(fn))
table = {}
for (name, carve), v in res.items():
    table.setdefault(name, {})["carveout %d" % carve] = round(float(np.median(v)), 4)
for k, v in table.items():
    print(k, v)
if len(sys.argv) > 1:
    json.dump({"what": "ms per launch at 1080p on the C2 terrain, L2 flushed, median of 4 rounds x 15 launches; carveout = percent of the "
                       "L1/shared array preferred for shared memory (-1 = driver default, 0 = largest L1)", "results": table},
              open(sys.argv[1], "w"), indent=1)
