"""Scratch: time primary/shadow/diffuse for a list of option dicts given on the command line as k=v,k=v ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

configs = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]] or [{}]
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish()
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
ctx.set_params(params)
n = w * h
d_hits = torch.zeros((n, 4), device="cuda")
d_rays = torch.zeros((n, 8), device="cuda")
d_sh = torch.zeros((n, 4), device="cuda")
ctx.set_option("scheduler", 0)
ctx.primary_device(w, h, d_hits, d_rays)
ctx.shadow_device(n, d_rays, d_hits, d_sh)
d_dr = torch.zeros((n * 4, 8), device="cuda")
d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
torch.cuda.synchronize()
nd = int(d_cnt.item())
d_dh = torch.zeros((nd, 4), device="cuda")
ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
torch.cuda.synchronize()
ref_p, ref_s, ref_d = d_hits.clone(), d_sh.clone(), d_dh.clone()
nhit = int((ref_p.view(torch.int32)[:, 0] >= 0).sum())
ntrav = 752541


def timeit(fn, iters=10):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for cfg in configs:
    for k in ("scheduler", "refill", "inner_exit", "blocks_per_sm", "exact_div", "tile_order"):
        default = {"scheduler": -1, "refill": 16, "inner_exit": 8, "blocks_per_sm": 0, "exact_div": 0, "tile_order": 0}[k]
        ctx.set_option(k, int(cfg.get(k, default)))
    tp = timeit(lambda: ctx.primary_device(w, h, d_hits))
    ts = timeit(lambda: ctx.shadow_device(n, d_rays, ref_p, d_sh))
    td = timeit(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh))
    ok = torch.equal(d_hits.view(torch.int32), ref_p.view(torch.int32)) and torch.equal(d_sh.view(torch.int32), ref_s.view(torch.int32)) \
        and torch.equal(d_dh.view(torch.int32), ref_d.view(torch.int32))
    print(f"{str(cfg):60s} primary {tp:.3f} ms ({ntrav / tp / 1e3:.0f})  shadow {ts:.3f} ms ({nhit / ts / 1e3:.0f})  diffuse {td:.3f} ms ({nd / td / 1e3:.0f} Mr/s)  identical={ok}", flush=True)
