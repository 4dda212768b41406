"""Scratch GPU sweep of scheduler options on the C2 scene: primary / shadow / diffuse timings.
usage: python tools/sweep.py [quads=707]"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 707
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(nq, 100.0).finish()
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


def timeit(fn, iters=8):
    with torch.cuda.stream(stream):
        for _ in range(2):
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


def run(tag):
    tp = timeit(lambda: ctx.primary_device(w, h, d_hits))
    ts = timeit(lambda: ctx.shadow_device(n, d_rays, ref_p, d_sh))
    td = timeit(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh))
    ok = torch.equal(d_hits.view(torch.int32), ref_p.view(torch.int32)) and torch.equal(d_sh.view(torch.int32), ref_s.view(torch.int32)) \
        and torch.equal(d_dh.view(torch.int32), ref_d.view(torch.int32))
    print(f"{tag:42s} primary {tp:.3f} ms  shadow {ts:.3f} ms ({nhit / ts / 1e3:.0f} Mr/s)  diffuse {td:.3f} ms ({nd / td / 1e3:.0f} Mr/s)  identical={ok}", flush=True)


run("batch scheduler")
ctx.set_option("scheduler", 1)
for refill, iexit in itertools.product([1, 4, 8, 16, 32], [0, 8, 16]):
    ctx.set_option("refill", refill)
    ctx.set_option("inner_exit", iexit)
    run(f"lanes refill={refill} inner_exit={iexit}")
for bps in (4, 6, 8):
    ctx.set_option("refill", 8)
    ctx.set_option("inner_exit", 0)
    ctx.set_option("blocks_per_sm", bps)
    run(f"lanes refill=8 blocks_per_sm={bps}")
