"""Small ncu target: a few launches of one kernel configuration on the C2 scene.
usage: python tools/prof_target.py {primary|shadow|diffuse} [scheduler=0|1] [refill] [inner_exit] [quads]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

what = sys.argv[1] if len(sys.argv) > 1 else "primary"
sched = int(sys.argv[2]) if len(sys.argv) > 2 else 0
refill = int(sys.argv[3]) if len(sys.argv) > 3 else 16
iexit = int(sys.argv[4]) if len(sys.argv) > 4 else 8
nq = int(sys.argv[5]) if len(sys.argv) > 5 else 707
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(nq, 100.0).finish()
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
ctx.set_params(params)
n = w * h
d_hits = torch.zeros((n, 4), device="cuda")
d_rays = torch.zeros((n, 8), device="cuda")
d_sh = torch.zeros((n, 4), device="cuda")
ctx.set_option("scheduler", 0)
ctx.primary_device(w, h, d_hits, d_rays)
ctx.synchronize()
if what == "diffuse":
    d_dr = torch.zeros((n * 4, 8), device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
    ctx.synchronize()
    nd = int(d_cnt.item())
    d_dh = torch.zeros((nd, 4), device="cuda")
ctx.set_option("scheduler", sched)
ctx.set_option("refill", refill)
ctx.set_option("inner_exit", iexit)
for _ in range(4):
    if what == "primary":
        ctx.primary_device(w, h, d_hits)
    elif what == "shadow":
        ctx.shadow_device(n, d_rays, d_hits, d_sh)
    else:
        ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
ctx.synchronize()
print("done", what, sched)
