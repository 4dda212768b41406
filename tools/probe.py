"""Scratch GPU probe: parity + throughput of the primary / shadow / diffuse paths on one scene.
usage: python tools/probe.py [terrain_quads=707] [w=1920] [h=1080]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from oracle import oracle_py as O

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 707
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
check = os.environ.get("PROBE_CHECK", "1") == "1"

t0 = time.time()
mesh = rtb200.Mesh().terrain(nq, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
print(f"scene: {A['indices'].size // 3} tris, {bvh.num_nodes} nodes, {bvh.num_refs} refs, build {time.time() - t0:.1f}s, cpus={os.cpu_count()}")
params, eye = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])

torch.cuda.init()
ctx = rtb200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.time()
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
print("upload", time.time() - t0, ctx.scene_info())
ctx.set_params(params)
n = w * h
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
d_shits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
d_srays = torch.empty((n, 8), dtype=torch.float32, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


ctx.primary_device(w, h, d_hits, d_rays)
torch.cuda.synchronize()
hits = d_hits.cpu().numpy().view(rtb200.HIT_DTYPE).reshape(-1)
rays = d_rays.cpu().numpy()
print("primary hit fraction", (hits["idx"] >= 0).mean())
if check:
    sc = O.OracleScene(A, bvh.nodes, bvh.tri_indices)
    orays, gate = O.primary_rays(params, w, h)
    print("raygen bit-identical:", np.array_equal(orays.view(np.uint32), rays.view(np.uint32)))
    t0 = time.time()
    want, cnt = sc.trace(0, orays)
    dt = time.time() - t0
    want["idx"][gate == 0] = -1
    want["t"][gate == 0] = rtb200.T_INIT
    want["u"][gate == 0] = 0
    want["v"][gate == 0] = 0
    print(f"oracle: {n / dt / 1e6:.2f} Mrays/s on {O.lib().orc_num_threads()} threads; per ray I={cnt['inner'] / n:.2f} L={cnt['leaf'] / n:.2f} T={cnt['tris'] / n:.2f} maxstack={cnt['max_stack']}")
    for k in ("idx", "t", "u", "v"):
        print("  primary", k, "identical:", np.array_equal(hits[k].view(np.uint32), want[k].view(np.uint32)))

for opt in [dict(), dict(smem_top=1023), dict(smem_top=2047), dict(blocks_per_sm=4), dict(blocks_per_sm=8), dict(blocks_per_sm=12), dict(blocks_per_sm=16)]:
    ctx.set_option("smem_top", opt.get("smem_top", 0))
    ctx.set_option("blocks_per_sm", opt.get("blocks_per_sm", 0))
    med, mn = timeit(lambda: ctx.primary_device(w, h, d_hits))
    print(f"primary {opt}: median {med:.3f} ms min {mn:.3f} ms -> {n / med / 1e3:.1f} Mrays/s")
ctx.set_option("smem_top", 0)
ctx.set_option("blocks_per_sm", 0)

# shadow rays
ctx.primary_device(w, h, d_hits, d_rays)
ctx.shadow_device(n, d_rays, d_hits, d_shits, d_srays)
torch.cuda.synchronize()
shits = d_shits.cpu().numpy().view(rtb200.HIT_DTYPE).reshape(-1)
nvalid = int((hits["idx"] >= 0).sum())
if check:
    srays, valid = O.shadow_rays(params, orays, want)
    got_srays = d_srays.cpu().numpy()
    v = valid.astype(bool)
    print("shadow raygen identical:", np.array_equal(srays[v].view(np.uint32), got_srays[v].view(np.uint32)))
    want_s, cnt_s = sc.trace(1, srays[v])
    print(f"  shadow per ray I={cnt_s['inner'] / v.sum():.2f} T={cnt_s['tris'] / v.sum():.2f} occluded={(want_s['idx'] >= 0).mean():.3f}")
    for k in ("idx", "t", "u", "v"):
        print("  shadow", k, "identical:", np.array_equal(shits[k][v].view(np.uint32), want_s[k].view(np.uint32)))
med, mn = timeit(lambda: ctx.shadow_device(n, d_rays, d_hits, d_shits))
print(f"shadow: median {med:.3f} ms -> {nvalid / med / 1e3:.1f} Mrays/s ({nvalid} rays)")

# diffuse
spp = 4
d_drays = torch.empty((n * spp, 8), dtype=torch.float32, device="cuda")
d_count = torch.zeros(1, dtype=torch.int64, device="cuda")
ctx.diffuse_rays_device(n, d_rays, d_hits, spp, 0x5EED, d_drays, d_count)
torch.cuda.synchronize()
nd = int(d_count.item())
d_dhits = torch.empty((nd, 4), dtype=torch.float32, device="cuda")
ctx.trace_device(rtb200.CLOSEST, nd, d_drays, d_dhits)
torch.cuda.synchronize()
dh = d_dhits.cpu().numpy().view(rtb200.HIT_DTYPE).reshape(-1)
print("diffuse rays", nd, "hit frac", (dh["idx"] >= 0).mean())
if check:
    dr = d_drays[:nd].cpu().numpy()
    sub = slice(0, min(nd, 400000))
    want_d, cnt_d = sc.trace(0, dr[sub])
    print(f"  diffuse per ray I={cnt_d['inner'] / want_d.size:.2f} T={cnt_d['tris'] / want_d.size:.2f}")
    for k in ("idx", "t", "u", "v"):
        print("  diffuse", k, "identical:", np.array_equal(dh[k][sub].view(np.uint32), want_d[k].view(np.uint32)))
med, mn = timeit(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_drays, d_dhits))
print(f"diffuse: median {med:.3f} ms -> {nd / med / 1e3:.1f} Mrays/s")

# frame
d_img = torch.empty((h, w), dtype=torch.int32, device="cuda")
ctx.render_frame_device(w, h, d_img)
torch.cuda.synchronize()
med, mn = timeit(lambda: ctx.render_frame_device(w, h, d_img), iters=5)
print(f"render_frame: median {med:.3f} ms")
if check and w * h <= 640 * 480:
    ref_img, _ = sc.render_frame(params, w, h)
    img = d_img.cpu().numpy().view(np.uint32)
    diff = np.abs(((img[..., None] >> np.array([0, 8, 16])) & 255).astype(np.int32) - ((ref_img[..., None] >> np.array([0, 8, 16])) & 255).astype(np.int32))
    print("frame max LSB diff", diff.max(), "pixels differing", (diff.max(axis=-1) > 0).sum())
print(ctx.counters())
