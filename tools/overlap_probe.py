"""rt_render_frame_begin/_end with two frames in flight: frames on per-slot streams (overlap_frames=1) vs on the context
stream (0); 1080p terrain, pinned destinations, a host-side consumer (checksum) per frame."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7)); A = mesh.arrays()
cache = "/tmp/rtb200_ab_bvh.bin"
bvh = rtb200.FlatBVH.load(cache) if os.path.exists(cache) else rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0); ctx.upload_scene(A, bvh.nodes, bvh.tri_indices); ctx.set_params(params)
bufs = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for _ in range(4)]
nf = 300
def run(slots, consume):
    acc = 0
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(params)
        ctx.render_frame_begin(w, h, bufs[k % slots], k % slots)
        if k >= slots - 1:
            j = (k - (slots - 1)) % slots
            ctx.render_frame_end(j)
            if consume: acc += int(bufs[j].numpy()[::8, ::8].sum())
    for k in range(nf - (slots - 1), nf):
        ctx.render_frame_end(k % slots)
    return (time.perf_counter() - t0) / nf * 1e3
for rnd in range(2):
  for sg in (-1, 0, 2):
    ctx.set_option("store_group", sg)
    for ov in (0, 1):
        ctx.set_option("overlap_frames", ov)
        print(f"store_group={sg} overlap_frames={ov}: " + "  ".join(f"slots={s} consume={c}: {run(s, c):.3f} ms/frame" for s in (1, 2, 3) for c in (False, True)), flush=True)
