"""Option "l2_warm" (bulk L2 prefetch of node pairs + packed triangles at the start of a launch, kernels.cuh warm_l2) on the
bench scene: every camera-ray pass and the ray-buffer passes, L2 flushed before each timed launch (the bench's rule) and
not flushed (a renderer drawing frame after frame), modes interleaved over several rounds.
usage: python tools/l2warm_probe.py [out.json]"""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, use_flush, steps=15, warm=4):
    for _ in range(warm):
        if use_flush:
            flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        if use_flush:
            flush.fill_(1)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def passes(w, h, part, n_parts, band_rows):
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
    ctx.set_params(params)
    n = w * h
    d_hits = torch.zeros((n, 4), device="cuda")
    d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    kw = dict(part=part, n_parts=n_parts, band_rows=band_rows)
    out = {"primary": lambda: ctx.primary_device(w, h, d_hits, None, **kw),
           "fused": lambda: ctx.primary_shadow_device(w, h, None, None, d_vis, **kw),
           "frame": lambda: ctx.render_frame_device(w, h, d_img, **kw)}
    if n_parts == 1:
        d_rays = torch.zeros((n, 8), device="cuda")
        d_sh = torch.zeros((n, 4), device="cuda")
        ctx.primary_device(w, h, d_hits, d_rays)
        d_dr = torch.empty((n * 4, 8), device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
        nd = int(d_cnt.item())
        d_dh = torch.empty((nd, 4), device="cuda")
        out["shadow"] = lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh)
        out["diffuse"] = lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
    return out, (d_hits, d_vis, d_img)


MODES = [(0, 16), (1, 16), (2, 16), (3, 16), (1, 4), (1, 64), (1, 256)]
res = {}
for label, geom in (("1080p", (1920, 1080, 0, 1, 16)), ("4K part 3 of 8", (3840, 2160, 3, 8, 16))):
    fns, bufs = passes(*geom)
    sums = {}
    for rnd in range(3):
        for mode, chunk in MODES:
            ctx.set_option("l2_warm", mode)
            ctx.set_option("l2_warm_chunk_kb", chunk)
            for name, fn in fns.items():
                for fl in (True, False):
                    if not fl and (mode, chunk) not in ((0, 16), (1, 16)):
                        continue
                    res.setdefault((label, name, "flushed" if fl else "not flushed", mode, chunk), []).append(timeit(fn, fl))
            torch.cuda.synchronize()
            sums.setdefault((mode, chunk), [int(b.view(torch.int32).to(torch.int64).sum().item()) for b in bufs])
    assert len({tuple(v) for v in sums.values()}) == 1, "results changed with l2_warm"
ctx.set_option("l2_warm", 0)
table = {}
for (label, name, fl, mode, chunk), v in res.items():
    table.setdefault(f"{label} | {name} | L2 {fl}", {})[f"mode {mode} chunk {chunk} KB"] = round(float(np.median(v)), 4)
for k, v in table.items():
    print(k, v)
if len(sys.argv) > 1:
    json.dump({"what": "ms per launch, median of 3 rounds x 15 launches; l2_warm mode 0 off, 1 pairs+triangles, 2 pairs, 3 triangles",
               "scene": "C2 terrain, 999 698 triangles, traversal data %.1f MB" % ((64 * bvh.nodes.shape[0] / 2 + 48 * len(bvh.tri_indices)) / 1e6),
               "results": table}, open(sys.argv[1], "w"), indent=1)
