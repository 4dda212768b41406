"""primary pass under the tile_order options (0 row-major from y=0, 1 reversed, 2 multiplicative permutation, 3 permuted
64-tile chunks) on the bench scene, default and top-down camera."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7)); A = mesh.arrays()
cache = "/tmp/rtb200_ab_bvh.bin"
bvh = rtb200.FlatBVH.load(cache) if os.path.exists(cache) else rtb200.FlatBVH.build(mesh)
ctx = rtb200.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
d_hits = torch.empty((w * h, 4), device="cuda")
d_img = torch.empty((h, w), dtype=torch.int32, device="cuda")
def timeit(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for name, kw in (("default camera", {}), ("top-down", {"d_beta": 35.0, "d_radius": -50.0})):
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], **kw)
    ctx.set_params(params)
    base = None
    for rnd in range(2):
        out = []
        for order in (0, 1, 2, 3):
            ctx.set_option("tile_order", order)
            tp = timeit(lambda: ctx.primary_device(w, h, d_hits))
            s = int(d_hits.view(torch.int32)[:, 0].to(torch.int64).sum().item())
            base = s if base is None else base
            assert s == base
            tf = timeit(lambda: ctx.render_frame_device(w, h, d_img), iters=10)
            out.append(f"order {order}: primary {tp:.4f} frame {tf:.4f}")
        print(name, " | ".join(out), flush=True)
ctx.set_option("tile_order", 0)
