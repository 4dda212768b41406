"""Scratch: speed and mismatch count of the NON-exact fast box test (closest-hit, batch scheduler) vs the exact kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
rtb200.hostlib.set_num_threads(os.cpu_count())
ctx = rtb200.Context(0); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def timeit(fn, iters=10):
    with torch.cuda.stream(stream):
        for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(); fn(); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
def run(tag, mesh, w, h, d_radius=0.0):
    A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], d_radius=d_radius)
    ctx.upload_scene(A, bvh.nodes, bvh.tri_indices, shading=False); ctx.set_params(params)
    n = w * h
    hits = torch.zeros((n, 4), device="cuda"); rays = torch.zeros((n, 8), device="cuda")
    ctx.set_option("fast_box", 0); ctx.set_option("scheduler", 0)
    ctx.primary_device(w, h, hits, rays); torch.cuda.synchronize()
    t0 = timeit(lambda: ctx.primary_device(w, h, hits)); ref = hits.clone()
    nh = int((ref.view(torch.int32)[:, 0] >= 0).sum())
    dr = torch.zeros((nh * 4, 8), device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.diffuse_rays_device(n, rays, ref, 4, 0x5EED, dr, cnt); torch.cuda.synchronize(); nd = int(cnt.item())
    dh = torch.zeros((nd, 4), device="cuda")
    td0 = timeit(lambda: ctx.trace_device(0, nd, dr, dh)); refd = dh.clone()
    ctx.set_option("fast_box", 1)
    t1 = timeit(lambda: ctx.primary_device(w, h, hits))
    td1 = timeit(lambda: ctx.trace_device(0, nd, dr, dh))
    def diff(a, b):
        ai, bi = a.view(torch.int32), b.view(torch.int32)
        idx_mis = int((ai[:, 0] != bi[:, 0]).sum()); bit_mis = int((ai != bi).any(dim=1).sum())
        both = (ai[:, 0] >= 0) & (bi[:, 0] >= 0) & (ai[:, 0] != bi[:, 0])
        rel = float(((a[both, 1] - b[both, 1]).abs() / b[both, 1].abs()).max()) if both.any() else 0.0
        hm = int(((ai[:, 0] >= 0) != (bi[:, 0] >= 0)).sum())
        return idx_mis, bit_mis, hm, rel
    print(f"{tag}: primary exact {t0:.3f} ms fast {t1:.3f} ms ({t0/t1:.2f}x) mismatches(idx, any-bit, hit/miss, max rel dt)={diff(hits, ref)} of {n} | "
          f"diffuse(batch) exact {td0:.3f} fast {td1:.3f} ms ({td0/td1:.2f}x) mismatches={diff(dh, refd)} of {nd}", flush=True)
    ctx.set_option("fast_box", 0); ctx.set_option("scheduler", -1)
run("C1 sphere", rtb200.Mesh().icosphere(6, 50.0).finish(), 1920, 1080)
run("C2 terrain", rtb200.Mesh().terrain(707, 100.0).finish(), 1920, 1080)
run("sticks+mix", rtb200.Mesh().terrain(160, 100.0).icosphere(5, 25.0, (20.0, 30.0, -10.0)).sticks(4000, 5, 150.0).finish(), 1920, 1080)
run("C4 spheres 10M", rtb200.Mesh().sphere_field(11, 120.0, 6, 50.0).finish(), 1920, 1080, d_radius=1320.0)
