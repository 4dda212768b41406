"""A/B timing of two builds of librtb200.so on the bench scene: primary / shadow / diffuse / frame, interleaved rounds.
usage: python tools/ab_time.py libA.so libB.so [rounds=5]   (each library is loaded in its own subprocess per round)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import rtb200

    w, h = 1920, 1080
    mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
    A = mesh.arrays()
    cache = "/tmp/rtb200_ab_bvh.bin"
    bvh = rtb200.FlatBVH.load(cache) if os.path.exists(cache) else rtb200.FlatBVH.build(mesh)
    if not os.path.exists(cache):
        bvh.save(cache)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
    ctx = rtb200.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
    ctx.set_params(params)
    n = w * h
    d_hits = torch.empty((n, 4), device="cuda")
    d_rays = torch.empty((n, 8), device="cuda")
    d_sh = torch.empty((n, 4), device="cuda")
    d_img = torch.empty((h, w), dtype=torch.int32, device="cuda")

    def timeit(fn, iters=20, warm=5):
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
        return float(np.median(ts))

    ctx.primary_device(w, h, d_hits, d_rays)
    torch.cuda.synchronize()
    out = {"primary": timeit(lambda: ctx.primary_device(w, h, d_hits)),
           "shadow": timeit(lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh)),
           "frame": timeit(lambda: ctx.render_frame_device(w, h, d_img), iters=10)}
    d_dr = torch.empty((n * 4, 8), device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
    nd = int(d_cnt.item())
    d_dh = torch.empty((nd, 4), device="cuda")
    out["diffuse"] = timeit(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh), iters=10)
    out["hits_sum"] = int(d_hits.view(torch.int32)[:, 0].to(torch.int64).sum().item())
    print(json.dumps(out))
    sys.exit(0)

libs = sys.argv[1:3]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 5
res = {l: [] for l in libs}
for r in range(rounds):
    for l in libs:
        env = dict(os.environ, RTB200_LIB=os.path.abspath(l))
        o = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if o.returncode:
            print(o.stderr[-2000:])
            sys.exit(1)
        res[l].append(json.loads(o.stdout.strip().splitlines()[-1]))
for l in libs:
    keys = [k for k in res[l][0] if k != "hits_sum"]
    print(l, {k: round(sorted(x[k] for x in res[l])[len(res[l]) // 2], 4) for k in keys}, "hits_sum", res[l][0]["hits_sum"])
