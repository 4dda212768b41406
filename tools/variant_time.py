"""Timing of several builds of librtb200.so (make -C real-time-opencl-raytracer_b200/csrc variant ...) on the bench scene at 1080p:
primary / fused primary+shadow / shadow rays / shaded frame, L2 flushed before every timed launch, rounds interleaved, each
library in its own subprocess; a checksum of every result buffer must agree between the builds.
usage: python tools/variant_time.py rounds libA.so libB.so ..."""
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

    rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    d_hits = torch.zeros((n, 4), device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    d_sh = torch.zeros((n, 4), device="cuda")
    d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")

    def timeit(fn, iters=15, warm=4):
        for _ in range(warm):
            flush.fill_(1)
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.fill_(1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    ctx.primary_device(w, h, d_hits, d_rays)
    out = {"primary": timeit(lambda: ctx.primary_device(w, h, d_hits)),
           "fused": timeit(lambda: ctx.primary_shadow_device(w, h, None, None, d_vis)),
           "shadow": timeit(lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh)),
           "frame": timeit(lambda: ctx.render_frame_device(w, h, d_img))}
    out["checksum"] = [int(t.view(torch.int32).to(torch.int64).sum().item()) for t in (d_hits, d_sh, d_vis, d_img)]
    print(json.dumps(out))
    sys.exit(0)

rounds = int(sys.argv[1])
libs = sys.argv[2:]
res = {l: [] for l in libs}
for r in range(rounds):
    for l in libs:
        env = dict(os.environ, RTB200_LIB=os.path.abspath(l))
        o = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if o.returncode:
            print(l, "FAILED", o.stderr[-1500:])
            continue
        res[l].append(json.loads(o.stdout.strip().splitlines()[-1]))
sums = {json.dumps(x["checksum"]) for v in res.values() for x in v}
print("checksums agree" if len(sums) == 1 else "CHECKSUMS DIFFER: %s" % sums)
for l in libs:
    if res[l]:
        print(os.path.basename(l).ljust(34), {k: round(sorted(x[k] for x in res[l])[len(res[l]) // 2], 4) for k in ("primary", "fused", "shadow", "frame")})
