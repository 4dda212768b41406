"""Frame time of rt_render_frame_device under option settings, on the bench terrain and on cubes2 (1080p), interleaved
rounds; frames must stay bit-identical. usage: python tools/frame_opts.py "wf_shadow_lanes=0" "wf_shadow_lanes=1" ..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from conftest import load_scene, mesh_dict

w, h = 1920, 1080
settings = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv) for s in sys.argv[1:]] or [{}]
scenes = {}
g = load_scene("cubes2")
scenes["cubes2"] = (mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"], g["aabb_min"], g["aabb_max"])
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
A = mesh.arrays()
cache = os.path.join(ROOT, "gpurun_out", "ab_bvh.bin")
bvh = rtb200.FlatBVH.load(cache) if os.path.exists(cache) else rtb200.FlatBVH.build(mesh)
scenes["terrain1M"] = (A, bvh.nodes, bvh.tri_indices, A["aabb_min"], A["aabb_max"])
for name, (md, nodes, tri, mn, mx) in scenes.items():
    for light in ((-23.0, 200.0, 3.0), (-150.0, 25.0, 3.0)):
        ctx = rtb200.Context(0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.upload_scene(md, nodes, tri)
        params, _ = rtb200.camera_params(w, h, mn, mx, light_pos=light)
        ctx.set_params(params)
        img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        base = None
        times = [[] for _ in settings]
        for rnd in range(4):
            for k, st in enumerate(settings):
                for o, v in st.items():
                    ctx.set_option(o, v)
                for _ in range(2):
                    ctx.render_frame_device(w, h, img)
                torch.cuda.synchronize()
                for _ in range(6):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    ctx.render_frame_device(w, h, img)
                    b.record()
                    torch.cuda.synchronize()
                    times[k].append(a.elapsed_time(b))
                got = img.cpu().numpy().copy()
                if base is None:
                    base = got
                elif not np.array_equal(base, got):
                    print("FRAME DIFFERS under", st)
        print(name, light, " | ".join(f"{st}: {np.median(t):.4f} ms" for st, t in zip(settings, times)))
        ctx.close()
