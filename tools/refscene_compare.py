"""The reference's own scene (cubes2.DAE, from the committed fixture tests/golden/scene_cubes2.npz), its default camera
and its default window (1024x768, RayTracer.cpp:39-40): reference OpenCL kernel vs rt_render_frame_device on the same GPU.
Also 1920x1080 and 3840x2160. Writes gpurun_out/refscene_compare.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from conftest import channel_diff, load_scene, mesh_dict
from oracle import oracle_py as O

g = load_scene("cubes2")
md = mesh_dict(g)
ref = O.RefCLScene(md, g["ref_nodes"], g["ref_tri_indices"])
ctx = rtb200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.upload_scene(md, g["ref_nodes"], g["ref_tri_indices"])
rows = []
for w, h in ((1024, 768), (1920, 1080), (3840, 2160)):
    for light in ((-23.0, 200.0, 3.0), (-150.0, 25.0, 3.0)):
        params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"], light_pos=light)
        ms_ref = []
        for _ in range(8):
            img, ms = ref.render_frame(params, w, h)
            ms_ref.append(ms)
        ctx.set_params(params)
        d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        ts = {}
        for mode in (0,):
            for _ in range(3):
                ctx.render_frame_device(w, h, d_img)
            torch.cuda.synchronize()
            t = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.render_frame_device(w, h, d_img)
                b.record()
                torch.cuda.synchronize()
                t.append(a.elapsed_time(b))
            ts[mode] = float(np.median(t))
        ours = d_img.cpu().numpy().view(np.uint32)
        d = channel_diff(img, ours).max(-1)
        row = {"frame": [w, h], "light": light, "reference_opencl_kernel_ms": float(np.median(ms_ref[2:])), "ours_frame_ms": ts[0], "pixels_differing": int((d > 0).sum()), "pixels_differing_by_more_than_1_lsb": int((d > 1).sum()),
               "coverage_mismatch": int(np.logical_xor(img != 0, ours != 0).sum())}
        rows.append(row)
        print(row)
out = {"scene": "cubes2.DAE (23 392 triangles, 13 materials), reference default camera", "device": ref.device_name(),
       "reference_build": ref.build_note(), "rows": rows}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "refscene_compare.json"), "w"), indent=1)
