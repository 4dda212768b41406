import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from conftest import load_scene, mesh_dict
w, h = 1920, 1080
g = load_scene("cubes2"); md = mesh_dict(g)
ctx = rtb200.Context(0); ctx.upload_scene(md, g["ref_nodes"], g["ref_tri_indices"])
params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"]); ctx.set_params(params)
img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
ctx.set_option("frame_mode", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(3): ctx.render_frame_device(w, h, img)
ctx.synchronize(); print("done")
