import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
w, h = 1920, 1080
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7)); A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
ctx = rtb200.Context(0); ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"]); ctx.set_params(params)
img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
ctx.set_option("frame_mode", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(3): ctx.render_frame_device(w, h, img)
ctx.synchronize(); print("done")
