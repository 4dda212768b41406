"""Run the reference's own OpenCL kernel (oracle/_ref/libref_cl.so) on this box's OpenCL device and compare its frames
with the CPU oracle and with the CUDA path; time it on the bench scene."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from oracle import oracle_py as O
from conftest import SCENES, load_scene, mesh_dict, channel_diff

opts = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    O.refcl(opts)
except O.RefCLUnavailable as e:
    print("reference OpenCL unavailable:", e); sys.exit(0)
ctx = rtb200.Context(0)
for name in SCENES:
    g = load_scene(name); md = mesh_dict(g)
    ref = O.RefCLScene(md, g["ref_nodes"], g["ref_tri_indices"], opts)
    w, h = (int(v) for v in g["wh"])
    img, ms = ref.render_frame(g["params"], w, h)
    ctx.upload_scene(md, g["ref_nodes"], g["ref_tri_indices"]); ctx.set_params(g["params"])
    cuda = ctx.render_frame(w, h)
    d_o, d_c = channel_diff(img, g["frame"]), channel_diff(img, cuda)
    print(f"{name:10s} device={ref.device_name()} ref-vs-oracle: max {d_o.max()} LSB, pixels differing {(d_o.max(-1) > 0).sum()} (>1 LSB: {(d_o.max(-1) > 1).sum()}) of {w*h}; "
          f"ref-vs-cuda: max {d_c.max()}, differing {(d_c.max(-1) > 0).sum()}; coverage mismatch {(np.logical_xor(img != 0, g['frame'] != 0)).sum()}")
# bench scene
mesh = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7)); A = mesh.arrays(); bvh = rtb200.FlatBVH.build(mesh)
w, h = 1920, 1080
params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
ref = O.RefCLScene(A, bvh.nodes, bvh.tri_indices, opts)
times = []
for _ in range(6):
    t0 = time.perf_counter(); img, ms = ref.render_frame(params, w, h); times.append((ms, (time.perf_counter() - t0) * 1e3))
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices); ctx.set_params(params)
cuda = ctx.render_frame(w, h)
t0 = time.perf_counter()
for _ in range(10): cuda = ctx.render_frame(w, h)
ours_ms = (time.perf_counter() - t0) * 100
d = channel_diff(img, cuda)
print(f"C2 1080p frame: reference OpenCL kernel {np.median([t[0] for t in times]):.3f} ms (kernel), {np.median([t[1] for t in times]):.3f} ms (call incl. readback); "
      f"ours rt_render_frame {ours_ms:.3f} ms (call incl. readback); ref-vs-cuda max {d.max()} LSB, pixels differing {(d.max(-1) > 0).sum()} (>1: {(d.max(-1) > 1).sum()}) of {w*h}")
