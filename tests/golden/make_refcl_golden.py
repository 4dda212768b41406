"""Generate tests/golden/refcl_frames.npz: frames rendered by the reference's OWN OpenCL kernel (unmodified
volumeRender.cl, embedded in oracle/_ref/libref_cl.so) for every golden scene. Needs an OpenCL device, i.e. run it on
the GPU box (NVIDIA's OpenCL ICD on the B200):

    gpurun -- 'python tests/golden/make_refcl_golden.py gpurun_out/refcl_frames.npz'

and commit the result under tests/golden/. The CPU test suite then checks the oracle's frames against these
reference-produced frames without needing a GPU."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import SCENES, load_scene, mesh_dict  # noqa: E402

from oracle import oracle_py as O  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "refcl_frames.npz")
data = {}
for name in SCENES:
    g = load_scene(name)
    ref = O.RefCLScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = (int(v) for v in g["wh"])
    img, _ = ref.render_frame(g["params"], w, h)
    data[name] = img
    data[name + "_grazing"] = ref.render_frame(np.concatenate([g["params"][:16], np.float32([-150, 25, 3, 1]), g["params"][20:]]), w, h)[0]
    print(name, "coverage", float((img != 0).mean()), "device", ref.device_name(), ref.build_note())
data["device"] = np.array(ref.device_name())
data["build_note"] = np.array(ref.build_note())
np.savez_compressed(out, **data)
print("wrote", out)
