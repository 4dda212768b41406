"""VERDICT r1 task 5a/5b: does the reference's OWN kernel (unmodified volumeRender.cl) become pixel-exact against the
oracle when it is BUILT differently -- build options only, no source edit -- and what is the cubes2 4K coverage mismatch?

    python tools/refcl_options_probe.py            # driver: one subprocess per option set -> gpurun_out/refcl_options.json
    python tools/refcl_options_probe.py --one IDX  # worker

For every option set: the 14 golden frames (7 scenes x default / grazing light, 96x64) and the 1080p bench frame against
the uncontracted fp32 definition (the CPU oracle for the golden scenes, the CUDA path -- itself bit-identical to the oracle
-- at 1080p), counting pixels that differ at all, by more than 1 LSB, and in coverage (hit vs background).
The FP_CONTRACT pragma is injected from the command line by redefining the `__kernel` keyword, in the spirit of the
-D__write_only= the NVIDIA compiler already needs; it covers everything after the first kernel (line 88: traversal,
ray_box, Moller-Trumbore, GGX) but not the two helpers above it (reflect, get_normal_at_tri_point)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PRAGMA = '_Pragma("OPENCL FP_CONTRACT OFF")'
OPTION_SETS = [
    ("stock (clBuildProgram without options, RayTracer.cpp:2173)", ""),
    ("IEEE divide/sqrt", "-cl-fp32-correctly-rounded-divide-sqrt"),
    ("no optimisation", "-cl-opt-disable"),
    ("no optimisation + IEEE divide/sqrt", "-cl-opt-disable -cl-fp32-correctly-rounded-divide-sqrt"),
    ("FP_CONTRACT OFF via -D__kernel (unquoted)", f"-D__kernel={PRAGMA}__kernel"),
    ("FP_CONTRACT OFF via -D__kernel (quoted) + IEEE divide/sqrt",
     f"-cl-fp32-correctly-rounded-divide-sqrt -D '__kernel={PRAGMA} __kernel'"),
    ("FP_CONTRACT OFF via -D__kernel (double-quoted) + IEEE divide/sqrt",
     '-cl-fp32-correctly-rounded-divide-sqrt -D "__kernel=_Pragma(\\"OPENCL FP_CONTRACT OFF\\") __kernel"'),
    ("FP_CONTRACT OFF via -D__kernel (tab-separated) + IEEE divide/sqrt",
     '-cl-fp32-correctly-rounded-divide-sqrt -D__kernel=_Pragma("OPENCL\tFP_CONTRACT\tOFF")__kernel'),
    ("nv opt level 0", "-cl-nv-opt-level=0"),
    ("nv opt level 0 + IEEE divide/sqrt", "-cl-nv-opt-level=0 -cl-fp32-correctly-rounded-divide-sqrt"),
]


def worker(idx):
    import torch  # noqa: F401  (CUDA context for the comparison frame)

    import rtb200
    from conftest import SCENES, channel_diff, load_scene, mesh_dict
    from oracle import oracle_py as O

    label, opts = OPTION_SETS[idx]
    res = {"label": label, "options": opts}
    try:
        O.refcl(opts)
    except O.RefCLUnavailable as e:
        res["build_error"] = str(e)[-600:]
        print(json.dumps(res))
        return
    rows = {}
    tot = {"pixels": 0, "differ": 0, "differ_gt1": 0, "coverage": 0}
    grazing = np.float32([-150, 25, 3, 1])
    for name in SCENES:
        g = load_scene(name)
        md = mesh_dict(g)
        ref = O.RefCLScene(md, g["ref_nodes"], g["ref_tri_indices"], opts)
        sc = O.OracleScene(md, g["ref_nodes"], g["ref_tri_indices"])
        w, h = (int(v) for v in g["wh"])
        for tag, params in (("", g["params"]), ("_grazing", np.concatenate([g["params"][:16], grazing, g["params"][20:]]))):
            img, _ = ref.render_frame(params, w, h)
            want, _ = sc.render_frame(params, w, h)
            d = channel_diff(img, want).max(-1)
            row = {"differ": int((d > 0).sum()), "differ_gt1": int((d > 1).sum()), "coverage": int(np.logical_xor(img != 0, want != 0).sum())}
            rows[name + tag] = row
            tot["pixels"] += w * h
            for k in ("differ", "differ_gt1", "coverage"):
                tot[k] += row[k]
    res["build_note"] = ref.build_note()
    res["golden_14_frames"] = tot
    res["golden_rows"] = rows
    # the 1080p bench frame and the cubes2 frames against the CUDA path (bit-identical to the oracle, tests/test_gpu_parity.py)
    ctx = rtb200.Context(0)

    def compare(md, nodes, tri, params, w, h, explain=False):
        ref = O.RefCLScene(md, nodes, tri, opts)
        ms = []
        for _ in range(4):
            img, t = ref.render_frame(params, w, h)
            ms.append(t)
        ctx.upload_scene(md, nodes, tri)
        ctx.set_params(params)
        ours = ctx.render_frame(w, h)
        d = channel_diff(img, ours).max(-1)
        cov = np.logical_xor(img != 0, ours != 0)
        row = {"frame": [w, h], "pixels": w * h, "differ": int((d > 0).sum()), "differ_gt1": int((d > 1).sum()),
               "coverage": int(cov.sum()), "reference_kernel_ms": float(np.median(ms[1:]))}
        if explain and cov.any():
            sc = O.OracleScene(md, nodes, tri)
            rays, gate = O.primary_rays(params, w, h)
            ex = []
            for y, x in np.argwhere(cov)[:4]:
                i = int(y) * w + int(x)
                hit, cnt = sc.trace(0, rays[i:i + 1])
                brute, _ = sc.trace(0, rays[i:i + 1], brute=True)
                # neighbours: what do the adjacent pixels hit?
                nb = {}
                for dy, dx in ((0, -1), (0, 1), (-1, 0), (1, 0)):
                    yy, xx = int(y) + dy, int(x) + dx
                    if 0 <= yy < h and 0 <= xx < w:
                        hh, _ = sc.trace(0, rays[yy * w + xx:yy * w + xx + 1])
                        nb[f"{dx:+d},{dy:+d}"] = [int(hh["idx"][0]), float(hh["t"][0])]
                ex.append({"pixel": [int(x), int(y)], "reference_pixel": int(img[y, x]), "ours_pixel": int(ours[y, x]),
                           "gate": int(gate[i]), "oracle_hit": [int(hit["idx"][0]), float(hit["t"][0]), float(hit["u"][0]), float(hit["v"][0])],
                           "oracle_brute_force_hit": [int(brute["idx"][0]), float(brute["t"][0])], "neighbours_idx_t": nb,
                           "ray": [float(v) for v in rays[i]]})
            row["coverage_mismatch_explained"] = ex
        return row

    m = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
    A = m.arrays()
    rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
    b = rtb200.FlatBVH.build(m)
    p, _ = rtb200.camera_params(1920, 1080, A["aabb_min"], A["aabb_max"])
    res["terrain_1M_1080p"] = compare(A, b.nodes, b.tri_indices, p, 1920, 1080)
    g = load_scene("cubes2")
    for light in ((-23.0, 200.0, 3.0), (-150.0, 25.0, 3.0)):
        p, _ = rtb200.camera_params(3840, 2160, g["aabb_min"], g["aabb_max"], light_pos=light)
        res[f"cubes2_4k_light_{int(light[0])}"] = compare(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"], p, 3840, 2160, explain=True)
    print(json.dumps(res))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        return worker(int(sys.argv[2]))
    out = []
    for i, (label, _o) in enumerate(OPTION_SETS):
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(i)], capture_output=True, text=True, timeout=600)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if line:
            out.append(json.loads(line[-1]))
        else:
            out.append({"label": label, "options": OPTION_SETS[i][1], "failed": (p.stderr or p.stdout)[-800:]})
        r = out[-1]
        print(label, "->", r.get("build_error", "")[:200] or r.get("failed", "")[:300] or
              {"golden": r["golden_14_frames"], "terrain": {k: r["terrain_1M_1080p"][k] for k in ("differ", "differ_gt1", "coverage", "reference_kernel_ms")},
               "cubes2_4k": [r[k]["coverage"] for k in r if k.startswith("cubes2_4k")]}, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"what": __doc__.splitlines()[0], "option_sets": out}, open(os.path.join(ROOT, "gpurun_out", "refcl_options.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
