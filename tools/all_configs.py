"""All BASELINE.json single-GPU configurations (C1-C4) on one B200: throughput + parity spot checks against
the CPU oracle. Writes profiles/<out>.json.   usage: python tools/all_configs.py [out_name] [--skip-c4]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200
from oracle import oracle_py as O

out_name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "r1_configs"
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
ctx = rtb200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
HIT = rtb200.HIT_DTYPE
results = {"gpu": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count(), "configs": {}}


def timeit(fn, iters=10, warm=3):
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def same(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in ("idx", "t", "u", "v"))


def oracle_rate(sc, mode, rays, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        hits, cnt = sc.trace(mode, rays)
        ts.append(time.perf_counter() - t0)
    return hits, cnt, rays.shape[0] / min(ts) / 1e6


def gate_mask(hits, rays_np):
    return None


def run_scene(tag, mesh, w, h, light=(-23.0, 200.0, 3.0), d_radius=0.0, d_beta=0.0, diffuse=True, check_stride=1, note=""):
    A = mesh.arrays()
    t0 = time.time()
    bvh = rtb200.FlatBVH.build(mesh)
    build_s = time.time() - t0
    params, eye = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=light, d_radius=d_radius, d_beta=d_beta)
    ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
    ctx.set_params(params)
    info = ctx.scene_info()
    n = w * h
    d_hits = torch.zeros((n, 4), device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    d_sh = torch.zeros((n, 4), device="cuda")
    d_sr = torch.zeros((n, 8), device="cuda")
    ctx.primary_device(w, h, d_hits, d_rays)
    ctx.shadow_device(n, d_rays, d_hits, d_sh, d_sr)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
    rays = d_rays.cpu().numpy()
    sh = d_sh.cpu().numpy().view(HIT).reshape(-1)
    srays = d_sr.cpu().numpy()
    sc = O.OracleScene(A, bvh.nodes, bvh.tri_indices)
    orays, gate = O.primary_rays(params, w, h)
    g = gate.astype(bool)
    ntrav, nhit = int(g.sum()), int((hits["idx"] >= 0).sum())
    r = {"triangles": A["indices"].size // 3, "bvh_nodes": bvh.num_nodes, "refs": bvh.num_refs, "duplicated_refs": bvh.duplicates,
         "sbvh_build_s": build_s, "frame": [w, h], "blob_MB": info["blob_bytes"] / 1e6, "tree_depth": info["max_depth"], "note": note,
         "pixels": n, "primary_traversed": ntrav, "primary_hits": nhit}
    # ---- parity on a subsample (stride) ----
    sel = np.flatnonzero(g)[::check_stride]
    want, cnt, cpu_rate = oracle_rate(sc, 0, np.ascontiguousarray(orays[sel]))
    r["primary_parity"] = {"rays_checked": int(sel.size), "raygen_identical": bool(np.array_equal(orays.view(np.uint32), rays.view(np.uint32))),
                           "idx_t_u_v_identical": bool(same(hits[sel], want))}
    r["primary_visits_per_traversed_ray"] = {"inner": cnt["inner"] / sel.size, "tris": cnt["tris"] / sel.size, "max_stack": cnt["max_stack"]}
    r["cpu_oracle_primary_mrays_s"] = cpu_rate
    ms = timeit(lambda: ctx.primary_device(w, h, d_hits))
    r["primary"] = {"ms": ms, "mrays_s_traversed": ntrav / ms / 1e3, "mpixels_s": n / ms / 1e3}
    # ---- shadow ----
    hsel = np.flatnonzero(hits["idx"] >= 0)[::check_stride]
    if hsel.size:
        want_s, cnt_s, cpu_rate_s = oracle_rate(sc, 1, np.ascontiguousarray(srays[hsel]))
        r["shadow_parity"] = {"rays_checked": int(hsel.size), "idx_t_u_v_identical": bool(same(sh[hsel], want_s)),
                              "occluded_fraction": float((want_s["idx"] >= 0).mean())}
        r["shadow_visits_per_ray"] = {"inner": cnt_s["inner"] / hsel.size, "tris": cnt_s["tris"] / hsel.size}
        r["cpu_oracle_shadow_mrays_s"] = cpu_rate_s
        ms = timeit(lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh))
        r["shadow_anyhit"] = {"ms": ms, "mrays_s": nhit / ms / 1e3, "rays": nhit}
    # ---- diffuse 4 spp ----
    if diffuse and nhit:
        spp = 4
        d_dr = torch.zeros((nhit * spp, 8), device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.diffuse_rays_device(n, d_rays, d_hits, spp, 0x5EED, d_dr, d_cnt)
        torch.cuda.synchronize()
        nd = int(d_cnt.item())
        d_dh = torch.zeros((nd, 4), device="cuda")
        ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
        torch.cuda.synchronize()
        dh = d_dh.cpu().numpy().view(HIT).reshape(-1)
        dsel = np.arange(0, nd, max(1, check_stride * 4))
        drays = d_dr.cpu().numpy()
        want_d, cnt_d, cpu_rate_d = oracle_rate(sc, 0, np.ascontiguousarray(drays[dsel]))
        r["diffuse_parity"] = {"rays_checked": int(dsel.size), "idx_t_u_v_identical": bool(same(dh[dsel], want_d)),
                               "hit_fraction": float((want_d["idx"] >= 0).mean())}
        r["diffuse_visits_per_ray"] = {"inner": cnt_d["inner"] / dsel.size, "tris": cnt_d["tris"] / dsel.size, "max_stack": cnt_d["max_stack"]}
        r["cpu_oracle_diffuse_mrays_s"] = cpu_rate_d
        for sched, name in ((1, "lanes"), (0, "batch")):
            ctx.set_option("scheduler", sched)
            ms = timeit(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh), iters=6)
            a_bytes = (64 * cnt_d["inner"] + 48 * cnt_d["tris"]) / dsel.size + 48
            r[f"diffuse_4spp_{name}"] = {"ms": ms, "mrays_s": nd / ms / 1e3, "rays": nd, "algorithmic_GBps": a_bytes * nd / ms / 1e6}
        ctx.set_option("scheduler", -1)
        ms = timeit(lambda: ctx.trace_sorted_device(rtb200.CLOSEST, nd, d_dr, d_dh), iters=6)
        r["diffuse_4spp_lanes_with_sort_prepass"] = {"ms_incl_sort_gather_scatter": ms, "mrays_s": nd / ms / 1e3}
        del d_dr, d_dh
    # ---- frame ----
    d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: ctx.render_frame_device(w, h, d_img), iters=5)
    r["full_frame"] = {"ms": ms, "fps": 1e3 / ms}
    results["configs"][tag] = r
    print(tag, json.dumps(r), flush=True)
    del d_hits, d_rays, d_sh, d_sr, d_img
    torch.cuda.empty_cache()


# C1: icosphere (81 920 tris) written to a COLLADA file and loaded through ColladaLoader -> Mesh::init
with tempfile.TemporaryDirectory() as td:
    src = rtb200.Mesh().icosphere(6, 50.0).finish(diffuse=(0.8, 0.3, 0.2))
    path = os.path.join(td, "c1.dae")
    src.write_dae(path)
    t0 = time.time()
    m1 = rtb200.Mesh().load_dae(path)
    load_s = time.time() - t0
run_scene("C1_640x480_icosphere_81920_dae", m1, 640, 480, note=f"COLLADA file of 81 920 triangles parsed + baked in {load_s:.2f} s")
# C2 / C3: 1 M-triangle terrain, default camera + default light; grazing light; top-down full-coverage view
terrain = rtb200.Mesh().terrain(707, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
run_scene("C2_C3_1080p_terrain_1M_default", terrain, 1920, 1080, check_stride=8)
run_scene("C3_1080p_terrain_1M_grazing_light", terrain, 1920, 1080, light=(-150.0, 25.0, 3.0), diffuse=False, check_stride=8)
run_scene("C2_1080p_terrain_1M_topdown_full_coverage", terrain, 1920, 1080, d_radius=-50.0, d_beta=float(np.deg2rad(35.0)), diffuse=False, check_stride=16)
del terrain
if "--skip-c4" not in sys.argv:
    field = rtb200.Mesh().sphere_field(11, 120.0, 6, 50.0).finish(diffuse=(0.6, 0.6, 0.8))
    run_scene("C4_1080p_sphere_field_10M", field, 1920, 1080, d_radius=1320.0, check_stride=16)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", out_name + ".json"), "w"), indent=1)
