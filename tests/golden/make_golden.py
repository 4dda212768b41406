"""Generate the committed golden fixtures. Run HERE (the container that has /root/reference):

    python tests/golden/make_golden.py

Sources of truth, in order of authority:
  * ref_*.npz   -- outputs of the reference's OWN compiled code (oracle/_ref/ref_host, built from
                   /root/reference unmodified by oracle/Makefile): Moller-Trumbore, scene-box gate,
                   dot/cross/normalize probes, and SplitBVHBuilder + BVH_Cuda flat trees.
  * trace_*.npz -- outputs of the CPU oracle (oracle/oracle.c) walking those reference-built trees.
                   The reference cannot produce these itself (its kernel is OpenCL, no ICD here).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import rtb200  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

assert O.ref_host_available(), "build oracle/_ref first: make -C oracle ref"
rng = np.random.default_rng(20261018)
CUBES2 = "/root/reference/x64/Release/data/collada/cubes2.DAE"
ONLY = sys.argv[1] if len(sys.argv) > 1 else None  # regenerate one scene only (it then draws from its own seed)
FILE_SCENES = {"cubes2"}                            # meshes that come from a file: keep the file's normals / materials


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def scenes():
    yield "test0", rtb200.Mesh().bvh_test(0)
    yield "test1", rtb200.Mesh().bvh_test(1)
    yield "ico2", rtb200.Mesh().icosphere(2, 50.0)
    yield "terrain12", rtb200.Mesh().terrain(12, 100.0)
    yield "sticks150", rtb200.Mesh().sticks(150, 3, 100.0)
    yield "mix", rtb200.Mesh().icosphere(2, 30.0, (10.0, 20.0, -5.0)).terrain(10, 80.0).sticks(40, 9, 60.0)
    # the scene the reference itself loads (RayTracer.cpp:862): its COLLADA file through the new loader, 23 392 triangles,
    # 13 materials, smooth normals from the file; seen from the reference's default camera
    yield "cubes2", rtb200.Mesh().load_dae(CUBES2)


def probes(td):
    # ---- single-function probes against the reference's compiled common.h / vectors_math.cpp ----
    n = 4096
    o = f32(rng.normal(size=(n, 3)) * 50)
    d = f32(rng.normal(size=(n, 3)))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v0 = f32(rng.normal(size=(n, 3)) * 20)
    e1 = f32(rng.normal(size=(n, 3)) * 30)
    e2 = f32(rng.normal(size=(n, 3)) * 30)
    # make most rays actually hit: aim at a random point inside the triangle
    bu, bv = rng.random(n) * 0.6, rng.random(n) * 0.4
    target = v0 + e1 * bu[:, None] + e2 * bv[:, None]
    aim = rng.random(n) < 0.8
    dd = f32(target - o)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    d[aim] = dd[aim]
    mt_in = f32(np.hstack([o, d, v0, e1, e2]))
    mt_out = np.frombuffer(O.ref_host("mt", np.int32(n).tobytes() + mt_in.tobytes(), td), dtype=np.float32).copy()

    bmin = f32(rng.normal(size=(n, 3)) * 40 - 30)
    bmax = f32(bmin + np.abs(rng.normal(size=(n, 3))) * 60)
    org = f32(rng.normal(size=(n, 3)) * 80)
    dirs = f32(rng.normal(size=(n, 3)))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    box_in = f32(np.hstack([bmin, bmax, org, (1.0 / dirs).astype(np.float32)]))
    raw = np.frombuffer(O.ref_host("box", np.int32(n).tobytes() + box_in.tobytes(), td), dtype=np.uint32).reshape(n, 3).copy()

    vec_in = f32(rng.normal(size=(n, 6)) * 10)
    vec_out = np.frombuffer(O.ref_host("vec", np.int32(n).tobytes() + vec_in.tobytes(), td), dtype=np.float32).reshape(n, 7).copy()
    np.savez_compressed(os.path.join(HERE, "ref_probes.npz"), mt_in=mt_in, mt_out=mt_out, box_in=box_in, box_out=raw,
                        vec_in=vec_in, vec_out=vec_out)
    print("ref_probes: MT hits", int((mt_out > 0).sum()), "box hits", int(raw[:, 0].sum()))



with tempfile.TemporaryDirectory() as td:
    if ONLY is None:
        probes(td)

    # ---- reference-built flat BVHs + oracle traces over them --------------------------------
    for name, mesh in scenes():
        if ONLY is not None and name != ONLY:
            continue
        if name in FILE_SCENES:
            import zlib
            rng = np.random.default_rng(zlib.crc32(name.encode()))
        else:
            mesh.finish(diffuse=(0.8, 0.55, 0.3))
        A = mesh.arrays()
        payload = np.array([A["verts"].shape[0], A["indices"].size // 3], dtype=np.int32).tobytes() + \
            f32(A["verts"]).tobytes() + np.ascontiguousarray(A["indices"], dtype=np.int32).tobytes()
        out = O.ref_host("build", payload, td)
        N, R = np.frombuffer(out[:8], dtype=np.int32)
        nodes = np.frombuffer(out[8:8 + 48 * N], dtype=np.float32).reshape(N, 12).copy()
        tri = np.frombuffer(out[8 + 48 * N:], dtype=np.int32).copy()
        assert tri.size == R

        w, h = 96, 64
        params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=(-150.0, 25.0, 3.0) if name == "mix" else (-23.0, 200.0, 3.0),
                                         d_radius=-60.0 if name.startswith("test") else 0.0)
        sc = O.OracleScene(A, nodes, tri)
        rays, gate = O.primary_rays(params, w, h)
        hits, cnt = sc.trace(0, rays)
        srays, valid = O.shadow_rays(params, rays, hits)
        shits, _ = sc.trace(1, srays)
        shits[valid == 0] = (-1, O.T_INIT, 0, 0)
        # incoherent rays: random origins inside the (inflated) scene box, random directions
        k = 2048
        lo, hi = A["aabb_min"] - 5, A["aabb_max"] + 5
        ro = f32(lo + rng.random((k, 3)) * (hi - lo))
        rd = f32(rng.normal(size=(k, 3)))
        rd /= np.linalg.norm(rd, axis=1, keepdims=True)
        rrays = np.zeros((k, 8), dtype=np.float32)
        rrays[:, 0:3], rrays[:, 3], rrays[:, 4:7] = ro, O.T_INIT, rd
        rrays[::7, 3] = f32(rng.random(rrays[::7].shape[0]) * 60 + 1)  # some finite tmax
        rh_c, _ = sc.trace(0, rrays)
        rh_a, _ = sc.trace(1, rrays)
        img, _ = sc.render_frame(params, w, h)
        np.savez_compressed(os.path.join(HERE, f"scene_{name}.npz"), verts=f32(A["verts"]), indices=A["indices"].astype(np.int32),
                            normals=f32(A["normals"]), normal_indices=A["normal_indices"].astype(np.int32),
                            materials=f32(A["materials"]), tri_to_material=A["tri_to_material"].astype(np.int32),
                            aabb_min=A["aabb_min"], aabb_max=A["aabb_max"],
                            ref_nodes=nodes, ref_tri_indices=tri, params=params, wh=np.array([w, h]),
                            primary_rays=rays, primary_gate=gate, primary_hits=hits, shadow_rays=srays, shadow_valid=valid,
                            shadow_hits=shits, random_rays=rrays, random_hits_closest=rh_c, random_hits_any=rh_a, frame=img)
        print(f"{name}: T={A['indices'].size // 3} N={N} R={R} primary hit {np.mean(hits['idx'] >= 0):.2f} "
              f"shadow occl {np.mean(shits['idx'][valid > 0] >= 0) if valid.any() else 0:.2f} random hit {np.mean(rh_c['idx'] >= 0):.2f}")
