"""Host-side packing (rt_pack_scene_host: the step rt_upload_scene performs before the upload) on CPU:
layout of the blob, validation errors, and -- with a pure-numpy walk over the PACKED arrays that follows the
kernel's loop (csrc/traverse.cuh) -- equality with the oracle walking the REFERENCE arrays."""
import numpy as np
import pytest
from conftest import SCENES, load_scene, mesh_dict

import rtb200
from oracle import oracle_py as O

HDR = np.dtype([("magic", "<u4"), ("version", "<u4"), ("total_bytes", "<u8"), ("root_ref", "<i4"), ("num_pairs", "<i4"),
                ("num_tris", "<i4"), ("top_pairs", "<i4"), ("max_depth", "<i4"), ("V", "<i4"), ("T", "<i4"), ("Vn", "<i4"),
                ("M", "<i4"), ("num_ref_nodes", "<i4"), ("coords_in_window", "<i4"), ("reserved", "<i4"),
                ("off", "<u8", 8), ("root_min", "<f4", 3), ("root_max", "<f4", 3)])
POISON = -(2 ** 31)
TMIN = np.float32(0.001)


def unpack(blob):
    h = np.frombuffer(blob[:HDR.itemsize], dtype=HDR)[0]
    off = h["off"]
    pairs = np.frombuffer(blob, dtype=np.float32, count=int(h["num_pairs"]) * 16, offset=int(off[0])).reshape(-1, 16)
    tris = np.frombuffer(blob, dtype=np.float32, count=(int(h["num_tris"]) + 1) * 12, offset=int(off[1])).reshape(-1, 12)
    return h, pairs, tris


def walk(h, pairs, tris, o, d, tmax, any_hit):
    """csrc/traverse.cuh restated with numpy float32 scalars (true division, NaN-ignoring min/max)."""
    f = np.float32
    o, d = o.astype(f), d.astype(f)
    stack, cur, t_hit, idx, uu, vv = [], int(h["root_ref"]), f(tmax), -1, f(0), f(0)
    pi, ti = pairs.view(np.int32), tris.view(np.int32)
    with np.errstate(all="ignore"):
        while True:
            while cur >= 0:
                q = pairs[cur]
                res = []
                for base in (0, 8):  # child box = {min.x, max.x, min.y, max.y} {min.z, max.z, ref, 0}
                    t0 = (q[[base, base + 2, base + 4]] - o) / d
                    t1 = (q[[base + 1, base + 3, base + 5]] - o) / d
                    tn = np.fmax(np.fmax(np.fmin(t0[0], t1[0]), np.fmin(t0[1], t1[1])), np.fmin(t0[2], t1[2]))
                    tf = np.fmin(np.fmin(np.fmax(t0[0], t1[0]), np.fmax(t0[1], t1[1])), np.fmax(t0[2], t1[2]))
                    res.append((tn, tf, bool(tn <= tf and tf >= TMIN and tn <= t_hit)))
                c0, c1 = int(pi[cur, 6]), int(pi[cur, 14])
                if res[0][2] and res[1][2]:
                    if res[0][0] > res[1][0]:
                        c0, c1 = c1, c0
                    if len(stack) >= 64:
                        return -1, t_hit, f(0), f(0)
                    stack.append(c1)
                    cur = c0
                elif res[0][2]:
                    cur = c0
                elif res[1][2]:
                    cur = c1
                elif not stack:
                    return idx, t_hit, uu, vv
                else:
                    cur = stack.pop()
            if cur == POISON:
                return -1, t_hit, f(0), f(0)
            k = ~cur
            while True:
                v0, e1, e2 = tris[k, 0:3], tris[k, 4:7], tris[k, 8:11]
                tvec = o - v0
                pvec = np.array([d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]], dtype=f)
                det = f(1.0) / (e1[0] * pvec[0] + e1[1] * pvec[1] + e1[2] * pvec[2])
                u = (tvec[0] * pvec[0] + tvec[1] * pvec[1] + tvec[2] * pvec[2]) * det
                t = f(-1.0)
                if not (u < 0 or u > 1):
                    qv = np.array([tvec[1] * e1[2] - tvec[2] * e1[1], tvec[2] * e1[0] - tvec[0] * e1[2], tvec[0] * e1[1] - tvec[1] * e1[0]], dtype=f)
                    v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * det
                    if not (v < 0 or (u + v) > 1):
                        t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * det
                if t < t_hit and t > TMIN:
                    t_hit, idx, uu, vv = t, int(ti[k, 3]), u, v
                    if any_hit:
                        return idx, t_hit, uu, vv
                if ti[k, 7] != 0:
                    break
                k += 1
            if not stack:
                return idx, t_hit, uu, vv
            cur = stack.pop()


@pytest.mark.parametrize("name", SCENES)
def test_blob_layout_matches_reference_arrays(name):
    g = load_scene(name)
    blob = rtb200.pack_scene_host(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    h, pairs, tris = unpack(blob)
    nodes, ints = g["ref_nodes"], g["ref_nodes"].view(np.int32)
    inner = np.flatnonzero(ints[:, 8] >= 0)
    assert h["magic"] == 0x42325452 and h["total_bytes"] == blob.size and h["num_ref_nodes"] == nodes.shape[0]
    assert h["num_pairs"] == inner.size and h["num_tris"] == g["ref_tri_indices"].size and all(int(o) % 256 == 0 for o in h["off"])
    # every pair holds exactly the two child boxes of one reference inner node, left child first
    want = {tuple(np.concatenate([nodes[ints[i, 8], 0:3], nodes[ints[i, 8], 4:7], nodes[ints[i, 9], 0:3], nodes[ints[i, 9], 4:7]]).view(np.uint32))
            for i in inner}
    got = {tuple(p[[0, 2, 4, 1, 3, 5, 8, 10, 12, 9, 11, 13]].view(np.uint32)) for p in pairs}
    assert got == want
    # triangles: v0 | 3*triId, e1 | last, e2, in tri_indices order
    ti = tris.view(np.int32)
    V, I = g["verts"], g["indices"]
    for k, tri in enumerate(g["ref_tri_indices"]):
        v0, v1, v2 = V[I[tri], :3], V[I[tri + 1], :3], V[I[tri + 2], :3]
        assert ti[k, 3] == tri and np.array_equal(tris[k, 0:3], v0) and np.array_equal(tris[k, 4:7], v1 - v0) and np.array_equal(tris[k, 8:11], v2 - v0)
    leaves = np.flatnonzero(ints[:, 8] < 0)
    last = np.zeros(g["ref_tri_indices"].size + 1, dtype=bool)
    for i in leaves:
        if ints[i, 11] > 0:
            last[ints[i, 10] + ints[i, 11] - 1] = True
    last[-1] = True  # sentinel
    assert np.array_equal(ti[:, 7] != 0, last)


@pytest.mark.parametrize("name", ["test0", "test1", "ico2", "mix"])
@pytest.mark.parametrize("top_pairs", [0, 5, 2047])
def test_walking_the_packed_blob_equals_the_oracle(name, top_pairs):
    """the packed layout + the kernel's loop (restated in numpy) give the oracle's hits on the reference layout, for any
    choice of the breadth-first prefix"""
    g = load_scene(name)
    h, pairs, tris = unpack(rtb200.pack_scene_host(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"], top_pairs=top_pairs))
    assert h["top_pairs"] == min(top_pairs, h["num_pairs"])
    rays = g["random_rays"][:160]
    for mode, key in ((0, "random_hits_closest"), (1, "random_hits_any")):
        want = g[key].view(O.HIT_DTYPE).reshape(-1)[:160]
        for i, r in enumerate(rays):
            idx, t, u, v = walk(h, pairs, tris, r[0:3], r[4:7], r[3], mode == 1)
            assert idx == want["idx"][i] and np.float32(t).view(np.uint32) == want["t"][i].view(np.uint32), (name, mode, i)
            if idx >= 0:
                assert np.float32(u).view(np.uint32) == want["u"][i].view(np.uint32) and np.float32(v).view(np.uint32) == want["v"][i].view(np.uint32)


def test_validation_errors():
    g = load_scene("ico2")
    md = mesh_dict(g)
    ints = g["ref_nodes"].view(np.int32)
    leaf = int(np.flatnonzero(ints[:, 8] < 0)[0])
    bad = g["ref_nodes"].copy()
    bad.view(np.int32)[leaf, 11] = 10 ** 6
    with pytest.raises(rtb200.RtError, match="triangle range"):
        rtb200.pack_scene_host(md, bad, g["ref_tri_indices"])
    tri = g["ref_tri_indices"].copy()
    tri[3] = 3 * (g["indices"].size // 3)
    with pytest.raises(rtb200.RtError, match="outside the index buffer"):
        rtb200.pack_scene_host(md, g["ref_nodes"], tri)
    cyc = g["ref_nodes"].copy()
    inner = int(np.flatnonzero(ints[:, 8] >= 0)[1])
    cyc.view(np.int32)[inner, 8:10] = (0, 0)
    with pytest.raises(rtb200.RtError, match="reachable twice"):
        rtb200.pack_scene_host(md, cyc, g["ref_tri_indices"])
    idx = g["indices"].copy()
    idx[0] = -1
    with pytest.raises(rtb200.RtError, match="outside"):
        rtb200.pack_scene_host({**md, "indices": idx}, g["ref_nodes"], g["ref_tri_indices"])
    # a malformed inner node is not an error: the reference returns -1 there, the packer marks it
    poison = g["ref_nodes"].copy()
    poison.view(np.int32)[0, 9] = 10 ** 6
    h, _pairs, _tris = unpack(rtb200.pack_scene_host(md, poison, g["ref_tri_indices"]))
    assert h["root_ref"] == POISON and h["num_pairs"] == 0


def test_window_flag_and_shading_sections():
    g = load_scene("mix")
    md = mesh_dict(g)
    h, _, _ = unpack(rtb200.pack_scene_host(md, g["ref_nodes"], g["ref_tri_indices"]))
    assert h["coords_in_window"] == 1 and h["Vn"] == g["normals"].shape[0] and h["M"] == 1
    far = g["ref_nodes"].copy()
    far[1, 4] = 3e38
    assert unpack(rtb200.pack_scene_host(md, far, g["ref_tri_indices"]))[0]["coords_in_window"] == 0
    h2, _, _ = unpack(rtb200.pack_scene_host(md, g["ref_nodes"], g["ref_tri_indices"], shading=False))
    assert h2["Vn"] == 0 and h2["M"] == 0 and h2["total_bytes"] < h["total_bytes"]
