"""The CPU oracle against (1) the reference's own compiled code, via committed golden vectors and --
where oracle/_ref exists -- live; (2) the author's brute-force loop; (3) its own committed outputs."""
import ctypes as C
import tempfile

import numpy as np
import pytest
from conftest import SCENES, assert_hits_identical, load_scene, mesh_dict, same_bits

from oracle import oracle_py as O


@pytest.fixture(scope="module")
def probes():
    import os

    from conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "ref_probes.npz"))
    return {k: z[k] for k in z.files}


def _mt(rows):
    L = O.lib()
    out = np.empty(rows.shape[0], dtype=np.float32)
    for i, r in enumerate(rows):
        r = np.ascontiguousarray(r)
        p = r.ctypes.data
        out[i] = L.orc_ray_triangle(p, p + 12, p + 24, p + 36, p + 48, None, None)
    return out


def test_moller_trumbore_matches_reference_common_h(probes):
    """oracle RayTriangleIntersection == spec::RayTriangleIntersection (reference common.h:193-219), bit for bit"""
    got = _mt(probes["mt_in"])
    assert (probes["mt_out"] > 0).sum() > 1000  # the vectors do exercise the hit path
    assert same_bits(got, probes["mt_out"])


def test_scene_box_gate_matches_reference_common_h(probes):
    """oracle RayBoxIntersection == spec::RayBoxIntersection (reference common.h:172-190)"""
    L = O.lib()
    rows = probes["box_in"]
    got = np.empty((rows.shape[0], 3), dtype=np.uint32)
    for i, r in enumerate(rows):
        r = np.ascontiguousarray(r)
        p = r.ctypes.data
        tmin, tmax = C.c_float(), C.c_float()
        hit = L.orc_scene_box_gate(p, p + 12, p + 24, p + 36, C.byref(tmin), C.byref(tmax))
        got[i] = (hit, np.float32(tmin.value).view(np.uint32), np.float32(tmax.value).view(np.uint32))
    assert np.array_equal(got, probes["box_out"])


def test_dot_cross_normalize_match_reference_vectors_math(probes):
    L = O.lib()
    rows = probes["vec_in"]
    got = np.empty((rows.shape[0], 7), dtype=np.float32)
    for i, r in enumerate(rows):
        r = np.ascontiguousarray(r)
        L.orc_vec_probe(r.ctypes.data, r.ctypes.data + 12, got[i].ctypes.data)
    assert same_bits(got, probes["vec_out"])


@pytest.mark.skipif(not O.ref_host_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_moller_trumbore_live_against_ref_host():
    rng = np.random.default_rng(5)
    n = 3000
    rows = (rng.normal(size=(n, 15)) * 10).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        want = np.frombuffer(O.ref_host("mt", np.int32(n).tobytes() + rows.tobytes(), td), dtype=np.float32)
    assert same_bits(_mt(rows), want)


@pytest.mark.parametrize("name", SCENES)
def test_traversal_equals_authors_bruteforce_loop(name):
    """BVH traversal (vR.cl:776-1006) == brute force over all triangles (vR.cl:690-712) for closest hit.
    Index may differ only on exact-t ties (the two loops visit triangles in different orders)."""
    g = load_scene(name)
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    for rays in (g["primary_rays"], g["random_rays"]):
        a, _ = sc.trace(0, rays)
        b, _ = sc.trace(0, rays, brute=True)
        assert same_bits(a["t"], b["t"])
        differ = a["idx"] != b["idx"]
        assert differ.mean() < 0.01  # ties only (same t, checked above)


@pytest.mark.parametrize("name", SCENES)
def test_oracle_reproduces_committed_golden(name):
    g = load_scene(name)
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = (int(v) for v in g["wh"])
    rays, gate = O.primary_rays(g["params"], w, h)
    assert same_bits(rays, g["primary_rays"]) and same_bits(gate, g["primary_gate"])
    hits, _ = sc.trace(0, rays)
    assert_hits_identical(hits, g["primary_hits"].view(O.HIT_DTYPE).reshape(-1), "primary")
    srays, valid = O.shadow_rays(g["params"], rays, hits)
    assert same_bits(srays, g["shadow_rays"]) and same_bits(valid, g["shadow_valid"])
    rc, _ = sc.trace(0, g["random_rays"])
    ra, _ = sc.trace(1, g["random_rays"])
    assert_hits_identical(rc, g["random_hits_closest"].view(O.HIT_DTYPE).reshape(-1), "random closest")
    assert_hits_identical(ra, g["random_hits_any"].view(O.HIT_DTYPE).reshape(-1), "random any")
    img, _ = sc.render_frame(g["params"], w, h)
    assert np.array_equal(img, g["frame"])


def test_any_hit_is_first_accepted_not_closest():
    """any-hit returns the first accepted triangle in traversal order: its t is >= the closest t, and the
    occlusion boolean agrees with closest-hit (vR.cl:986)."""
    g = load_scene("mix")
    c = g["random_hits_closest"].view(O.HIT_DTYPE).reshape(-1)
    a = g["random_hits_any"].view(O.HIT_DTYPE).reshape(-1)
    assert np.array_equal(c["idx"] >= 0, a["idx"] >= 0)
    hit = c["idx"] >= 0
    assert np.all(a["t"][hit] >= c["t"][hit])
    assert np.any(a["t"][hit] > c["t"][hit])  # and it is genuinely order dependent


def test_tmax_and_tmin_rules():
    """accept iff t < tHit && t > 0.001 (strict both sides, vR.cl:978)"""
    g = load_scene("terrain12")
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    rays = g["primary_rays"].copy()
    hits = g["primary_hits"].view(O.HIT_DTYPE).reshape(-1)
    sel = np.flatnonzero(hits["idx"] >= 0)[:200]
    r = rays[sel].copy()
    r[:, 3] = hits["t"][sel]  # tmax == t exactly -> strict '<' rejects that triangle
    h2, _ = sc.trace(0, r)
    assert np.all((h2["idx"] != hits["idx"][sel]) | (h2["t"] < hits["t"][sel]))
    assert np.all(h2["t"] <= hits["t"][sel])


def test_empty_batch():
    g = load_scene("test0")
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    hits, cnt = sc.trace(0, np.zeros((0, 8), dtype=np.float32))
    assert hits.size == 0 and cnt["inner"] == 0
