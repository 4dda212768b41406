"""Host scene layer: the SBVH builder / flattener must emit byte-identical arrays to the reference's
own SplitBVHBuilder + BVH_Cuda (golden fixtures made with oracle/_ref/ref_host; live where available)."""
import os
import tempfile

import numpy as np
import pytest
from conftest import SCENES, load_scene, same_bits

import rtb200
from oracle import oracle_py as O


def _build(g, threshold=16384):
    m = rtb200.Mesh().set(g["verts"], g["indices"])
    return rtb200.FlatBVH.build(m, threshold)


@pytest.mark.parametrize("name", SCENES)
def test_builder_matches_reference_golden(name):
    g = load_scene(name)
    b = _build(g)
    assert b.num_nodes == g["ref_nodes"].shape[0] and b.num_refs == g["ref_tri_indices"].size
    assert same_bits(b.nodes, g["ref_nodes"])
    assert np.array_equal(b.tri_indices, g["ref_tri_indices"])


@pytest.mark.parametrize("threshold", [0, 64, 1000])
def test_parallel_build_is_deterministic(threshold):
    """task-parallel subtree builds splice back into the serial (reference) emission order"""
    m = rtb200.Mesh().icosphere(4, 50.0).terrain(40, 100.0).sticks(300, 11, 120.0)
    a = rtb200.FlatBVH.build(m, 0)
    b = rtb200.FlatBVH.build(m, threshold)
    assert same_bits(a.nodes, b.nodes) and np.array_equal(a.tri_indices, b.tri_indices)
    assert a.duplicates == b.duplicates and a.duplicates > 0  # spatial splits did happen


def _ref_build(A, td):
    payload = np.array([A["verts"].shape[0], A["indices"].size // 3], dtype=np.int32).tobytes() + \
        np.ascontiguousarray(A["verts"], dtype=np.float32).tobytes() + np.ascontiguousarray(A["indices"], dtype=np.int32).tobytes()
    out = O.ref_host("build", payload, td)
    N, R = np.frombuffer(out[:8], dtype=np.int32)
    return np.frombuffer(out[8:8 + 48 * N], dtype=np.float32).reshape(N, 12), np.frombuffer(out[8 + 48 * N:], dtype=np.int32)


@pytest.mark.skipif(not O.ref_host_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("maker", [
    lambda: rtb200.Mesh().icosphere(4, 50.0),
    lambda: rtb200.Mesh().terrain(64, 100.0),
    lambda: rtb200.Mesh().sticks(2500, 21, 100.0),
    lambda: rtb200.Mesh().icosphere(3, 40.0, (5.0, 0.0, 0.0)).icosphere(3, 40.0, (-5.0, 3.0, 1.0)).sticks(500, 2, 90.0),
])
def test_builder_matches_reference_live(maker):
    m = maker()
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    with tempfile.TemporaryDirectory() as td:
        nodes, tri = _ref_build(A, td)
    assert same_bits(b.nodes, nodes) and np.array_equal(b.tri_indices, tri)


def test_flat_layout_invariants():
    """48-byte nodes, pre-order (left child = index+1), w = 1, leaves -1/-1, refs pre-multiplied by 3"""
    m = rtb200.Mesh().terrain(30, 100.0)
    b = rtb200.FlatBVH.build(m)
    ints = b.nodes.view(np.int32)
    left, right, off, cnt = ints[:, 8], ints[:, 9], ints[:, 10], ints[:, 11]
    inner = left >= 0
    assert np.all(left[inner] == np.flatnonzero(inner) + 1)
    assert np.all(right[inner] > left[inner]) and np.all(off[inner] == -1) and np.all(cnt[inner] == 0)
    assert np.all(right[~inner] == -1) and np.all(cnt[~inner] >= 1) and np.all(cnt[~inner] <= 8)
    assert np.all(b.nodes[:, 3] == 1.0) and np.all(b.nodes[:, 7] == 1.0)
    assert np.all(b.tri_indices % 3 == 0) and cnt[~inner].sum() == b.num_refs
    # every triangle is referenced at least once
    assert np.unique(b.tri_indices // 3).size == m.arrays()["indices"].size // 3


def test_flat_bvh_disk_cache_roundtrip(tmp_path):
    m = rtb200.Mesh().icosphere(3, 10.0)
    b = rtb200.FlatBVH.build(m)
    p = str(tmp_path / "x.fbvh")
    b.save(p)
    c = rtb200.FlatBVH.load(p)
    assert same_bits(b.nodes, c.nodes) and np.array_equal(b.tri_indices, c.tri_indices)


def test_camera_default_pose_and_params():
    """reference Camera(): radius 200, alpha 225 deg, beta 45 deg -> eye ~ (-100, 141.42, -100) (Camera.cpp:6-19)"""
    params, eye = rtb200.camera_params(640, 480, (-1, -2, -3), (4, 5, 6))
    assert np.allclose(eye, [-100.0, 141.42136, -100.0], atol=1e-3)
    p = params.reshape(8, 4)
    assert np.all(p[:, 3] == 1.0)
    a, b, c, campos = p[0, :3], p[1, :3], p[2, :3], p[3, :3]
    half = np.tan(np.float32(60.0 * 3.1415 * 0.5 / 180.0))
    assert np.isclose(np.linalg.norm(a), 2 * half * 640 / 480, rtol=1e-5) and np.isclose(np.linalg.norm(b), 2 * half, rtol=1e-5)
    assert abs(np.dot(a, b)) < 1e-4
    centre = c + 0.5 * a + 0.5 * b  # image-plane centre = eye + dir
    assert np.isclose(np.linalg.norm(centre - campos), 1.0, atol=1e-5)
    assert np.allclose(p[4, :3], [-23, 200, 3]) and np.allclose(p[6, :3], [-1, -2, -3]) and np.allclose(p[7, :3], [4, 5, 6])


# ---- bit-level pin of Camera + make_params (task 5e / weak 9) -------------------------------------------
# A numpy restatement in explicit fp32 of reference Camera.cpp:6-68 and RayTracer.cpp:634-671 (updateCamera), operation
# for operation. Double sub-expressions of the source (M_PI, the 3.1415 literal) are evaluated in double and rounded where
# the source assigns them to a float. cos / sin / tan of a float are taken as the correctly rounded fp32 value (computed
# in double, rounded once) -- the reference calls its C library's float overloads, whose last bit is not specified; glibc's
# cosf / sinf / tanf agree with the correctly rounded value on every pose below.
F = np.float32


def _f3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


def _dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))             # vectors_math.cpp:73-75


def _cross(a, b):
    return _f3(F(F(a[1] * b[2]) - F(a[2] * b[1])), F(F(a[2] * b[0]) - F(a[0] * b[2])), F(F(a[0] * b[1]) - F(a[1] * b[0])))  # :77-79


def _normalize(v):
    inv = F(F(1.0) / np.sqrt(_dot(v, v), dtype=np.float32))                    # rsqrtf = 1.0f / sqrtf, :18-20, 81-84
    return _f3(F(inv * v[0]), F(inv * v[1]), F(inv * v[2]))


def _trig(fn, x):
    import math
    return F(fn(float(x)))


def params_restated(w, h, light_pos, light_color, aabb_min, aabb_max, d_radius=0.0, rotations=()):
    import math
    PI = math.pi
    radius, alpha, beta = F(200.0), F(0.0), F(0.0)
    up = _f3(0, 1, 0)
    center = _f3(0, 0, 0)

    def add_rotate(da, db):                                                     # Camera.cpp:26-48
        nonlocal alpha, beta, up
        alpha = F(alpha + F(da))
        beta = F(beta + F(db))
        if beta < F(0.0):
            beta = F(float(beta) + 2.0 * PI)                                    # float += double
        elif float(beta) > 2.0 * PI:
            beta = F(float(beta) - 2.0 * PI)
        flipped = (float(beta) > PI / 2.0) and (float(beta) < 3.0 * PI / 2.0)
        up = _f3(0, -1 if flipped else 1, 0)

    add_rotate(F(45 * (3 + 2) * PI / 180.0), F(45 * PI / 180.0))                # Camera.cpp:18
    if d_radius:
        radius = F(radius + F(d_radius))                                        # add_radius, :21-24
    for da, db in rotations:
        add_rotate(F(da), F(db))
    cb, sb = _trig(math.cos, beta), _trig(math.sin, beta)                       # update_eye, :50-59
    eye = _f3(F(center[0] + F(F(radius * cb) * _trig(math.cos, alpha))),
              F(center[1] + F(radius * sb)),
              F(center[2] + F(F(radius * cb) * _trig(math.sin, alpha))))
    direction = _normalize(_f3(*(F(center[i] - eye[i]) for i in range(3))))     # update_full, :61-68
    right = _normalize(_cross(direction, up))
    cr = _cross(direction, right)
    cam_up = _normalize(_f3(-cr[0], -cr[1], -cr[2]))
    theta = F((60.0 * 3.1415 * 0.5) / 180.0)                                    # RayTracer.cpp:639-640
    half = _trig(math.tan, theta)
    aspect = F(F(w) / F(h))
    u0, v0 = F(F(-half) * aspect), F(-half)
    u1, v1 = F(half * aspect), half
    du, dv = F(u1 - u0), F(v1 - v0)
    a = _f3(*(F(du * right[i]) for i in range(3)))                              # :650-652
    b = _f3(*(F(dv * cam_up[i]) for i in range(3)))
    c = _f3(*(F(F(F(eye[i] + F(u0 * right[i])) + F(v0 * cam_up[i])) + F(F(1.0) * direction[i])) for i in range(3)))
    out = np.ones((8, 4), dtype=np.float32)
    for i, row in enumerate((a, b, c, eye, light_pos, light_color, aabb_min, aabb_max)):
        out[i, :3] = np.asarray(row, dtype=np.float32)
    return out.reshape(-1), eye


@pytest.mark.parametrize("w,h", [(640, 480), (1920, 1080), (3840, 2160), (1024, 768), (328, 204)])
@pytest.mark.parametrize("pose", [dict(), dict(d_radius=1320.0), dict(d_alpha=0.3, d_beta=-0.2), dict(d_alpha=-2.5, d_beta=0.61),
                                  dict(d_radius=-50.0, d_alpha=1.0, d_beta=0.5), dict(d_alpha=0.0, d_beta=1.0)])
def test_camera_params_bit_identical_to_the_restated_reference(w, h, pose):
    """Camera::Camera / add_radius / add_rotate / make_params == reference Camera.cpp + updateCamera in fp32, bit for bit
    (the golden frames and every parity test inherit these 128 bytes)"""
    light, colour, lo, hi = (-23.0, 200.0, 3.0), (1.0, 1.0, 1.0), (-100.0, -7.4, -100.0), (100.0, 7.4, 100.0)
    got, eye = rtb200.camera_params(w, h, lo, hi, light_pos=light, light_color=colour, **pose)
    rot = [(pose.get("d_alpha", 0.0), pose.get("d_beta", 0.0))] if ("d_alpha" in pose or "d_beta" in pose) else []
    want, weye = params_restated(w, h, light, colour, lo, hi, d_radius=pose.get("d_radius", 0.0), rotations=rot)
    assert same_bits(eye, weye), (eye, weye)
    assert same_bits(got, want), np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))


def test_collada_roundtrip(tmp_path):
    """generated .dae -> ColladaLoader -> Mesh::init reproduces the mesh exactly (BASELINE config 1 path)"""
    m = rtb200.Mesh().icosphere(2, 50.0).finish(diffuse=(0.25, 0.5, 0.75))
    A = m.arrays()
    p = str(tmp_path / "s.dae")
    m.write_dae(p)
    B = rtb200.Mesh().load_dae(p).arrays()
    for k in ("verts", "indices", "normals", "normal_indices", "tri_to_material"):
        assert same_bits(A[k], B[k]), k
    assert np.allclose(A["materials"][0, 12:15], B["materials"][0, 12:15])
    assert same_bits(A["aabb_min"], B["aabb_min"]) and same_bits(A["aabb_max"], B["aabb_max"])


@pytest.mark.parametrize("damage", ["count_not_multiple", "count_huge", "count_negative", "vertex_index", "normal_index",
                                    "material_name", "p_count_huge"])
def test_collada_malformed_files_fail_cleanly(tmp_path, damage):
    """a float_array count that the text cannot back, or <p> indices / material names that point nowhere,
    must fail the load (IOError through rth_mesh_load_dae) -- no heap overflow, no exception through extern "C", no
    out-of-range index handed to the SBVH builder."""
    import re
    m = rtb200.Mesh().icosphere(1, 10.0).finish(diffuse=(0.25, 0.5, 0.75))
    p = str(tmp_path / "ok.dae")
    m.write_dae(p)
    txt = open(p).read()
    counts = re.findall(r'<float_array[^>]*count="(\d+)"', txt)
    assert counts, "generated file has float arrays"
    first = f'count="{counts[0]}"'
    if damage == "count_not_multiple":   # was: 1-2 floats written past the allocation; now the tail is ignored
        bad = txt.replace(first, f'count="{int(counts[0]) - 1}"', 1)
    elif damage == "count_huge":
        bad = txt.replace(first, 'count="2000000000"', 1)
    elif damage == "count_negative":
        bad = txt.replace(first, 'count="-5"', 1)
    elif damage == "vertex_index":
        bad = re.sub(r"<p>\s*(\d+) ", "<p>999999 ", txt, count=1)
    elif damage == "normal_index":
        bad = re.sub(r"<p>\s*(\d+) (\d+) ", r"<p>\1 -7 ", txt, count=1)
    elif damage == "material_name":
        bad = re.sub(r'(<polygons[^>]*material=")[^"]*"', r'\1no_such_effect"', txt, count=1)
        bad = re.sub(r"<library_effects>.*</library_effects>", "<library_effects></library_effects>", bad, flags=re.S)
    else:
        bad = re.sub(r'(<polygons[^>]*count=")\d+"', r'\g<1>2000000000"', txt, count=1)
    assert bad != txt
    q = str(tmp_path / "bad.dae")
    open(q, "w").write(bad)
    if damage in ("count_not_multiple", "p_count_huge"):
        try:   # either the load fails, or what it delivers is in range
            B = rtb200.Mesh().load_dae(q).arrays()
        except IOError:
            return
        assert B["indices"].size == 0 or B["indices"].max() < B["verts"].shape[0]
    else:
        with pytest.raises(IOError):
            rtb200.Mesh().load_dae(q)


REF_CUBES2 = "/root/reference/x64/Release/data/collada/cubes2.DAE"


@pytest.mark.skipif(not os.path.exists(REF_CUBES2), reason="the reference tree is only mounted in the build container")
def test_loader_on_the_reference_scene_file():
    """the COLLADA file the reference itself loads (RayTracer.cpp:862) through the new loader reproduces the committed
    fixture (tests/golden/scene_cubes2.npz): 23 392 triangles, 13 materials, the file's own normals"""
    A = rtb200.Mesh().load_dae(REF_CUBES2).arrays()
    g = load_scene("cubes2")
    assert A["indices"].size // 3 == 23392 and A["materials"].shape[0] == 13
    for key, gkey in (("verts", "verts"), ("indices", "indices"), ("normals", "normals"), ("normal_indices", "normal_indices"),
                      ("materials", "materials"), ("tri_to_material", "tri_to_material")):
        assert same_bits(np.asarray(A[key]), g[gkey]), key
