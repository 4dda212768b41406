"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the committed
golden fixtures. Bar: bit-exact hit index, t, u, v (integer/index work AND the fp32 results, because the
kernels keep the reference's operation order with FMA contraction off); frames within +-1 LSB (powf)."""
import numpy as np
import pytest
from conftest import SCENES, assert_hits_identical, channel_diff, load_scene, mesh_dict, same_bits, gpu_context

import rtb200
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
HIT = rtb200.HIT_DTYPE


@pytest.fixture(scope="module")
def ctx():
    c = gpu_context()
    yield c
    c.close()


def _upload(ctx, g):
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    ctx.set_params(g["params"])


@pytest.mark.parametrize("name", SCENES)
def test_trace_host_buffers_vs_golden(ctx, name):
    """rt_trace (host rays in, host hits out) == committed oracle outputs on reference-built trees"""
    g = load_scene(name)
    _upload(ctx, g)
    got = ctx.trace(rtb200.CLOSEST, g["primary_rays"])
    want = g["primary_hits"].view(HIT).reshape(-1)
    # rt_trace traces every ray it is given; the golden primary hits are ungated too
    assert_hits_identical(got, want, f"{name} primary closest")
    assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), f"{name} random closest")
    assert_hits_identical(ctx.trace(rtb200.ANY, g["random_rays"]), g["random_hits_any"].view(HIT).reshape(-1), f"{name} random any")
    v = g["shadow_valid"].astype(bool)
    if v.any():
        assert_hits_identical(ctx.trace(rtb200.ANY, g["shadow_rays"][v]), g["shadow_hits"].view(HIT).reshape(-1)[v], f"{name} shadow any")


@pytest.mark.parametrize("name", SCENES)
def test_fused_primary_and_shadow_kernels_vs_golden(ctx, name):
    """in-kernel ray generation (camera, shadow) is bit-identical to the oracle's, gate included"""
    import torch

    g = load_scene(name)
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    n = w * h
    d_hits = torch.full((n, 4), -7.0, device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    ctx.primary_device(w, h, d_hits, d_rays)
    ctx.synchronize()
    assert same_bits(d_rays.cpu().numpy(), g["primary_rays"])
    want = g["primary_hits"].view(HIT).reshape(-1).copy()
    gate = g["primary_gate"].astype(bool)
    want[~gate] = (-1, rtb200.T_INIT, 0, 0)  # the kernel does not traverse rays that fail the scene-AABB gate
    got = d_hits.cpu().numpy().view(HIT).reshape(-1)
    assert_hits_identical(got, want, f"{name} fused primary")

    d_sh = torch.zeros((n, 4), device="cuda")
    d_sr = torch.zeros((n, 8), device="cuda")
    ctx.shadow_device(n, d_rays, d_hits, d_sh, d_sr)
    ctx.synchronize()
    v = g["shadow_valid"].astype(bool) & gate
    assert same_bits(d_sr.cpu().numpy()[v], g["shadow_rays"][v])
    assert_hits_identical(d_sh.cpu().numpy().view(HIT).reshape(-1)[v], g["shadow_hits"].view(HIT).reshape(-1)[v], f"{name} fused shadow")
    miss = ~(want["idx"] >= 0)
    assert np.all(d_sh.cpu().numpy().view(HIT).reshape(-1)["idx"][miss] == -1)


@pytest.mark.parametrize("name", SCENES)
def test_render_frame_vs_golden(ctx, name):
    g = load_scene(name)
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    img = ctx.render_frame(w, h)
    d = channel_diff(img, g["frame"])
    assert d.max() <= 1, f"{name}: frame differs by {d.max()} LSB on {(d.max(axis=-1) > 1).sum()} pixels"
    assert (d.max(axis=-1) > 0).mean() < 0.02
    assert np.all(img >> 24 == 0)


def test_hoisted_division_is_the_ieee_quotient(ctx):
    """the box test's per-ray-hoisted division (FMUL + 2 FFMA) == `/` bit for bit on 2 G random operand pairs drawn
    from the admitted window, and the window has margin: exact at numerator exponents -104 and +105 too"""
    for seed in (1, 2):
        assert ctx.selftest(1 << 30, seed) == 0
    for ex in (-104, -100, -77, -20, 0, 61, 100, 105):
        assert ctx.selftest_range(1 << 22, ex) == 0
    assert ctx.selftest_range(1 << 22, -125) > 0  # and it does break outside: the guard is needed


def test_scenes_with_tiny_coordinates_use_the_hoisted_path(ctx):
    """icospheres contain coordinates like 1e-16 (and exact zeros): still inside the admitted window"""
    for name in ("ico2", "mix", "terrain12"):
        g = load_scene(name)
        _upload(ctx, g)
        assert ctx.scene_info()["hoisted_division"]
    huge = load_scene("ico2")
    nodes = huge["ref_nodes"].copy()
    nodes[1, 4] = 1e30  # one child-box coordinate outside the window -> the whole scene falls back to the full division
    ctx.upload_scene(mesh_dict(huge), nodes, huge["ref_tri_indices"])
    assert not ctx.scene_info()["hoisted_division"]


@pytest.mark.parametrize("exact_div", [0, 1])
def test_exact_div_option_gives_identical_results(ctx, exact_div):
    g = load_scene("mix")
    _upload(ctx, g)
    ctx.set_option("exact_div", exact_div)
    try:
        assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), "closest")
        assert_hits_identical(ctx.trace(rtb200.ANY, g["random_rays"]), g["random_hits_any"].view(HIT).reshape(-1), "any")
    finally:
        ctx.set_option("exact_div", 0)


@pytest.mark.parametrize("smem_top", [0, 63, 500])
def test_smem_top_variant_is_identical(ctx, smem_top):
    g = load_scene("mix")
    _upload(ctx, g)
    ctx.set_option("smem_top", smem_top)
    try:
        assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), "closest")
        assert_hits_identical(ctx.trace(rtb200.ANY, g["random_rays"]), g["random_hits_any"].view(HIT).reshape(-1), "any")
    finally:
        ctx.set_option("smem_top", 0)


def _medium_scene():
    m = rtb200.Mesh().terrain(160, 100.0).icosphere(4, 25.0, (20.0, 30.0, -10.0)).sticks(400, 5, 150.0).finish(diffuse=(0.6, 0.7, 0.8))
    return m, m.arrays(), rtb200.FlatBVH.build(m)


def test_medium_scene_all_ray_classes_vs_oracle(ctx):
    """~56 K triangles with duplicated references, 320x240: primary, grazing-light shadow (order-dependent
    any-hit), device-generated diffuse rays, and the frame -- all against the live oracle."""
    import torch

    m, A, b = _medium_scene()
    assert b.duplicates > 0
    w, h = 320, 240
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=(-150.0, 25.0, 3.0))
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    ctx.set_params(params)
    sc = O.OracleScene(A, b.nodes, b.tri_indices)
    rays, gate = O.primary_rays(params, w, h)
    want, _ = sc.trace(0, rays)
    got = ctx.trace(rtb200.CLOSEST, rays)
    assert_hits_identical(got, want, "primary")
    srays, valid = O.shadow_rays(params, rays, want)
    v = valid.astype(bool)
    want_s, _ = sc.trace(1, srays[v])
    assert 0.05 < (want_s["idx"] >= 0).mean() < 0.95  # the grazing light really occludes some
    assert_hits_identical(ctx.trace(rtb200.ANY, srays[v]), want_s, "shadow any-hit")

    n = w * h
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.from_numpy(want.view(np.float32).reshape(-1, 4).copy()).cuda()
    d_out = torch.zeros((n * 4, 8), device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_out, d_cnt)
    ctx.synchronize()
    nd = int(d_cnt.item())
    assert nd == 4 * int((want["idx"] >= 0).sum())
    drays = d_out[:nd].cpu().numpy()
    assert np.allclose(np.linalg.norm(drays[:, 4:7], axis=1), 1.0, atol=1e-5)
    want_d, _ = sc.trace(0, drays)
    assert_hits_identical(ctx.trace(rtb200.CLOSEST, drays), want_d, "diffuse")

    img = ctx.render_frame(w, h)
    ref_img, _ = sc.render_frame(params, w, h)
    d = channel_diff(img, ref_img)
    assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() < 0.02


def test_gate_pretest_never_drops_a_hit(ctx):
    """the conservative approximate scene-gate pre-test only skips pixels the exact gate rejects anyway: primary hits
    stay bit-identical to the oracle from far, near, inside-the-box, grazing and axis-parallel camera poses"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = 128, 96
    poses = [dict(), dict(d_radius=1500.0), dict(d_radius=-150.0), dict(d_radius=-195.0), dict(d_beta=-0.78), dict(d_beta=0.7),
             dict(d_alpha=0.7853982, d_beta=-0.7853982)]
    for pose in poses:
        params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"], **pose)
        ctx.set_params(params)
        rays, gate = O.primary_rays(params, w, h)
        want, _ = sc.trace(0, rays)
        want[gate == 0] = (-1, rtb200.T_INIT, 0, 0)
        d_hits = torch.zeros((w * h, 4), device="cuda")
        ctx.primary_device(w, h, d_hits)
        ctx.synchronize()
        assert_hits_identical(d_hits.cpu().numpy().view(HIT).reshape(-1), want, f"pose {pose}")
        img = ctx.render_frame(w, h)
        ref_img, _ = sc.render_frame(params, w, h)
        assert channel_diff(img, ref_img).max() <= 1, pose


@pytest.mark.parametrize("seed,scale", [(1, 1.0), (2, 1e-2), (3, 1e3), (4, 37.0), (5, 0.3)])
def test_random_triangle_soup_fuzz(ctx, seed, scale):
    """random soups with degenerate (zero-area, repeated-vertex), axis-aligned and overlapping triangles at several scales,
    rays that start inside boxes, run along axes (zero direction components, signed zeros) or carry finite tmax:
    closest and any-hit stay bit-identical to the oracle (NaN/inf behaviour of the slab and Moller-Trumbore tests included)"""
    rng = np.random.default_rng(seed)
    nt = 1500
    v = (rng.normal(size=(nt, 3, 3)) * 10).astype(np.float32)
    v[:, 1] = v[:, 0] + (rng.normal(size=(nt, 3)) * 1.5).astype(np.float32)
    v[:, 2] = v[:, 0] + (rng.normal(size=(nt, 3)) * 1.5).astype(np.float32)
    v[:40, 2] = v[:40, 1]                      # zero-area: two equal vertices
    v[40:80, 1] = v[40:80, 0]; v[40:80, 2] = v[40:80, 0]  # a point
    v[80:200, :, 1] = np.round(v[80:200, :1, 1])  # axis-aligned (flat in y, on integer planes): zero-thickness boxes
    v[200:260, :, 0] = 0.0                       # lying in the x = 0 plane
    v[260:300] = v[300:340]                      # exact duplicates
    v *= np.float32(scale)
    verts = np.concatenate([v.reshape(-1, 3), np.ones((nt * 3, 1), dtype=np.float32)], axis=1)
    idx = np.arange(nt * 3, dtype=np.int32)
    m = rtb200.Mesh().set(verts, idx).finish()
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    sc = O.OracleScene(A, b.nodes, b.tri_indices)
    n = 40000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = (rng.normal(size=(n, 3)) * 12 * scale).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[:, 3] = rtb200.T_INIT
    k = n // 8
    rays[:k, 4:7] = 0; rays[:k, 4 + (np.arange(k) % 3)] = np.where(np.arange(k) % 2, 1.0, -1.0)   # along an axis
    rays[k:2 * k, 5] = 0.0                                         # one zero component
    rays[2 * k:2 * k + 500, 5] = -0.0                              # negative zero
    rays[3 * k:4 * k, 0:3] = np.round(rays[3 * k:4 * k, 0:3] / scale) * scale  # origins on integer planes (on box faces)
    rays[4 * k:5 * k, 3] = (rng.random(k) * 20 * scale + 1e-3).astype(np.float32)  # finite tmax
    for mode in (rtb200.CLOSEST, rtb200.ANY):
        want, _ = sc.trace(mode, rays)
        assert (want["idx"] >= 0).mean() > 0.02
        for sched in (0, 1):
            ctx.set_option("scheduler", sched)
            assert_hits_identical(ctx.trace(mode, rays), want, f"seed {seed} scale {scale} mode {mode} scheduler {sched}")
    ctx.set_option("scheduler", -1)


def test_sorted_trace_gives_the_same_hits(ctx):
    """rt_trace_sorted_device (coherence pre-pass) == rt_trace_device, ray for ray"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    rays = np.tile(g["random_rays"], (40, 1))
    rng = np.random.default_rng(0)
    rays = rays[rng.permutation(rays.shape[0])].copy()
    n = rays.shape[0]
    d_r = torch.from_numpy(rays).cuda()
    for mode in (rtb200.CLOSEST, rtb200.ANY):
        a, b = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
        ctx.trace_device(mode, n, d_r, a)
        ctx.trace_sorted_device(mode, n, d_r, b)
        ctx.synchronize()
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    # and small / ragged batches
    for m in (1, 33, 1000):
        a, b = torch.zeros((m, 4), device="cuda"), torch.zeros((m, 4), device="cuda")
        ctx.trace_device(rtb200.CLOSEST, m, d_r, a)
        ctx.trace_sorted_device(rtb200.CLOSEST, m, d_r, b)
        ctx.synchronize()
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_band_partition_covers_frame(ctx):
    """interleaved row bands (the multi-GPU partition) of the fused primary + frame kernels tile the frame exactly"""
    import torch

    g = load_scene("terrain12")
    _upload(ctx, g)
    w, h = 96, 64
    whole = torch.zeros((w * h, 4), device="cuda")
    ctx.primary_device(w, h, whole)
    img_whole = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ctx.render_frame_device(w, h, img_whole)
    for n_parts, band in ((2, 16), (3, 8), (4, 4)):
        acc = torch.full((w * h, 4), float("nan"), device="cuda")
        img = torch.full((h, w), -1, dtype=torch.int32, device="cuda")
        for part in range(n_parts):
            ctx.primary_device(w, h, acc, None, part=part, n_parts=n_parts, band_rows=band)
            ctx.render_frame_device(w, h, img, part=part, n_parts=n_parts, band_rows=band)
        ctx.synchronize()
        assert same_bits(acc.cpu().numpy(), whole.cpu().numpy())
        assert torch.equal(img, img_whole)


def test_fused_gather_store_matches_hits(ctx):
    """rt_primary_gather_device: the 4-byte/pixel index frame (the store-fused gather target; here a local IPC-capable
    allocation) equals the idx field of the hit records, for whole frames and for band partitions, with and without hits"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    w, h = 96, 64
    hits = torch.zeros((w * h, 4), device="cuda")
    ctx.primary_device(w, h, hits)
    ctx.synchronize()
    want = hits.view(torch.int32)[:, 0].cpu().numpy()
    ptr, handle = ctx.ipc_alloc(w * h * 4)
    assert len(handle) == 64
    try:
        for n_parts, band, with_hits in ((1, 4, True), (2, 16, True), (4, 8, False)):
            got = np.full(w * h, 12345, dtype=np.int32)
            hits2 = torch.zeros((w * h, 4), device="cuda") if with_hits else None
            for part in range(n_parts):
                ctx.primary_gather_device(w, h, hits2, ptr, part=part, n_parts=n_parts, band_rows=band)
            ctx.synchronize()
            ctx.memcpy_to_host(got, ptr, w * h * 4)
            assert np.array_equal(got, want)
            if with_hits:
                assert torch.equal(hits2.view(torch.int32), hits.view(torch.int32))
    finally:
        ctx.ipc_free(ptr)


def test_primary_host_entry_zero_copy_and_staged(ctx):
    """rt_primary (host buffer out): pinned destination (kernel stores straight to host memory), pinned but staged,
    and pageable destination all give the device result bit for bit"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    dev = torch.zeros((w * h, 4), device="cuda")
    ctx.primary_device(w, h, dev)
    ctx.synchronize()
    want = dev.cpu().view(torch.int32)
    pinned = torch.zeros((w * h, 4)).pin_memory()
    for zc in (1, 0):
        ctx.set_option("zero_copy", zc)
        pinned.zero_()
        ctx.primary(w, h, pinned)
        assert torch.equal(pinned.view(torch.int32), want)
    ctx.set_option("zero_copy", 1)
    pageable = ctx.primary(w, h)
    assert np.array_equal(pageable.view(np.int32).reshape(-1, 4), want.numpy())
    for tile_order in (1, 2, 3):  # work-item order is a scheduling choice only
        ctx.set_option("tile_order", tile_order)
        dev2 = torch.zeros((w * h, 4), device="cuda")
        ctx.primary_device(w, h, dev2)
        ctx.synchronize()
        assert torch.equal(dev2.view(torch.int32).cpu(), want)
    ctx.set_option("tile_order", 0)


def test_trace_host_pipeline_large_batch(ctx):
    """rt_trace on a batch that spans several upload/trace/download chunks, pageable and pinned (zero-copy) buffers,
    against rt_trace_device on the same rays; frame call with a pinned destination"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    reps = 300  # 2048 * 300 = 614 400 rays -> 3 chunks
    rays = np.tile(g["random_rays"], (reps, 1))
    rays[:, 0:3] += (np.arange(rays.shape[0], dtype=np.float32)[:, None] % 7) * 0.01
    n = rays.shape[0]
    d_r = torch.from_numpy(rays).cuda()
    d_h = torch.zeros((n, 4), device="cuda")
    for mode in (rtb200.CLOSEST, rtb200.ANY):
        ctx.trace_device(mode, n, d_r, d_h)
        ctx.synchronize()
        want = d_h.cpu().view(torch.int32)
        got = ctx.trace(mode, rays)  # pageable numpy in/out
        assert np.array_equal(got.view(np.int32).reshape(-1, 4), want.numpy())
        pr, ph = torch.from_numpy(rays).pin_memory(), torch.zeros((n, 4)).pin_memory()
        ctx.trace(mode, pr, ph)  # pinned in, zero-copy out
        assert torch.equal(ph.view(torch.int32), want)
    w, h = (int(v) for v in g["wh"])
    img_pinned = torch.zeros((h, w), dtype=torch.int32).pin_memory()
    ctx.render_frame(w, h, img_pinned)
    assert np.array_equal(img_pinned.numpy().view(np.uint32), ctx.render_frame(w, h))


@pytest.mark.parametrize("scheduler", [0, 1])
def test_both_schedulers_identical(ctx, scheduler):
    """batch and persistent-lanes schedulers run the same per-ray operation sequence"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    ctx.set_option("scheduler", scheduler)
    try:
        for refill, iexit in ((16, 8), (1, 0), (32, 16)):
            ctx.set_option("refill", refill)
            ctx.set_option("inner_exit", iexit)
            assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), "closest")
            assert_hits_identical(ctx.trace(rtb200.ANY, g["random_rays"]), g["random_hits_any"].view(HIT).reshape(-1), "any")
            w, h = (int(v) for v in g["wh"])
            d_hits = torch.zeros((w * h, 4), device="cuda")
            d_rays = torch.zeros((w * h, 8), device="cuda")
            d_sh = torch.zeros((w * h, 4), device="cuda")
            ctx.primary_device(w, h, d_hits, d_rays)
            ctx.shadow_device(w * h, d_rays, d_hits, d_sh)
            ctx.synchronize()
            want = g["primary_hits"].view(HIT).reshape(-1).copy()
            gate = g["primary_gate"].astype(bool)
            want[~gate] = (-1, rtb200.T_INIT, 0, 0)
            assert_hits_identical(d_hits.cpu().numpy().view(HIT).reshape(-1), want, "primary")
            v = g["shadow_valid"].astype(bool) & gate
            assert_hits_identical(d_sh.cpu().numpy().view(HIT).reshape(-1)[v], g["shadow_hits"].view(HIT).reshape(-1)[v], "shadow")
    finally:
        ctx.set_option("scheduler", -1)
        ctx.set_option("refill", 16)
        ctx.set_option("inner_exit", 8)


def test_ragged_and_edge_inputs(ctx):
    g = load_scene("ico2")
    _upload(ctx, g)
    # empty batch
    assert ctx.trace(rtb200.CLOSEST, np.zeros((0, 8), dtype=np.float32)).size == 0
    # ragged sizes around the warp width
    want = g["random_hits_closest"].view(HIT).reshape(-1)
    for n in (1, 31, 32, 33, 1000):
        assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"][:n].copy()), want[:n], f"n={n}")
    # frame sizes that are not multiples of the 8x4 warp tile
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    for w, h in ((37, 21), (8, 4), (1, 1)):
        params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"])
        ctx.set_params(params)
        ref_img, _ = sc.render_frame(params, w, h)
        assert channel_diff(ctx.render_frame(w, h), ref_img).max() <= 1
    # zero direction components (division by zero in ray_box -> inf/NaN handled like the oracle)
    rays = np.zeros((64, 8), dtype=np.float32)
    rays[:, 0:3] = np.random.default_rng(3).normal(size=(64, 3)) * 20
    rays[:, 3] = rtb200.T_INIT
    rays[:32, 4] = 1.0   # +x axis-aligned
    rays[32:, 5] = -1.0  # -y axis-aligned
    want, _ = sc.trace(0, rays)
    assert_hits_identical(ctx.trace(rtb200.CLOSEST, rays), want, "axis-aligned rays")


def test_root_leaf_scene(ctx):
    """test0 flattens to a single leaf node (root is a leaf): traversal starts in the leaf loop"""
    g = load_scene("test0")
    assert g["ref_nodes"].shape[0] == 1
    _upload(ctx, g)
    assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), "root leaf")


def test_error_behaviour(ctx):
    c2 = rtb200.Context(0)
    with pytest.raises(rtb200.RtError):  # no scene yet
        c2.trace(rtb200.CLOSEST, np.zeros((4, 8), dtype=np.float32))
    g = load_scene("ico2")
    bad = g["ref_nodes"].copy()
    bad.view(np.int32)[0, 9] = 10 ** 6  # root's right child out of range -> reference returns -1 for every ray
    c2.upload_scene(mesh_dict(g), bad, g["ref_tri_indices"])
    sc = O.OracleScene(mesh_dict(g), bad, g["ref_tri_indices"])
    want, _ = sc.trace(0, g["random_rays"])
    assert np.all(want["idx"] == -1)
    assert_hits_identical(c2.trace(rtb200.CLOSEST, g["random_rays"]), want, "poisoned root")
    cyc = g["ref_nodes"].copy()
    cyc.view(np.int32)[1, 8:10] = (0, 0) if cyc.view(np.int32)[1, 8] >= 0 else cyc.view(np.int32)[1, 8:10]
    if cyc.view(np.int32)[1, 8] == 0:
        with pytest.raises(rtb200.RtError, match="reachable twice"):
            c2.upload_scene(mesh_dict(g), cyc, g["ref_tri_indices"])
    idx = g["indices"].copy()
    idx[5] = 10 ** 7
    md = mesh_dict(g)
    md["indices"] = idx
    with pytest.raises(rtb200.RtError, match="outside"):
        c2.upload_scene(md, g["ref_nodes"], g["ref_tri_indices"])
    with pytest.raises(rtb200.RtError):  # no params
        c3 = rtb200.Context(0)
        c3.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
        c3.render_frame(8, 8)
    c2.close()


def test_scene_blob_adopt_roundtrip(ctx):
    """the packed scene is one position-independent device buffer: copy it (as NCCL would) and adopt it"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    ptr, nbytes = ctx.scene_blob()
    assert ptr and nbytes == ctx.scene_info()["blob_bytes"]
    clone = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    ctx.copy_scene_blob(clone, nbytes)
    moved = clone.clone()  # a second copy at a different address: the blob is position independent
    del clone
    c2 = rtb200.Context(0)
    c2.adopt_scene_blob(moved.data_ptr(), nbytes)
    assert_hits_identical(c2.trace(rtb200.ANY, g["random_rays"]), g["random_hits_any"].view(HIT).reshape(-1), "adopted blob")
    # a header whose sections do not fit the allocation is rejected before any pointer is derived from it
    hdr = moved[:256].cpu().numpy().copy()
    for field_offset, value in ((64, nbytes), (72, 2**40), (20, 2**30), (16, 2**30)):  # off_pairs, off_tris, num_pairs, root_ref
        bad = moved.clone()
        h2 = hdr.copy()
        h2[field_offset:field_offset + 8 if field_offset >= 64 else field_offset + 4] = np.frombuffer(
            (np.uint64(value) if field_offset >= 64 else np.int32(value)).tobytes(), dtype=np.uint8)
        bad[:256] = torch.from_numpy(h2).cuda()
        with pytest.raises(rtb200.RtError, match="rejected|mismatch"):
            c2.adopt_scene_blob(bad.data_ptr(), nbytes)
    # ... and so are CONTENTS that point outside their sections: a child reference past the pair section, a
    # triangle id past the mesh, a missing leaf terminator on the sentinel -- checked by one kernel when a blob is adopted
    off_pairs = int(np.frombuffer(hdr[64:72].tobytes(), dtype=np.uint64)[0])
    off_tris = int(np.frombuffer(hdr[72:80].tobytes(), dtype=np.uint64)[0])
    num_tris = int(np.frombuffer(hdr[24:28].tobytes(), dtype=np.int32)[0])
    for byte_at, value in ((off_pairs + 16 + 8, 2**30),                    # pair 0, child 0 reference
                           (off_pairs + 64 * 3 + 48 + 8, -(2**30)),         # pair 3, child 1: leaf reference far past the triangles
                           (off_tris + 48 * 2 + 12, 3 * 10**8),             # triangle 2: id beyond the mesh
                           (off_tris + 48 * num_tris + 16 + 12, 0)):        # the sentinel's `last` flag
        bad = moved.clone()
        bad[byte_at:byte_at + 4] = torch.from_numpy(np.frombuffer(np.int32(value).tobytes(), dtype=np.uint8).copy()).cuda()
        with pytest.raises(rtb200.RtError, match="rejected"):
            c2.adopt_scene_blob(bad.data_ptr(), nbytes)
    c2.adopt_scene_blob(moved.data_ptr(), nbytes)  # the intact copy is still accepted afterwards
    with pytest.raises(rtb200.RtError):
        c2.adopt_scene_blob(moved.data_ptr(), nbytes - 256)
    shifted = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    shifted[16:16 + nbytes] = moved
    with pytest.raises(rtb200.RtError, match="aligned"):
        c2.adopt_scene_blob(shifted.data_ptr() + 16, nbytes)
    c2.close()


def test_full_size_properties(ctx):
    """BASELINE config 2 at full size (1920x1080, ~1M-triangle terrain): size-independent properties.
    (a) a 1/64 subsample of the pixels is bit-identical to the oracle; (b) closest t <= any-hit t on the same
    rays and the occlusion booleans agree; (c) tracing with tmax = t_closest + ulp returns the same hit, with
    tmax = t_closest returns a different/no hit (strict '<'); (d) re-running is idempotent (bitwise)."""
    import torch

    m = rtb200.Mesh().terrain(707, 100.0).finish()
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    w, h = 1920, 1080
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    ctx.set_params(params)
    n = w * h
    d_hits = torch.zeros((n, 4), device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    ctx.primary_device(w, h, d_hits, d_rays)
    ctx.synchronize()
    hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
    rays = d_rays.cpu().numpy()
    assert 0.30 < (hits["idx"] >= 0).mean() < 0.40
    d_hits2 = torch.zeros((n, 4), device="cuda")
    ctx.primary_device(w, h, d_hits2)
    ctx.synchronize()
    assert torch.equal(d_hits.view(torch.int32), d_hits2.view(torch.int32))  # (d)

    sc = O.OracleScene(A, b.nodes, b.tri_indices)
    sub = np.arange(0, n, 64)
    orays, gate = O.primary_rays(params, w, h)
    assert same_bits(orays[sub], rays[sub])
    want, _ = sc.trace(0, orays[sub])
    want[gate[sub] == 0] = (-1, rtb200.T_INIT, 0, 0)
    assert_hits_identical(hits[sub], want, "1/64 subsample")  # (a)

    d_any = torch.zeros((n, 4), device="cuda")
    ctx.trace_device(rtb200.ANY, n, d_rays, d_any)
    ctx.synchronize()
    anyh = d_any.cpu().numpy().view(HIT).reshape(-1)
    g = gate.astype(bool)
    assert np.array_equal(anyh["idx"][g] >= 0, hits["idx"][g] >= 0)  # (b)
    hit = g & (hits["idx"] >= 0)
    assert np.all(anyh["t"][hit] >= hits["t"][hit])

    sel = np.flatnonzero(hit)[::97]
    r2 = rays[sel].copy()
    r2[:, 3] = np.nextafter(hits["t"][sel], np.float32(np.inf))
    h2 = ctx.trace(rtb200.CLOSEST, r2)
    assert np.array_equal(h2["idx"], hits["idx"][sel]) and same_bits(h2["t"], hits["t"][sel])  # (c)
    r2[:, 3] = hits["t"][sel]
    h3 = ctx.trace(rtb200.CLOSEST, r2)
    assert np.all((h3["idx"] != hits["idx"][sel]) | (h3["t"] < hits["t"][sel]))


@pytest.mark.parametrize("overlap", [1, 0])
@pytest.mark.parametrize("pinned", [True, False])
def test_pipelined_frames_equal_frame_by_frame(ctx, pinned, overlap):
    """rt_render_frame_begin/_end with several frames in flight (each with its own Params block, set while earlier
    frames are still running) deliver exactly the frames rt_render_frame delivers one by one"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    ctx.set_option("overlap_frames", overlap)  # 1: every frame slot on its own stream (frames overlap on the GPU); 0: context stream
    w, h = (int(v) for v in g["wh"])
    frames = []
    for k in range(6):  # orbit: rotate a, b, c, campos about the y axis
        p = g["params"].copy().reshape(8, 4)
        ang = 0.2 * k
        cs, sn = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        for row in range(4):
            x, z = p[row, 0], p[row, 2]
            p[row, 0], p[row, 2] = cs * x + sn * z, -sn * x + cs * z
        frames.append(p.reshape(-1))
    want = []
    for p in frames:
        ctx.set_params(p)
        want.append(ctx.render_frame(w, h).copy())
    assert any(not np.array_equal(want[0], x) for x in want[1:])
    slots = 3
    bufs = [(torch.zeros((h, w), dtype=torch.int32).pin_memory() if pinned else np.zeros((h, w), dtype=np.uint32)) for _ in range(slots)]
    got = [None] * len(frames)
    for k, p in enumerate(frames):
        if k >= slots:
            ctx.render_frame_end((k - slots) % slots)
            b = bufs[(k - slots) % slots]
            got[k - slots] = (b.numpy().view(np.uint32) if pinned else b).copy()
        ctx.set_params(p)
        ctx.render_frame_begin(w, h, bufs[k % slots], k % slots)
    for k in range(len(frames) - slots, len(frames)):
        ctx.render_frame_end(k % slots)
        b = bufs[k % slots]
        got[k] = (b.numpy().view(np.uint32) if pinned else b).copy()
    ctx.set_option("overlap_frames", 1)
    for k in range(len(frames)):
        assert np.array_equal(got[k], want[k]), f"frame {k}"
    ctx.set_params(g["params"])


def test_pipelined_frame_slot_errors(ctx):
    g = load_scene("test0")
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    out = np.zeros((h, w), dtype=np.uint32)
    with pytest.raises(rtb200.RtError, match="no frame in flight"):
        ctx.render_frame_end(0)
    with pytest.raises(rtb200.RtError, match="outside"):
        ctx.render_frame_begin(w, h, out, 4)
    ctx.render_frame_begin(w, h, out, 1)
    with pytest.raises(rtb200.RtError, match="still has a frame in flight"):
        ctx.render_frame_begin(w, h, out, 1)
    ctx.render_frame_end(1)
    assert np.array_equal(out, g["frame"])


def test_cpp_host_driver_pipelined_frames_match(tmp_path):
    """host/rt_demo (RayTracer::setupCL / initRayTraceFromMesh / updateCamera / raytrace_gpgpu and the pipelined
    raytrace_gpgpu_begin/_end pair, all through the C ABI): the orbit rendered frame by frame and with two frames in
    flight gives identical frame checksums; the PPM it writes has the right size"""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(rtb200.device.LIB_PATH), "..", "host", "rt_demo")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is missing (build it with __graft_entry__.build())")
    ppm = tmp_path / "f.ppm"
    r = subprocess.run([exe, "terrain:40", "320", "200", "12", str(ppm)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "(identical)" in r.stdout
    data = ppm.read_bytes()
    assert data.startswith(b"P6\n320 200\n255\n") and len(data) == len(b"P6\n320 200\n255\n") + 320 * 200 * 3
