"""GPU tests of the round-2 device paths, through the C ABI: the fused primary + shadow launch, row assembly of
4-byte/pixel frames (128 / 512-byte stores into host or peer memory), the device-side ray counter, and the one-process
multi-GPU group. Bar as everywhere: bit-identical to the two-pass path, to the golden fixtures and to the live oracle."""
import numpy as np
import pytest
from conftest import SCENES, assert_hits_identical, load_scene, mesh_dict, gpu_context

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


def expected_vis(primary_hits, shadow_hits):
    """vis = -1 | 3*triId + occluded, occluded = shadow idx >= 0 && shadow t > 0.025 (vR.cl:1444-1449)"""
    occl = (shadow_hits["idx"] >= 0) & (shadow_hits["t"] > np.float32(0.025))
    return np.where(primary_hits["idx"] >= 0, primary_hits["idx"] + occl.astype(np.int32), -1).astype(np.int32)


@pytest.mark.parametrize("name", SCENES)
def test_fused_primary_shadow_vs_golden(ctx, name):
    """one launch == rt_primary_device followed by rt_shadow_device == the committed oracle outputs"""
    import torch

    g = load_scene(name)
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    n = w * h
    d_hits = torch.full((n, 4), -7.0, device="cuda")
    d_sh = torch.full((n, 4), -7.0, device="cuda")
    d_vis = torch.full((h, w), -7, dtype=torch.int32, device="cuda")
    ctx.primary_shadow_device(w, h, d_hits, d_sh, d_vis)
    ctx.synchronize()
    hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
    sh = d_sh.cpu().numpy().view(HIT).reshape(-1)
    gate = g["primary_gate"].astype(bool)
    want = g["primary_hits"].view(HIT).reshape(-1).copy()
    want[~gate] = (-1, rtb200.T_INIT, 0, 0)
    assert_hits_identical(hits, want, f"{name} fused: primary")
    v = g["shadow_valid"].astype(bool) & gate
    assert_hits_identical(sh[v], g["shadow_hits"].view(HIT).reshape(-1)[v], f"{name} fused: shadow")
    miss = want["idx"] < 0
    assert np.all(sh["idx"][miss] == -1) and np.all(sh["t"][miss] == rtb200.T_INIT)
    assert np.array_equal(d_vis.cpu().numpy().reshape(-1), expected_vis(hits, sh))
    # and the two-pass path says the same
    d_h2 = torch.zeros((n, 4), device="cuda")
    d_r2 = torch.zeros((n, 8), device="cuda")
    d_s2 = torch.zeros((n, 4), device="cuda")
    ctx.primary_device(w, h, d_h2, d_r2)
    ctx.shadow_device(n, d_r2, d_h2, d_s2)
    ctx.synchronize()
    assert torch.equal(d_h2.view(torch.int32), d_hits.view(torch.int32))
    assert torch.equal(d_s2.view(torch.int32), d_sh.view(torch.int32))


def test_fused_pass_grazing_light_vs_live_oracle(ctx):
    """order-dependent any-hit results (grazing light: many occluders per ray) through the fused launch"""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = 200, 120
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"], light_pos=(-150.0, 25.0, 3.0))
    ctx.set_params(params)
    d_hits = torch.zeros((w * h, 4), device="cuda")
    d_sh = torch.zeros((w * h, 4), device="cuda")
    ctx.primary_shadow_device(w, h, d_hits, d_sh, None)
    ctx.synchronize()
    sc = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    rays, gate = O.primary_rays(params, w, h)
    want, _ = sc.trace(0, rays)
    want[~gate.astype(bool)] = (-1, rtb200.T_INIT, 0, 0)
    hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
    assert_hits_identical(hits, want, "primary")
    srays, valid = O.shadow_rays(params, rays, want)
    v = valid.astype(bool) & gate.astype(bool)
    want_s, _ = sc.trace(1, np.ascontiguousarray(srays[v]))
    assert (want_s["idx"] >= 0).mean() > 0.05
    assert_hits_identical(d_sh.cpu().numpy().view(HIT).reshape(-1)[v], want_s, "shadow")


@pytest.mark.parametrize("wh", [(96, 52), (203, 77), (640, 36), (1000, 8)])
@pytest.mark.parametrize("group", [0, 2, 4])
def test_row_assembly_gives_the_same_frames(ctx, wh, group):
    """frames stored through the row assembler (stage + last-arriver copy-out in 128/512-byte rows) equal the directly
    stored ones: device, pinned-host and band-partitioned destinations; ragged widths (w % 32 != 0, w % 4 != 0)"""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = wh
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"])
    ctx.set_params(params)
    ctx.set_option("store_group", 0)
    ref_idx = torch.full((h, w), -9, dtype=torch.int32, device="cuda")
    ref_vis = torch.full((h, w), -9, dtype=torch.int32, device="cuda")
    ref_img = torch.full((h, w), -9, dtype=torch.int32, device="cuda")
    ctx.primary_gather_device(w, h, None, ref_idx)
    ctx.primary_shadow_device(w, h, None, None, ref_vis)
    ctx.render_frame_device(w, h, ref_img)
    ctx.synchronize()
    try:
        ctx.set_option("store_group", group)
        for dst in ("device", "pinned"):
            mk = (lambda: torch.full((h, w), -5, dtype=torch.int32, device="cuda")) if dst == "device" else \
                 (lambda: torch.full((h, w), -5, dtype=torch.int32).pin_memory())
            idx, vis, img = mk(), mk(), mk()
            ctx.primary_gather_device(w, h, None, idx)
            ctx.primary_shadow_device(w, h, None, None, vis)
            ctx.render_frame_device(w, h, img)
            ctx.synchronize()
            assert torch.equal(idx.cpu(), ref_idx.cpu()), f"idx frame, {dst}"
            assert torch.equal(vis.cpu(), ref_vis.cpu()), f"vis frame, {dst}"
            assert torch.equal(img.cpu(), ref_img.cpu()), f"shaded frame, {dst}"
        # three interleaved 4-row bands, each "rank" writing into the same frame
        idx = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        for part in range(3):
            ctx.primary_gather_device(w, h, None, idx, part=part, n_parts=3, band_rows=4)
        ctx.synchronize()
        assert torch.equal(idx, ref_idx)
    finally:
        ctx.set_option("store_group", -1)


def test_auto_row_assembly_for_host_frames_and_host_entry_point(ctx):
    """auto mode: a pinned destination is row-assembled, rt_primary_shadow (host buffer, pinned or pageable) delivers the
    same visibility frame as the device entry point"""
    import torch

    g = load_scene("terrain12")
    _upload(ctx, g)
    w, h = 320, 200
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"])
    ctx.set_params(params)
    d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ctx.primary_shadow_device(w, h, None, None, d_vis)
    ctx.synchronize()
    pinned = torch.zeros((h, w), dtype=torch.int32).pin_memory()
    ctx.primary_shadow(w, h, pinned)
    pageable = ctx.primary_shadow(w, h)
    assert np.array_equal(pinned.numpy(), d_vis.cpu().numpy()) and np.array_equal(pageable, d_vis.cpu().numpy())
    assert (d_vis >= 0).any() and (d_vis < 0).any()


def test_rays_traced_counter_counts_traversals(ctx):
    """RT_CNT_RAYS_TRACED is counted by the kernels: gate-passing primaries (+ one shadow ray per hit), caller rays"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    w, h = (int(v) for v in g["wh"])
    gate = int(g["primary_gate"].astype(bool).sum())
    want = g["primary_hits"].view(HIT).reshape(-1)
    nhit = int(((want["idx"] >= 0) & g["primary_gate"].astype(bool)).sum())
    d_hits = torch.zeros((w * h, 4), device="cuda")
    ctx.reset_counters()
    ctx.primary_device(w, h, d_hits)
    assert ctx.counters()["rays_traced"] == gate
    ctx.reset_counters()
    ctx.primary_shadow_device(w, h, d_hits, None, None)
    assert ctx.counters()["rays_traced"] == gate + nhit
    ctx.reset_counters()
    ctx.trace(rtb200.CLOSEST, g["random_rays"])
    ctx.trace(rtb200.ANY, g["random_rays"])
    assert ctx.counters()["rays_traced"] == 2 * g["random_rays"].shape[0]
    ctx.reset_counters()
    ctx.render_frame(w, h)
    c = ctx.counters()["rays_traced"]
    assert gate + nhit <= c <= 6 * gate  # <= 3 path segments of closest + shadow


def test_calls_leave_the_current_device_alone(ctx):
    """every entry point restores the calling thread's current CUDA device (torch and the library share a process)"""
    import torch

    g = load_scene("test0")
    before = torch.cuda.current_device()
    _upload(ctx, g)
    ctx.trace(rtb200.CLOSEST, g["random_rays"])
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() > 1:
        c1 = rtb200.Context(1)
        try:
            c1.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
            got = c1.trace(rtb200.CLOSEST, g["random_rays"])
            assert_hits_identical(got, g["random_hits_closest"].view(HIT).reshape(-1), "device 1")
            assert torch.cuda.current_device() == before
            x = torch.zeros(4, device="cuda")  # still lands on the device torch had selected
            assert x.device.index == before
        finally:
            c1.close()


# ---- one process, N GPUs (rt_group) ----------------------------------------------------------------------------------

def _group_sizes():
    import torch

    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [k for k in (1, 2, 4, 8) if k <= n] or [1]


@pytest.mark.parametrize("pinned", [True, False])
def test_group_frames_equal_single_context(ctx, pinned):
    """rt_render_frame_tiled / rt_primary_tiled over 1..N GPUs of this box == the single-context frames, bit for bit
    (N = 1 runs everywhere; N >= 2 needs a multi-GPU box: NCCL broadcast of the blob, peer / host stores of the bands)"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    w, h = 328, 204  # w % 32 != 0: ragged row groups; 51 tile rows: bands of unequal count
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"], light_pos=(-150.0, 25.0, 3.0))
    ctx.set_params(params)
    want_img = ctx.render_frame(w, h)
    want_vis = ctx.primary_shadow(w, h)
    d_idx = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ctx.primary_gather_device(w, h, None, d_idx)
    ctx.synchronize()
    want_idx = d_idx.cpu().numpy()
    for n in _group_sizes():
        grp = rtb200.Group(n)
        try:
            grp.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
            grp.set_params(params)
            for band_rows in (4, 16):
                grp.set_option("band_rows", band_rows)
                mk32 = (lambda: torch.full((h, w), -3, dtype=torch.int32).pin_memory()) if pinned else (lambda: np.full((h, w), -3, dtype=np.int32))
                img, vis, idx = mk32(), mk32(), mk32()
                grp.render_frame(w, h, img)
                grp.primary(w, h, True, vis)
                grp.primary(w, h, False, idx)
                as_np = (lambda t: t.numpy()) if pinned else (lambda a: a)
                assert np.array_equal(as_np(img).view(np.uint32), want_img), f"{n} GPUs, bands of {band_rows}: shaded frame"
                assert np.array_equal(as_np(vis), want_vis), f"{n} GPUs: visibility frame"
                assert np.array_equal(as_np(idx), want_idx), f"{n} GPUs: index frame"
                st = grp.stats()
                assert st["zero_copy"] == (1.0 if pinned else 0.0)
                assert len(st["rank_kernel_ms"]) == n and all(ms > 0 for ms in st["rank_kernel_ms"])
            if n > 1:
                assert grp.stats()["broadcast_ms"] > 0 and grp.stats()["blob_bytes"] > 0
                grp.set_option("broadcast", 1)  # peer copies instead of NCCL: same scene on every GPU
                grp.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
                assert np.array_equal(grp.render_frame(w, h), want_img)
        finally:
            grp.close()


def test_group_host_frame_at_an_odd_offset(ctx):
    """a page-locked frame the kernels cannot store words into (2 bytes off) goes through the gather frame + copy"""
    import torch

    g = load_scene("mix")
    _upload(ctx, g)
    w, h = 200, 120
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"])
    ctx.set_params(params)
    want = ctx.render_frame(w, h)
    for n in _group_sizes():
        grp = rtb200.Group(n)
        try:
            grp.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
            grp.set_params(params)
            store = torch.zeros(4 * w * h + 16, dtype=torch.uint8).pin_memory()
            raw = store.numpy()
            fr = np.frombuffer(raw.data, dtype=np.uint32, count=w * h, offset=2).reshape(h, w)
            grp.render_frame(w, h, fr)
            assert np.array_equal(fr, want), f"{n} GPUs"
            assert grp.stats()["zero_copy"] == 0.0
            assert (raw[:2] == 0).all() and (raw[2 + 4 * w * h:] == 0).all()
        finally:
            grp.close()


def test_group_errors():
    import torch

    with pytest.raises(rtb200.RtError):
        rtb200.Group(torch.cuda.device_count() + 1)
    grp = rtb200.Group(1)
    try:
        with pytest.raises(rtb200.RtError, match="no scene"):
            grp.render_frame(64, 64)
        with pytest.raises(rtb200.RtError, match="band_rows"):
            grp.set_option("band_rows", 6)
    finally:
        grp.close()
