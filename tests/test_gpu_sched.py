"""Temporal tile scheduling (csrc/kernels.cuh "tile scheduler"): hints recorded by one launch reorder and split the tiles of
the next launch of the same frame geometry. They must never change a result: every pixel is traced exactly once, by
whichever warp, with the same arithmetic. The thresholds are forced low here so that small test frames put a large share
of their tiles on the heavy and split lists."""
import numpy as np
import pytest
from conftest import gpu_context, load_scene, mesh_dict

import rtb200

pytestmark = pytest.mark.gpu
HIT = rtb200.HIT_DTYPE
GRAZING = (-150.0, 25.0, 3.0)


@pytest.fixture(scope="module")
def ctx():
    c = gpu_context()
    yield c
    c.close()


def _scene(ctx):
    m = rtb200.Mesh().terrain(96, 100.0).icosphere(3, 22.0, (15.0, 28.0, -20.0)).finish(diffuse=(0.6, 0.7, 0.8))
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    return A


def _force_hints(ctx, on=True):
    ctx.set_option("tile_hints", 1 if on else 0)
    ctx.set_option("hint_heavy_pct", 30)
    ctx.set_option("hint_split_pct", 8)
    ctx.set_option("hint_keep_pct", 2)


@pytest.mark.parametrize("wh", [(328, 204), (640, 360), (97, 53)])
@pytest.mark.parametrize("parts", [(0, 1, 4), (1, 3, 8), (2, 3, 4)])
def test_hints_never_change_a_frame(ctx, wh, parts):
    """primary hits, index frame (with and without row assembly), fused primary+shadow and the shaded frame: eight
    consecutive launches with hints == the launch without hints, bit for bit; and the hints do engage."""
    import torch

    w, h = wh
    part, n_parts, band_rows = parts
    A = _scene(ctx)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=GRAZING)
    ctx.set_params(params)
    kw = dict(part=part, n_parts=n_parts, band_rows=band_rows)

    def run_all(store_group):
        ctx.set_option("store_group", store_group)
        hits = torch.full((w * h, 4), -7.0, device="cuda")
        sh = torch.full((w * h, 4), -7.0, device="cuda")
        idx = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        vis = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        img = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        ctx.primary_device(w, h, hits, **kw)
        ctx.primary_gather_device(w, h, None, idx, **kw)
        ctx.primary_shadow_device(w, h, None, sh, vis, **kw)
        ctx.render_frame_device(w, h, img, **kw)
        ctx.synchronize()
        return [t.cpu().numpy().copy() for t in (hits, sh, idx, vis, img)]

    _force_hints(ctx, False)
    want = run_all(0)
    assert (want[0].view(HIT)["idx"] >= 0).any()
    _force_hints(ctx, True)
    engaged = {"split_rows": 0, "heavy_tiles": 0}
    for it in range(8):
        got = run_all((0, 2, 4)[it % 3])
        for name, g, wnt in zip(("hits", "shadow hits", "index frame", "visibility frame", "shaded frame"), got, want):
            assert np.array_equal(g.view(np.uint32), wnt.view(np.uint32)), f"launch {it}: {name} changed with tile hints on"
        st = ctx.tile_hint_stats()
        for k in engaged:
            engaged[k] = max(engaged[k], st[k])
    assert engaged["split_rows"] > 0 and engaged["heavy_tiles"] > 0, engaged
    _force_hints(ctx, False)
    ctx.set_option("store_group", -1)
    ctx.set_option("tile_hints", 1)
    for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
        ctx.set_option(k, v)


def test_hints_follow_a_moving_camera_and_a_new_scene(ctx):
    """hints recorded for one camera pose / scene are only hints for the next: orbiting the camera and replacing the scene
    between launches still gives the frames of a hint-free context"""
    w, h = 320, 200
    _scene(ctx)
    ref = rtb200.Context(0)
    ref.set_option("tile_hints", 0)
    try:
        _force_hints(ctx, True)
        for step in range(6):
            if step == 3:  # a different scene under the same frame geometry
                g = load_scene("cubes2")
                for c in (ctx, ref):
                    c.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
                lo, hi = g["aabb_min"], g["aabb_max"]
            elif step == 0:
                m = rtb200.Mesh().terrain(96, 100.0).icosphere(3, 22.0, (15.0, 28.0, -20.0)).finish(diffuse=(0.6, 0.7, 0.8))
                A = m.arrays()
                b = rtb200.FlatBVH.build(m)
                ref.upload_scene(A, b.nodes, b.tri_indices)
                lo, hi = A["aabb_min"], A["aabb_max"]
            params, _ = rtb200.camera_params(w, h, lo, hi, d_alpha=0.35 * step, d_beta=-0.05 * step)
            for c in (ctx, ref):
                c.set_params(params)
            assert np.array_equal(ctx.render_frame(w, h), ref.render_frame(w, h)), f"step {step}"
            assert np.array_equal(ctx.primary_shadow(w, h), ref.primary_shadow(w, h)), f"step {step}"
    finally:
        ref.close()
        ctx.set_option("tile_hints", 1)
        for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
            ctx.set_option(k, v)


def test_frames_in_flight_keep_their_own_hints(ctx):
    """rt_render_frame_begin/_end with two slots in flight: each slot has its own hint buffers"""
    import torch

    w, h = 384, 216
    A = _scene(ctx)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"])
    ctx.set_params(params)
    ctx.set_option("tile_hints", 0)
    want = ctx.render_frame(w, h)
    _force_hints(ctx, True)
    bufs = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
    try:
        for it in range(10):
            s = it % 2
            if it >= 2:
                ctx.render_frame_end(s)
                assert np.array_equal(bufs[s].numpy().view(np.uint32), want), f"frame {it - 2}"
            bufs[s].zero_()
            ctx.render_frame_begin(w, h, bufs[s], s)
        for s in range(2):
            ctx.render_frame_end(s)
            assert np.array_equal(bufs[s].numpy().view(np.uint32), want)
    finally:
        ctx.set_option("tile_hints", 1)
        for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
            ctx.set_option(k, v)


@pytest.mark.parametrize("pose", [dict(), dict(d_radius=-120.0), dict(d_radius=-195.0), dict(d_radius=400.0, d_alpha=0.7, d_beta=0.3),
                                  dict(d_alpha=2.0, d_beta=-0.6), dict(d_radius=-150.0, d_beta=-0.7)],
                         ids=["default", "close", "inside_the_box", "far_rotated", "rotated", "low_and_close"])
@pytest.mark.parametrize("parts", [(0, 1, 4), (1, 3, 8)])
def test_scene_box_culling_never_changes_a_result(ctx, pose, parts):
    """the tile queue enumerates only the screen rectangle of the scene box and fills the rest (cull_setup): every output of
    every camera-ray entry point, and the number of rays traced, equal the unculled launch -- camera far, close, inside the
    box (no rectangle: whole frame), with and without row assembly, with the tile hints on"""
    import torch

    w, h = 424, 244
    part, n_parts, band_rows = parts
    A = _scene(ctx)
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=GRAZING, **pose)
    ctx.set_params(params)
    kw = dict(part=part, n_parts=n_parts, band_rows=band_rows)

    def run_all(store_group):
        ctx.set_option("store_group", store_group)
        hits = torch.full((w * h, 4), -7.0, device="cuda")
        sh = torch.full((w * h, 4), -7.0, device="cuda")
        idx = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        vis = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        img = torch.full((h, w), -5, dtype=torch.int32, device="cuda")
        ctx.reset_counters()
        ctx.primary_device(w, h, hits, **kw)
        ctx.primary_gather_device(w, h, None, idx, **kw)
        ctx.primary_shadow_device(w, h, None, sh, vis, **kw)
        ctx.render_frame_device(w, h, img, **kw)
        ctx.synchronize()
        return [t.cpu().numpy().copy() for t in (hits, sh, idx, vis, img)], ctx.counters()["rays_traced"]

    try:
        ctx.set_option("gate_cull", 0)
        ctx.set_option("tile_hints", 0)
        want, want_rays = run_all(0)
        ctx.set_option("gate_cull", 1)
        _force_hints(ctx, True)
        for it in range(6):
            got, rays = run_all((0, 2, 4)[it % 3])
            for name, g, wnt in zip(("hits", "shadow hits", "index frame", "visibility frame", "shaded frame"), got, want):
                assert np.array_equal(g.view(np.uint32), wnt.view(np.uint32)), f"launch {it}: {name} changed with the culled queue"
            assert rays == want_rays, (rays, want_rays)
    finally:
        ctx.set_option("store_group", -1)
        ctx.set_option("gate_cull", 1)
        ctx.set_option("tile_hints", 1)
        for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
            ctx.set_option(k, v)


def test_culled_queue_follows_a_moving_camera(ctx):
    """hints recorded under one rectangle, a different rectangle in the next launch (the camera zooms and orbits): tiles
    change sides between launches and still every frame equals the unculled, hint-free one"""
    w, h = 384, 216
    A = _scene(ctx)
    ref = rtb200.Context(0)
    ref.set_option("tile_hints", 0)
    ref.set_option("gate_cull", 0)
    m = rtb200.Mesh().terrain(96, 100.0).icosphere(3, 22.0, (15.0, 28.0, -20.0)).finish(diffuse=(0.6, 0.7, 0.8))
    b = rtb200.FlatBVH.build(m)
    ref.upload_scene(m.arrays(), b.nodes, b.tri_indices)
    try:
        _force_hints(ctx, True)
        for step in range(8):
            params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], d_radius=-40.0 * step + 100.0, d_alpha=0.25 * step, d_beta=0.04 * step)
            for c in (ctx, ref):
                c.set_params(params)
            assert np.array_equal(ctx.render_frame(w, h), ref.render_frame(w, h)), f"step {step}"
            assert np.array_equal(ctx.primary_shadow(w, h), ref.primary_shadow(w, h)), f"step {step}"
    finally:
        ref.close()
        ctx.set_option("tile_hints", 1)
        for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
            ctx.set_option(k, v)


def test_inner_exit_instances_give_identical_results(ctx):
    """the batch kernels exist in two instances of the traversal loop (traverse.cuh INNER_EXIT: leave the inner loop early
    while others wait on a leaf; picked per scene by its size against L2). Forcing either on the same scene must not change
    one bit of any entry point: rays from buffers (closest / any), camera rays, shadow rays, the fused pass, frames."""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    w, h = 328, 204
    params, _ = rtb200.camera_params(w, h, g["aabb_min"], g["aabb_max"], light_pos=GRAZING)
    ctx.set_params(params)
    n = w * h
    ctx.set_option("scheduler", 0)  # the batch kernels (ray buffers default to the lanes kernel)

    def run_all():
        hits = torch.zeros((n, 4), device="cuda")
        rays = torch.zeros((n, 8), device="cuda")
        sh = torch.zeros((n, 4), device="cuda")
        anyh = torch.zeros((n, 4), device="cuda")
        clo = torch.zeros((n, 4), device="cuda")
        vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        ctx.primary_device(w, h, hits, rays)
        ctx.shadow_device(n, rays, hits, sh)
        ctx.trace_device(rtb200.ANY, n, rays, anyh)
        ctx.trace_device(rtb200.CLOSEST, n, rays, clo)
        ctx.primary_shadow_device(w, h, None, None, vis)
        ctx.render_frame_device(w, h, img)
        ctx.synchronize()
        return [t.cpu().numpy().copy() for t in (hits, sh, anyh, clo, vis, img)]

    try:
        ctx.set_option("inner_exit_batch", 0)
        want = run_all()
        ctx.set_option("inner_exit_batch", 1)
        for _ in range(3):
            got = run_all()
            for name, a, b in zip(("primary", "shadow", "any-hit buffer", "closest buffer", "visibility frame", "shaded frame"), got, want):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{name} differs between the two loop instances"
        golden = g["random_hits_closest"].view(HIT).reshape(-1)
        from conftest import assert_hits_identical
        assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), golden, "INNER_EXIT instance vs golden")
    finally:
        ctx.set_option("inner_exit_batch", -1)
        ctx.set_option("scheduler", -1)
