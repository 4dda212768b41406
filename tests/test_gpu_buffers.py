"""The C ABI takes plain pointers, so what it does around the caller's buffers is part of the contract (include/rtb200.h,
"Alignment"): (1) no pass stores one byte outside the buffer it was given -- tile padding, culled regions filled by `fill`
items, split rows, assembled 128/512-byte rows and band parts all write exact extents; (2) a device buffer the kernels
could not address with their 16-byte vectors is refused up front instead of faulting (a fault is a sticky CUDA error that
ends the process' CUDA work); (3) host buffers of any alignment work -- oddly placed ones go through the copy."""
import numpy as np
import pytest
from conftest import assert_hits_identical, gpu_context, load_scene, mesh_dict

import rtb200

pytestmark = pytest.mark.gpu
HIT = rtb200.HIT_DTYPE
GRAZING = (-150.0, 25.0, 3.0)
GUARD_WORDS = 4096  # 16 KiB of pattern either side of every buffer
PATTERN = 0x5A5A5A5A
FILL = 0x7B3C5D1E  # what the buffers hold before a launch: no hit index, visibility word or pixel colour looks like it


@pytest.fixture(scope="module")
def ctx():
    c = gpu_context()
    yield c
    c.close()


class Arena:
    """One device allocation cut into [guard | buffer | guard | buffer | ... | guard]; buffers start at `align` bytes."""

    def __init__(self, sizes_words, align=512, skew_words=0):
        import torch

        self.offsets = []
        at = GUARD_WORDS
        step = align // 4
        for n in sizes_words:
            at = (at + step - 1) // step * step + skew_words
            self.offsets.append((at, n))
            at += n + GUARD_WORDS
        self.mem = torch.full((at,), PATTERN, dtype=torch.int32, device="cuda")
        assert self.mem.data_ptr() % 512 == 0

    def ptr(self, i):
        return self.mem.data_ptr() + 4 * self.offsets[i][0]

    def view(self, i):
        at, n = self.offsets[i]
        return self.mem[at:at + n]

    def fill(self, value):
        for i in range(len(self.offsets)):
            self.view(i).fill_(value)

    def guards_intact(self):
        import torch

        mask = torch.ones_like(self.mem, dtype=torch.bool)
        for at, n in self.offsets:
            mask[at:at + n] = False
        return bool((self.mem[mask] == PATTERN).all())


def _terrain_scene(ctx):
    m = rtb200.Mesh().terrain(96, 100.0).icosphere(3, 22.0, (15.0, 28.0, -20.0)).finish(diffuse=(0.6, 0.7, 0.8))
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    ctx.upload_scene(A, b.nodes, b.tri_indices)
    return A


def _restore(ctx):
    ctx.set_option("store_group", -1)
    ctx.set_option("tile_hints", 1)
    for k, v in (("hint_heavy_pct", 12), ("hint_split_pct", 80), ("hint_keep_pct", 30)):
        ctx.set_option(k, v)


@pytest.mark.parametrize("wh", [(97, 53), (328, 204), (642, 362)])
@pytest.mark.parametrize("parts", [(0, 1, 4), (2, 3, 4), (1, 3, 8)])
@pytest.mark.parametrize("skew_words", [0, 4, 1])  # 512-byte, 16-byte and (frames only) 4-byte aligned buffers
def test_no_store_lands_outside_the_callers_buffers(ctx, wh, parts, skew_words):
    """every device-buffer pass, with hints forced on (split rows, heavy/light lists) and a camera that leaves part of the
    frame to the `fill` items, run into buffers with 16 KiB of pattern either side: the pattern survives, and pixel rows
    the part does not own keep their previous contents"""
    import torch

    w, h = wh
    part, n_parts, band_rows = parts
    n = w * h
    A = _terrain_scene(ctx)
    # zoomed out + orbiting: the scene's rectangle covers part of the frame only, so culling and fill items engage
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=GRAZING, d_radius=160.0, d_alpha=0.3)
    ctx.set_params(params)
    ctx.set_option("tile_hints", 1)
    ctx.set_option("hint_heavy_pct", 30)
    ctx.set_option("hint_split_pct", 8)
    ctx.set_option("hint_keep_pct", 2)
    kw = dict(part=part, n_parts=n_parts, band_rows=band_rows)
    records_ok = skew_words % 4 == 0
    #        0 hits   1 rays   2 shadow hits  3 idx frame  4 vis frame  5 shaded frame  6 hits of the fused pass
    sizes = [4 * n, 8 * n, 4 * n, n, n, n, 4 * n]
    ar = Arena(sizes, skew_words=skew_words)
    rows = np.arange(h)
    owned = ((rows // band_rows) % n_parts) == part
    try:
        for launch in range(4):
            ctx.set_option("store_group", (0, 2, 4, -1)[launch])
            ar.fill(FILL)
            if records_ok:
                ctx.primary_device(w, h, ar.ptr(0), ar.ptr(1), **kw)
                ctx.primary_shadow_device(w, h, ar.ptr(6), ar.ptr(2), ar.ptr(4), **kw)
                ctx.primary_gather_device(w, h, ar.ptr(0), ar.ptr(3), **kw)
            else:
                ctx.primary_shadow_device(w, h, None, None, ar.ptr(4), **kw)
                ctx.primary_gather_device(w, h, None, ar.ptr(3), **kw)
            ctx.render_frame_device(w, h, ar.ptr(5), **kw)
            ctx.synchronize()
            assert ar.guards_intact(), f"launch {launch}: a store landed outside a buffer"
            for i in (3, 4, 5):
                fr = ar.view(i).cpu().numpy().reshape(h, w)
                assert (fr[~owned] == FILL).all(), f"launch {launch}: buffer {i}: a row of another part was written"
                assert (fr[owned] != FILL).all(), f"launch {launch}: buffer {i}: an owned pixel was not written"
            if records_ok:
                for i, words in ((0, 4), (1, 8), (2, 4), (6, 4)):
                    rec = ar.view(i).cpu().numpy().reshape(h, w * words)
                    assert (rec[~owned] == FILL).all(), f"launch {launch}: buffer {i}: a record of another part was written"
                idx = ar.view(0).cpu().numpy().reshape(h, w, 4)[..., 0]
                assert (idx[owned] != FILL).all()
    finally:
        _restore(ctx)


def test_ray_buffer_passes_stay_inside_their_buffers(ctx):
    """rt_trace_device / rt_trace_sorted_device / rt_shadow_device / rt_diffuse_rays_device with ray counts that are not
    multiples of the 32-ray work item, into guarded buffers"""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    ctx.set_params(g["params"])
    rays_h = np.ascontiguousarray(g["random_rays"], dtype=np.float32)
    for n in (1, 31, 33, 1000, rays_h.shape[0]):
        for sched in (0, 1):
            ctx.set_option("scheduler", sched)
            spp = 3
            #        0 rays  1 hits  2 any hits  3 sorted hits  4 shadow hits  5 shadow rays  6 diffuse rays   7 count
            ar = Arena([8 * n, 4 * n, 4 * n, 4 * n, 4 * n, 8 * n, 8 * n * spp, 2], skew_words=4)
            ar.fill(FILL)
            ar.view(0).copy_(torch.from_numpy(rays_h[:n].view(np.int32).reshape(-1)).cuda())
            ctx.trace_device(rtb200.CLOSEST, n, ar.ptr(0), ar.ptr(1))
            ctx.trace_device(rtb200.ANY, n, ar.ptr(0), ar.ptr(2))
            ctx.trace_sorted_device(rtb200.CLOSEST, n, ar.ptr(0), ar.ptr(3))
            ctx.shadow_device(n, ar.ptr(0), ar.ptr(1), ar.ptr(4), ar.ptr(5))
            ctx.diffuse_rays_device(n, ar.ptr(0), ar.ptr(1), spp, 7, ar.ptr(6), ar.ptr(7))
            ctx.synchronize()
            assert ar.guards_intact(), f"n={n} scheduler={sched}: a store landed outside a buffer"
            got = ar.view(1).cpu().numpy().view(HIT).reshape(-1)
            assert_hits_identical(got, g["random_hits_closest"].view(HIT).reshape(-1)[:n], f"n={n} closest")
            assert np.array_equal(ar.view(3).cpu().numpy(), ar.view(1).cpu().numpy()), "sorted trace differs"
    ctx.set_option("scheduler", -1)


def test_misaligned_device_buffers_are_refused_before_any_launch(ctx):
    """a record buffer at a 4-byte offset would make the kernels' 16-byte loads / stores fault; every entry point that takes
    one returns RT_E_INVALID instead, and the context keeps working"""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    ctx.set_params(g["params"])
    w, h = 64, 32
    n = w * h
    buf = torch.zeros(16 * n + 64, dtype=torch.int32, device="cuda")
    good, odd, odd2 = buf.data_ptr(), buf.data_ptr() + 4, buf.data_ptr() + 2
    launches_before = ctx.counters()["kernel_launches"]
    calls = [
        lambda: ctx.trace_device(rtb200.CLOSEST, n, odd, good),
        lambda: ctx.trace_device(rtb200.ANY, n, good, odd),
        lambda: ctx.trace_sorted_device(rtb200.CLOSEST, n, odd, good),
        lambda: ctx.trace_sorted_device(rtb200.CLOSEST, n, good, odd),
        lambda: ctx.primary_device(w, h, odd),
        lambda: ctx.primary_device(w, h, good, odd),
        lambda: ctx.primary_gather_device(w, h, None, odd2),
        lambda: ctx.primary_shadow_device(w, h, odd, None, None),
        lambda: ctx.primary_shadow_device(w, h, None, odd, None),
        lambda: ctx.primary_shadow_device(w, h, None, None, odd2),
        lambda: ctx.shadow_device(n, odd, good, good),
        lambda: ctx.shadow_device(n, good, odd, good),
        lambda: ctx.shadow_device(n, good, good, odd),
        lambda: ctx.shadow_device(n, good, good, good, odd),
        lambda: ctx.diffuse_rays_device(n, odd, good, 1, 0, good, None),
        lambda: ctx.diffuse_rays_device(n, good, good, 1, 0, odd, None),
        lambda: ctx.diffuse_rays_device(n, good, good, 1, 0, good, odd),
        lambda: ctx.render_frame_device(w, h, odd2),
    ]
    for i, call in enumerate(calls):
        with pytest.raises(rtb200.RtError, match="aligned"):
            call()
    assert ctx.counters()["kernel_launches"] == launches_before, "a refused call launched something"
    ctx.synchronize()  # no sticky error
    assert_hits_identical(ctx.trace(rtb200.CLOSEST, g["random_rays"]), g["random_hits_closest"].view(HIT).reshape(-1), "after refusals")


@pytest.mark.parametrize("zero_copy", [1, 0])
def test_host_buffers_of_any_alignment(ctx, zero_copy):
    """page-locked and pageable host buffers at odd offsets: rt_trace / rt_primary / rt_primary_shadow / rt_render_frame
    and the frames-in-flight pair fall back to the copy and give the same bytes as aligned buffers"""
    import torch

    g = load_scene("mix")
    ctx.upload_scene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"])
    ctx.set_params(g["params"])
    ctx.set_option("zero_copy", zero_copy)
    w, h = (int(v) for v in g["wh"])
    n = w * h
    rays = np.ascontiguousarray(g["random_rays"], dtype=np.float32)
    nr = rays.shape[0]
    try:
        want_hits = ctx.trace(rtb200.CLOSEST, rays).copy()
        want_prim = ctx.primary(w, h).copy()
        want_vis = ctx.primary_shadow(w, h).copy()
        want_img = ctx.render_frame(w, h).copy()
        for pinned in (True, False):
            for off in (4, 2, 1):
                room = max(16 * max(n, nr), 32 * nr) + 64
                store = torch.zeros(room, dtype=torch.uint8, pin_memory=pinned)
                raw = store.numpy()
                base = store.data_ptr()
                assert base % 16 == 0

                def carve(dtype, count):
                    raw[:] = 0xEE
                    return np.frombuffer(raw.data, dtype=dtype, count=count, offset=off)

                src = carve(np.float32, 8 * nr)
                src[:] = rays.reshape(-1)
                out_store = torch.zeros(room, dtype=torch.uint8, pin_memory=pinned)
                out_raw = out_store.numpy()
                out = np.frombuffer(out_raw.data, dtype=HIT, count=nr, offset=off)
                lib_out = ctx.trace(rtb200.CLOSEST, src.reshape(nr, 8), out)
                assert np.array_equal(lib_out.view(np.uint8), want_hits.view(np.uint8)), (pinned, off, "rt_trace")
                out = np.frombuffer(out_raw.data, dtype=HIT, count=n, offset=off)
                assert np.array_equal(ctx.primary(w, h, out).view(np.uint8), want_prim.view(np.uint8)), (pinned, off, "rt_primary")
                fr = np.frombuffer(out_raw.data, dtype=np.int32, count=n, offset=off).reshape(h, w)
                assert np.array_equal(ctx.primary_shadow(w, h, fr), want_vis), (pinned, off, "rt_primary_shadow")
                fr = np.frombuffer(out_raw.data, dtype=np.uint32, count=n, offset=off).reshape(h, w)
                assert np.array_equal(ctx.render_frame(w, h, fr), want_img), (pinned, off, "rt_render_frame")
                fr[:] = 0
                ctx.render_frame_begin(w, h, fr, 1)
                ctx.render_frame_end(1)
                assert np.array_equal(fr, want_img), (pinned, off, "rt_render_frame_begin/end")
                assert (out_raw[:off] == 0).all(), "bytes before the carved buffer were written"
    finally:
        ctx.set_option("zero_copy", 1)
