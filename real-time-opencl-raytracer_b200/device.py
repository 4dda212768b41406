"""ctypes binding of the C ABI in include/rtb200.h (librtb200.so, hand-written sm_100a CUDA).

There is no CPU fallback: if the library is missing, or no CUDA device is present, the calls raise.
Host-side mirror of the reference's launch seam (reference RayTracer.cpp: setupCL :2370,
initRayTrace :858, updateCamera :609, raytrace_gpgpu :330, cleanup :1263)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(_HERE, "csrc", "librtb200.so")  # env override: kernel-variant experiments
_lib = None

CLOSEST, ANY = 0, 1
T_INIT = np.float32(4294967296.0)
RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmax", np.float32), ("d", np.float32, 3), ("reserved", np.float32)])
HIT_DTYPE = np.dtype([("idx", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])

ABI_SYMBOLS = [
    "rt_create", "rt_destroy", "rt_last_error", "rt_version", "rt_set_stream", "rt_synchronize", "rt_upload_scene",
    "rt_scene_blob", "rt_copy_scene_blob", "rt_adopt_scene_blob", "rt_set_params", "rt_render_frame", "rt_render_frame_begin", "rt_render_frame_end", "rt_trace", "rt_primary", "rt_trace_device", "rt_trace_sorted_device",
    "rt_primary_device", "rt_primary_gather_device", "rt_primary_shadow_device", "rt_primary_shadow", "rt_ipc_alloc", "rt_ipc_open", "rt_ipc_close", "rt_ipc_free",
    "rt_memcpy_to_host", "rt_host_register", "rt_host_unregister", "rt_signal", "rt_shadow_device", "rt_diffuse_rays_device", "rt_render_frame_device", "rt_get_counters",
    "rt_create_group", "rt_destroy_group", "rt_group_last_error", "rt_group_size", "rt_group_context", "rt_group_upload_scene",
    "rt_group_set_params", "rt_group_set_option", "rt_render_frame_tiled", "rt_primary_tiled", "rt_group_stats",
    "rt_reset_counters", "rt_set_option", "rt_scene_info", "rt_selftest", "rt_selftest_range", "rt_selftest_exhaustive", "rt_tile_hint_stats", "rt_cull_rect_host", "rt_pack_scene_host", "rt_free_host",
]


class RtError(RuntimeError):
    pass


def cull_rect(params, w, h):
    """host-only: inclusive pixel rectangle (x0, x1, y0, y1) outside of which no pixel can pass the scene gate"""
    p = np.ascontiguousarray(params, dtype=np.float32)
    out = (C.c_int64 * 4)()
    if lib().rt_cull_rect_host(p.ctypes.data, w, h, out):
        raise RtError("rt_cull_rect_host: bad arguments")
    return tuple(int(v) for v in out)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtError(f"{LIB_PATH} is missing (build it with __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.rt_create.argtypes = [i32, C.POINTER(vp)]
        L.rt_destroy.argtypes = [vp]
        L.rt_last_error.argtypes = [vp]
        L.rt_last_error.restype = C.c_char_p
        L.rt_version.restype = C.c_char_p
        L.rt_set_stream.argtypes = [vp, vp]
        L.rt_synchronize.argtypes = [vp]
        L.rt_upload_scene.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, vp]
        L.rt_scene_blob.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.rt_adopt_scene_blob.argtypes = [vp, vp, C.c_size_t]
        L.rt_copy_scene_blob.argtypes = [vp, vp, C.c_size_t]
        L.rt_set_params.argtypes = [vp, vp]
        L.rt_render_frame.argtypes = [vp, i32, i32, vp]
        L.rt_render_frame_begin.argtypes = [vp, i32, i32, vp, i32]
        L.rt_render_frame_end.argtypes = [vp, i32]
        L.rt_trace.argtypes = [vp, i32, i64, vp, vp]
        L.rt_trace_device.argtypes = [vp, i32, i64, vp, vp]
        L.rt_primary.argtypes = [vp, i32, i32, vp]
        L.rt_trace_sorted_device.argtypes = [vp, i32, i64, vp, vp]
        L.rt_primary_device.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
        L.rt_shadow_device.argtypes = [vp, i64, vp, vp, vp, vp]
        L.rt_primary_gather_device.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
        L.rt_primary_shadow_device.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
        L.rt_primary_shadow.argtypes = [vp, i32, i32, vp]
        L.rt_ipc_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
        L.rt_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.rt_ipc_close.argtypes = [vp, vp]
        L.rt_ipc_free.argtypes = [vp, vp]
        L.rt_memcpy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
        L.rt_host_register.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp)]
        L.rt_host_unregister.argtypes = [vp, vp]
        L.rt_signal.argtypes = [vp, vp, C.c_uint32]
        L.rt_diffuse_rays_device.argtypes = [vp, i64, vp, vp, i32, C.c_uint32, vp, vp]
        L.rt_render_frame_device.argtypes = [vp, i32, i32, i32, i32, i32, vp]
        L.rt_get_counters.argtypes = [vp, vp]
        L.rt_reset_counters.argtypes = [vp]
        L.rt_set_option.argtypes = [vp, C.c_char_p, i32]
        L.rt_scene_info.argtypes = [vp, vp]
        L.rt_selftest.argtypes = [vp, i64, C.c_uint32, C.POINTER(C.c_uint64)]
        L.rt_pack_scene_host.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, i32, C.POINTER(vp),
                                         C.POINTER(C.c_size_t), C.c_char_p, i32]
        L.rt_free_host.argtypes = [vp]
        L.rt_free_host.restype = None
        L.rt_selftest_range.argtypes = [vp, i64, C.c_uint32, i32, C.POINTER(C.c_uint64)]
        L.rt_tile_hint_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.rt_cull_rect_host.argtypes = [vp, i32, i32, C.POINTER(C.c_int64)]
        L.rt_selftest_exhaustive.argtypes = [vp, C.c_uint32, C.c_uint32, i32, i32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.rt_create_group.argtypes = [i32, vp, C.POINTER(vp)]
        L.rt_destroy_group.argtypes = [vp]
        L.rt_group_last_error.argtypes = [vp]
        L.rt_group_last_error.restype = C.c_char_p
        L.rt_group_size.argtypes = [vp]
        L.rt_group_context.argtypes = [vp, i32]
        L.rt_group_context.restype = vp
        L.rt_group_upload_scene.argtypes = L.rt_upload_scene.argtypes
        L.rt_group_set_params.argtypes = [vp, vp]
        L.rt_group_set_option.argtypes = [vp, C.c_char_p, i32]
        L.rt_render_frame_tiled.argtypes = [vp, i32, i32, vp]
        L.rt_primary_tiled.argtypes = [vp, i32, i32, i32, vp]
        L.rt_group_stats.argtypes = [vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    """numpy array -> host pointer; torch tensor / int -> device pointer; None -> NULL"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    return a.data_ptr()


def pack_scene_host(mesh, bvh_nodes, tri_indices, top_pairs=2047, shading=True):
    """Host-only packing (no GPU needed): returns the blob as a uint8 numpy array, or raises RtError."""
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    verts, indices = c(mesh["verts"], np.float32), c(mesh["indices"], np.int32)
    nodes, tri = c(bvh_nodes, np.float32), c(tri_indices, np.int32)
    use_shading = shading and len(mesh.get("normals", ())) > 0 and len(mesh.get("materials", ())) > 0
    if use_shading:
        normals, nidx = c(mesh["normals"], np.float32), c(mesh["normal_indices"], np.int32)
        mats, t2m = c(mesh["materials"], np.float32), c(mesh["tri_to_material"], np.int32)
        Vn, M = normals.shape[0], mats.shape[0]
    else:
        normals = nidx = mats = t2m = None
        Vn = M = 0
    out, nbytes, err = C.c_void_p(), C.c_size_t(), C.create_string_buffer(256)
    rc = lib().rt_pack_scene_host(_ptr(verts), verts.shape[0], _ptr(indices), indices.size // 3, _ptr(nodes), nodes.shape[0],
                                  _ptr(tri), tri.size, _ptr(normals), Vn, _ptr(nidx), _ptr(mats), M, _ptr(t2m), top_pairs,
                                  C.byref(out), C.byref(nbytes), err, 256)
    if rc:
        raise RtError(f"[{rc}] {err.value.decode()}")
    try:
        return np.frombuffer((C.c_char * nbytes.value).from_address(out.value), dtype=np.uint8).copy()
    finally:
        lib().rt_free_host(out)


def _scene_args(mesh, bvh_nodes, tri_indices, shading=True):
    """the 14 scene arguments of rt_upload_scene / rt_group_upload_scene + the arrays that must stay alive during the call"""
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    verts, indices = c(mesh["verts"], np.float32), c(mesh["indices"], np.int32)
    nodes, tri = c(bvh_nodes, np.float32), c(tri_indices, np.int32)
    use_shading = shading and len(mesh.get("normals", ())) > 0 and len(mesh.get("materials", ())) > 0
    if use_shading:
        normals, nidx = c(mesh["normals"], np.float32), c(mesh["normal_indices"], np.int32)
        mats, t2m = c(mesh["materials"], np.float32), c(mesh["tri_to_material"], np.int32)
        Vn, M = normals.shape[0], mats.shape[0]
    else:
        normals = nidx = mats = t2m = None
        Vn = M = 0
    keep = (verts, indices, nodes, tri, normals, nidx, mats, t2m)
    return (_ptr(verts), verts.shape[0], _ptr(indices), indices.size // 3, _ptr(nodes), nodes.shape[0], _ptr(tri), tri.size,
            _ptr(normals), Vn, _ptr(nidx), _ptr(mats), M, _ptr(t2m)), keep


class Group:
    """`rt_group`: one process driving N GPUs of one box (scene broadcast with NCCL, frames split in interleaved row bands,
    every GPU storing straight into the caller's frame). Mirrors Context for the frame-level calls."""

    STAT_NAMES = {"broadcast_ms": 0, "comm_init_ms": 1, "blob_bytes": 2, "last_call_ms": 3, "zero_copy": 4}

    def __init__(self, n_gpus, devices=None):
        self._h = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * n_gpus)(*devices)
        rc = lib().rt_create_group(n_gpus, dev, C.byref(self._h))
        if rc:
            raise RtError(f"rt_create_group({n_gpus}) failed [{rc}]: {lib().rt_group_last_error(None).decode()}")
        self.n = n_gpus

    def close(self):
        if getattr(self, "_h", None):
            lib().rt_destroy_group(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RtError(f"[{rc}] {lib().rt_group_last_error(self._h).decode()}")

    def upload_scene(self, mesh, bvh_nodes, tri_indices, shading=True):
        args, _keep = _scene_args(mesh, bvh_nodes, tri_indices, shading)
        self._ck(lib().rt_group_upload_scene(self._h, *args))

    def set_params(self, params32):
        p = np.ascontiguousarray(params32, dtype=np.float32)
        assert p.size == 32
        self._ck(lib().rt_group_set_params(self._h, p.ctypes.data))

    def set_option(self, name, value):
        self._ck(lib().rt_group_set_option(self._h, name.encode(), int(value)))

    def render_frame(self, w, h, out=None):
        if out is None:
            out = np.empty((h, w), dtype=np.uint32)
        self._ck(lib().rt_render_frame_tiled(self._h, w, h, _ptr(out)))
        return out

    def primary(self, w, h, with_shadow=False, out=None):
        """(h, w) int32 frame: hit index (with_shadow=False) or visibility word (True)"""
        if out is None:
            out = np.empty((h, w), dtype=np.int32)
        self._ck(lib().rt_primary_tiled(self._h, w, h, 1 if with_shadow else 0, _ptr(out)))
        return out

    def stats(self):
        out = np.zeros(24, dtype=np.float64)
        self._ck(lib().rt_group_stats(self._h, out.ctypes.data))
        d = {k: float(out[i]) for k, i in self.STAT_NAMES.items()}
        d["rank_kernel_ms"] = [float(v) for v in out[8:8 + self.n]]
        return d

    def rays_traced(self):
        """sum of RT_CNT_RAYS_TRACED over the group's contexts"""
        total = 0
        for r in range(self.n):
            out = np.zeros(8, dtype=np.uint64)
            h = lib().rt_group_context(self._h, r)
            if lib().rt_get_counters(h, out.ctypes.data):
                raise RtError("rt_get_counters failed")
            total += int(out[1])
        return total


class Context:
    """One `rt_context` = one GPU. Mirrors the reference's global CL state + RayTraceData."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().rt_create(device, C.byref(self._h))
        if rc:
            raise RtError(f"rt_create({device}) failed [{rc}]: {lib().rt_last_error(None).decode()}")
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RtError(f"[{rc}] {lib().rt_last_error(self._h).decode()}")

    # --- plumbing ---
    def set_stream(self, cuda_stream_handle):
        """Run on an existing CUDA stream. Handle 0 (the legacy default stream, e.g. torch's default
        current stream) is passed as cudaStreamLegacy (0x1); None restores the context's own stream."""
        if cuda_stream_handle is None:
            handle = None
        else:
            handle = 1 if int(cuda_stream_handle) == 0 else int(cuda_stream_handle)
        self._ck(lib().rt_set_stream(self._h, handle))

    def synchronize(self):
        self._ck(lib().rt_synchronize(self._h))

    def set_option(self, name, value):
        self._ck(lib().rt_set_option(self._h, name.encode(), int(value)))

    def counters(self):
        out = np.zeros(8, dtype=np.uint64)
        self._ck(lib().rt_get_counters(self._h, out.ctypes.data))
        return {"kernel_launches": int(out[0]), "rays_traced": int(out[1]), "h2d_bytes": int(out[2]), "d2h_bytes": int(out[3])}

    def reset_counters(self):
        self._ck(lib().rt_reset_counters(self._h))

    def selftest(self, samples, seed=1):
        bad = C.c_uint64(0)
        self._ck(lib().rt_selftest(self._h, samples, seed, C.byref(bad)))
        return int(bad.value)

    def selftest_range(self, samples, x_exponent, seed=1):
        bad = C.c_uint64(0)
        self._ck(lib().rt_selftest_range(self._h, samples, seed, x_exponent, C.byref(bad)))
        return int(bad.value)

    def tile_hint_stats(self):
        out = (C.c_uint64 * 4)()
        self._ck(lib().rt_tile_hint_stats(self._h, out))
        return {"split_rows": int(out[0]), "heavy_tiles": int(out[1]), "tiles_timed": int(out[2]), "span_cycles": int(out[3])}

    def selftest_exhaustive(self, md_begin, md_count, ex_x=0, ex_d=0, signs=0):
        """all 2^23 numerator mantissas x md_count divisor mantissas: mismatches of the hoisted division vs `/`"""
        bad = C.c_uint64(0)
        self._ck(lib().rt_selftest_exhaustive(self._h, md_begin, md_count, ex_x, ex_d, signs, C.byref(bad)))
        return int(bad.value)

    def scene_info(self):
        out = np.zeros(6, dtype=np.int64)
        self._ck(lib().rt_scene_info(self._h, out.ctypes.data))
        return {"node_pairs": int(out[0]), "packed_tris": int(out[1]), "blob_bytes": int(out[2]), "max_depth": int(out[3]),
                "hoisted_division": bool(out[4]), "bfs_pairs": int(out[5])}

    # --- scene (reference initRayTrace buffers) ---
    def upload_scene(self, mesh, bvh_nodes, tri_indices, shading=True):
        """mesh: dict from hostlib.Mesh.arrays(); bvh_nodes (N,12) f32 words; tri_indices (R,) i32."""
        args, _keep = _scene_args(mesh, bvh_nodes, tri_indices, shading)
        self._ck(lib().rt_upload_scene(self._h, *args))

    def scene_blob(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(lib().rt_scene_blob(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def copy_scene_blob(self, dst, nbytes):
        self._ck(lib().rt_copy_scene_blob(self._h, _ptr(dst), nbytes))

    def adopt_scene_blob(self, device_ptr, nbytes):
        self._ck(lib().rt_adopt_scene_blob(self._h, device_ptr, nbytes))

    def set_params(self, params32):
        p = np.ascontiguousarray(params32, dtype=np.float32)
        assert p.size == 32
        self._ck(lib().rt_set_params(self._h, p.ctypes.data))

    # --- hot path, host buffers ---
    def render_frame(self, w, h, out=None):
        if out is None:
            out = np.empty((h, w), dtype=np.uint32)
        self._ck(lib().rt_render_frame(self._h, w, h, _ptr(out)))
        return out

    def render_frame_begin(self, w, h, out, slot=0):
        """enqueue the frame of the last set_params() into `out` (numpy uint32 (h, w) or pinned torch tensor, kept alive
        by the caller until render_frame_end(slot)) and return without waiting"""
        self._ck(lib().rt_render_frame_begin(self._h, w, h, _ptr(out), slot))

    def render_frame_end(self, slot=0):
        self._ck(lib().rt_render_frame_end(self._h, slot))

    def trace(self, mode, rays, hits=None):
        n = rays.shape[0] if isinstance(rays, np.ndarray) else rays.numel() * rays.element_size() // 32
        if hits is None:
            hits = np.empty(n, dtype=HIT_DTYPE)
        self._ck(lib().rt_trace(self._h, mode, n, _ptr(rays), _ptr(hits)))
        return hits

    def primary(self, w, h, hits=None):
        """primary-ray pass, host buffer out (numpy array or pinned torch tensor of w*h 16-byte records)"""
        if hits is None:
            hits = np.empty(w * h, dtype=HIT_DTYPE)
        self._ck(lib().rt_primary(self._h, w, h, _ptr(hits)))
        return hits

    # --- hot path, device buffers (torch tensors or raw device pointers) ---
    def trace_device(self, mode, n, d_rays, d_hits):
        self._ck(lib().rt_trace_device(self._h, mode, n, _ptr(d_rays), _ptr(d_hits)))

    def trace_sorted_device(self, mode, n, d_rays, d_hits):
        """rt_trace_device with the coherence-sorting pre-pass (same results, different execution order)"""
        self._ck(lib().rt_trace_sorted_device(self._h, mode, n, _ptr(d_rays), _ptr(d_hits)))

    def primary_device(self, w, h, d_hits, d_rays_out=None, part=0, n_parts=1, band_rows=4):
        self._ck(lib().rt_primary_device(self._h, w, h, part, n_parts, band_rows, _ptr(d_hits), _ptr(d_rays_out)))

    def primary_gather_device(self, w, h, d_hits, d_idx_frame, part=0, n_parts=1, band_rows=4):
        """primary pass that also (or only) stores the 4-byte hit index per pixel into `d_idx_frame` (may be peer memory)"""
        self._ck(lib().rt_primary_gather_device(self._h, w, h, part, n_parts, band_rows, _ptr(d_hits), _ptr(d_idx_frame)))

    def primary_shadow_device(self, w, h, d_hits=None, d_shadow_hits=None, d_vis_frame=None, part=0, n_parts=1, band_rows=4):
        """primary + any-hit shadow pass in one launch; vis word = -1 | 3*triId + occluded (see include/rtb200.h)"""
        self._ck(lib().rt_primary_shadow_device(self._h, w, h, part, n_parts, band_rows, _ptr(d_hits), _ptr(d_shadow_hits), _ptr(d_vis_frame)))

    def primary_shadow(self, w, h, vis=None):
        """host-buffer form: (h, w) int32 visibility frame (numpy array or pinned torch tensor)"""
        if vis is None:
            vis = np.empty((h, w), dtype=np.int32)
        self._ck(lib().rt_primary_shadow(self._h, w, h, _ptr(vis)))
        return vis

    def ipc_alloc(self, nbytes):
        p, handle = C.c_void_p(), C.create_string_buffer(64)
        self._ck(lib().rt_ipc_alloc(self._h, nbytes, C.byref(p), handle))
        return p.value, handle.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._ck(lib().rt_ipc_open(self._h, handle, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._ck(lib().rt_ipc_close(self._h, ptr))

    def ipc_free(self, ptr):
        self._ck(lib().rt_ipc_free(self._h, ptr))

    def host_register(self, host_ptr, nbytes):
        """page-lock host memory; returns the device alias kernels may store into"""
        p = C.c_void_p()
        self._ck(lib().rt_host_register(self._h, host_ptr, nbytes, C.byref(p)))
        return p.value

    def host_unregister(self, host_ptr):
        self._ck(lib().rt_host_unregister(self._h, host_ptr))

    def signal(self, d_word, value):
        """stream-ordered: set the 32-bit word at device-visible address `d_word` once the work enqueued so far is done"""
        self._ck(lib().rt_signal(self._h, d_word, int(value) & 0xFFFFFFFF))

    def memcpy_to_host(self, dst, src_ptr, nbytes):
        self._ck(lib().rt_memcpy_to_host(self._h, _ptr(dst), src_ptr, nbytes))

    def shadow_device(self, n, d_rays, d_hits, d_shadow_hits, d_shadow_rays_out=None):
        self._ck(lib().rt_shadow_device(self._h, n, _ptr(d_rays), _ptr(d_hits), _ptr(d_shadow_hits), _ptr(d_shadow_rays_out)))

    def diffuse_rays_device(self, n, d_rays, d_hits, spp, seed, d_out_rays, d_count):
        self._ck(lib().rt_diffuse_rays_device(self._h, n, _ptr(d_rays), _ptr(d_hits), spp, seed, _ptr(d_out_rays), _ptr(d_count)))

    def render_frame_device(self, w, h, d_out, part=0, n_parts=1, band_rows=4):
        self._ck(lib().rt_render_frame_device(self._h, w, h, part, n_parts, band_rows, _ptr(d_out)))
