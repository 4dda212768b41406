"""ctypes view of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_HOST = os.path.join(_HERE, "_ref", "ref_host")
T_INIT = np.float32(4294967296.0)
HIT_DTYPE = np.dtype([("idx", np.int32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])
_lib = None


class Scene(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("indices", C.c_void_p), ("nodes", C.c_void_p), ("tri_indices", C.c_void_p),
                ("num_verts", C.c_int), ("num_tris", C.c_int), ("num_bvh_tris", C.c_int), ("num_bvh_nodes", C.c_int),
                ("normals", C.c_void_p), ("normal_indices", C.c_void_p), ("materials", C.c_void_p),
                ("tri_to_material", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=False, stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB_PATH)
        L.orc_trace.argtypes = [C.POINTER(Scene), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_trace_bruteforce.argtypes = [C.POINTER(Scene), C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_shadow_rays.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render_frame.argtypes = [C.POINTER(Scene), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ray_triangle.argtypes = [C.c_void_p] * 7
        L.orc_ray_triangle.restype = C.c_float
        L.orc_scene_box_gate.argtypes = [C.c_void_p] * 6
        L.orc_vec_probe.argtypes = [C.c_void_p] * 3
        L.orc_ray_box.argtypes = [C.c_void_p] * 5
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


class OracleScene:
    """Holds contiguous copies of the reference-layout arrays and the orc_scene struct over them."""

    def __init__(self, mesh, bvh_nodes, tri_indices):
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        self.verts, self.indices = c(mesh["verts"], np.float32), c(mesh["indices"], np.int32)
        self.nodes, self.tri = c(bvh_nodes, np.float32), c(tri_indices, np.int32)
        s = Scene()
        s.verts, s.indices, s.nodes, s.tri_indices = (self.verts.ctypes.data, self.indices.ctypes.data,
                                                      self.nodes.ctypes.data, self.tri.ctypes.data)
        s.num_verts, s.num_tris = self.verts.shape[0], self.indices.size // 3
        s.num_bvh_tris, s.num_bvh_nodes = self.tri.size, self.nodes.shape[0]
        if len(mesh.get("normals", ())) and len(mesh.get("materials", ())):
            self.normals, self.nidx = c(mesh["normals"], np.float32), c(mesh["normal_indices"], np.int32)
            self.mats, self.t2m = c(mesh["materials"], np.float32), c(mesh["tri_to_material"], np.int32)
            s.normals, s.normal_indices = self.normals.ctypes.data, self.nidx.ctypes.data
            s.materials, s.tri_to_material = self.mats.ctypes.data, self.t2m.ctypes.data
        self.c = s

    def trace(self, mode, rays, brute=False):
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        hits = np.empty(n, dtype=HIT_DTYPE)
        counters = np.zeros(4, dtype=np.uint64)
        if brute:
            lib().orc_trace_bruteforce(C.byref(self.c), mode, n, rays.ctypes.data, hits.ctypes.data)
        else:
            lib().orc_trace(C.byref(self.c), mode, n, rays.ctypes.data, hits.ctypes.data, counters.ctypes.data)
        return hits, {"inner": int(counters[0]), "leaf": int(counters[1]), "tris": int(counters[2]), "max_stack": int(counters[3])}

    def render_frame(self, params, w, h):
        p = np.ascontiguousarray(params, dtype=np.float32)
        out = np.empty((h, w), dtype=np.uint32)
        counters = np.zeros(4, dtype=np.uint64)
        lib().orc_render_frame(C.byref(self.c), p.ctypes.data, w, h, out.ctypes.data, counters.ctypes.data)
        return out, {"inner": int(counters[0]), "leaf": int(counters[1]), "tris": int(counters[2])}


def primary_rays(params, w, h):
    p = np.ascontiguousarray(params, dtype=np.float32)
    rays = np.empty((w * h, 8), dtype=np.float32)
    gate = np.empty(w * h, dtype=np.uint8)
    lib().orc_primary_rays(p.ctypes.data, w, h, rays.ctypes.data, gate.ctypes.data)
    return rays, gate


def shadow_rays(params, rays, hits):
    p = np.ascontiguousarray(params, dtype=np.float32)
    n = rays.shape[0]
    out = np.empty((n, 8), dtype=np.float32)
    valid = np.empty(n, dtype=np.uint8)
    lib().orc_shadow_rays(p.ctypes.data, n, np.ascontiguousarray(rays).ctypes.data, np.ascontiguousarray(hits).ctypes.data,
                          out.ctypes.data, valid.ctypes.data)
    return out, valid


def ref_host_available():
    return os.path.exists(REF_HOST)


def ref_host(mode, payload, tmpdir):
    """Run the reference's own compiled code (oracle/_ref/ref_host): payload bytes in, bytes out."""
    fin, fout = os.path.join(tmpdir, f"{mode}.in"), os.path.join(tmpdir, f"{mode}.out")
    with open(fin, "wb") as f:
        f.write(payload)
    subprocess.run([REF_HOST, mode, fin, fout], check=True, stdout=subprocess.DEVNULL)
    with open(fout, "rb") as f:
        return f.read()


# ---- the reference's own OpenCL kernel (oracle/_ref/libref_cl.so; needs an OpenCL ICD, i.e. a GPU box) ----------
REF_CL = os.path.join(_HERE, "_ref", "libref_cl.so")
_refcl = None


class RefCLUnavailable(RuntimeError):
    pass


def refcl(build_options=""):
    """Load + initialise the reference OpenCL program once per process; raises RefCLUnavailable if it cannot run here."""
    global _refcl
    if _refcl is None:
        if not os.path.exists(REF_CL):
            raise RefCLUnavailable("oracle/_ref/libref_cl.so not built (needs /root/reference at build time)")
        L = C.CDLL(REF_CL)
        L.refcl_last_error.restype = C.c_char_p
        L.refcl_device_name.restype = C.c_char_p
        L.refcl_build_note.restype = C.c_char_p
        L.refcl_init.argtypes = [C.c_char_p]
        L.refcl_upload_scene.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.refcl_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        if L.refcl_init(build_options.encode()):
            raise RefCLUnavailable(L.refcl_last_error().decode(errors="replace"))
        _refcl = L
    return _refcl


class RefCLScene:
    """The reference kernel `raytracer_bvh` on the reference-layout arrays, launched as RayTracer.cpp launches it."""

    def __init__(self, mesh, bvh_nodes, tri_indices, build_options=""):
        self.L = refcl(build_options)
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        v, i = c(mesh["verts"], np.float32), c(mesh["indices"], np.int32)
        n, t = c(bvh_nodes, np.float32), c(tri_indices, np.int32)
        nr, ni = c(mesh["normals"], np.float32), c(mesh["normal_indices"], np.int32)
        m, t2m = c(mesh["materials"], np.float32), c(mesh["tri_to_material"], np.int32)
        if self.L.refcl_upload_scene(v.ctypes.data, v.shape[0], i.ctypes.data, i.size // 3, n.ctypes.data, n.shape[0], t.ctypes.data,
                                     t.size, nr.ctypes.data, nr.shape[0], ni.ctypes.data, m.ctypes.data, m.shape[0], t2m.ctypes.data):
            raise RuntimeError(self.L.refcl_last_error().decode(errors="replace"))

    def device_name(self):
        return self.L.refcl_device_name().decode()

    def build_note(self):
        return self.L.refcl_build_note().decode()

    def render_frame(self, params, w, h):
        p = np.ascontiguousarray(params, dtype=np.float32)
        out = np.empty((h, w), dtype=np.uint32)
        ms = C.c_double(0)
        if self.L.refcl_render(p.ctypes.data, w, h, out.ctypes.data, C.byref(ms)):
            raise RuntimeError(self.L.refcl_last_error().decode(errors="replace"))
        return out, ms.value
