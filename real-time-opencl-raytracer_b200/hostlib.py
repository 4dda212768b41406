"""ctypes view of the host scene layer (librtb200_host.so): Mesh / SceneGen / ColladaLoader /
SplitBVHBuilder / BVH_Cuda / Camera. CPU only; the device boundary lives in device.py."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "host", "librtb200_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIB_PATH)
        L.rth_mesh_new.restype = C.c_void_p
        L.rth_bvh_build.restype = C.c_void_p
        L.rth_bvh_build.argtypes = [C.c_void_p, C.c_int]
        L.rth_bvh_load.restype = C.c_void_p
        L.rth_bvh_load.argtypes = [C.c_char_p]
        L.rth_bvh_save.argtypes = [C.c_void_p, C.c_char_p]
        L.rth_bvh_free.argtypes = [C.c_void_p]
        L.rth_bvh_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_double)]
        for name in ("rth_mesh_free", "rth_mesh_clear", "rth_mesh_finish"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.rth_mesh_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.rth_mesh_icosphere.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4
        L.rth_mesh_terrain.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.rth_mesh_sphere_field.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float]
        L.rth_mesh_sticks.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_float]
        L.rth_mesh_bvh_test.argtypes = [C.c_void_p, C.c_int]
        L.rth_mesh_set_diffuse.argtypes = [C.c_void_p] + [C.c_float] * 3
        L.rth_mesh_write_dae.argtypes = [C.c_void_p, C.c_char_p]
        L.rth_mesh_load_dae.argtypes = [C.c_void_p, C.c_char_p]
        L.rth_mesh_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_float)]
        L.rth_camera_params.argtypes = [C.c_float] * 3 + [C.c_int] * 2 + [C.c_void_p] * 6
        _lib = L
    return _lib


def set_num_threads(n):
    """threads for the SBVH builder (OpenMP tasks); torchrun workers start with OMP_NUM_THREADS=1"""
    lib().rth_set_num_threads(int(n))


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Mesh:
    """Host `Mesh` (reference Mesh.h): numpy views alias the C++ vectors."""

    def __init__(self):
        self._h = C.c_void_p(lib().rth_mesh_new())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rth_mesh_free(self._h)
            self._h = None

    # --- generators (SceneGen.h) ---
    def icosphere(self, subdiv, radius=50.0, center=(0.0, 0.0, 0.0)):
        lib().rth_mesh_icosphere(self._h, subdiv, radius, *center)
        return self

    def terrain(self, nquads=707, half=100.0):
        lib().rth_mesh_terrain(self._h, nquads, half)
        return self

    def sphere_field(self, grid=11, pitch=120.0, subdiv=6, radius=50.0):
        lib().rth_mesh_sphere_field(self._h, grid, pitch, subdiv, radius)
        return self

    def sticks(self, count, seed=1, extent=100.0):
        lib().rth_mesh_sticks(self._h, count, seed, extent)
        return self

    def bvh_test(self, which):
        lib().rth_mesh_bvh_test(self._h, which)
        return self

    def set(self, verts4, indices):
        v = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
        i = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        lib().rth_mesh_set(self._h, v.ctypes.data, v.shape[0], i.ctypes.data, i.shape[0])
        return self

    def finish(self, diffuse=None):
        if diffuse is not None:
            lib().rth_mesh_set_diffuse(self._h, *diffuse)
        lib().rth_mesh_finish(self._h)
        return self

    def write_dae(self, path):
        if lib().rth_mesh_write_dae(self._h, os.fsencode(path)):
            raise IOError(f"cannot write {path}")

    def load_dae(self, path):
        if lib().rth_mesh_load_dae(self._h, os.fsencode(path)):
            raise IOError(f"cannot load {path}")
        return self

    def arrays(self):
        counts = (C.c_int * 4)()
        ptrs = (C.c_void_p * 6)()
        aabb = (C.c_float * 6)()
        lib().rth_mesh_ptrs(self._h, counts, ptrs, aabb)
        V, T, Vn, M = list(counts)
        return {
            "verts": _view(ptrs[0], (V, 4), np.float32),
            "indices": _view(ptrs[1], (T * 3,), np.int32),
            "normals": _view(ptrs[2], (Vn, 4), np.float32),
            "normal_indices": _view(ptrs[3], (T * 3 if Vn else 0,), np.int32),
            "materials": _view(ptrs[4], (M, 44), np.float32),
            "tri_to_material": _view(ptrs[5], (T if M else 0,), np.int32),
            "aabb_min": np.array(aabb[0:3], dtype=np.float32),
            "aabb_max": np.array(aabb[3:6], dtype=np.float32),
            "_owner": self,  # the views alias this Mesh's C++ vectors: keep it alive with the dict
        }


class FlatBVH:
    """`BVH_Cuda` (reference BVH_Cuda.h): nodes as (N,12) float32 words + tri_indices (R,) int32."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("BVH build/load failed")
        self._h = C.c_void_p(handle)
        counts = (C.c_int * 3)()
        ptrs = (C.c_void_p * 2)()
        secs = C.c_double(0)
        lib().rth_bvh_ptrs(self._h, counts, ptrs, C.byref(secs))
        self.num_nodes, self.num_refs, self.duplicates = list(counts)
        self.build_seconds = secs.value
        self.nodes = _view(ptrs[0], (self.num_nodes, 12), np.float32)
        self.tri_indices = _view(ptrs[1], (self.num_refs,), np.int32)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rth_bvh_free(self._h)
            self._h = None

    @classmethod
    def build(cls, mesh, parallel_threshold=16384):
        return cls(lib().rth_bvh_build(mesh._h, parallel_threshold))

    @classmethod
    def load(cls, path):
        return cls(lib().rth_bvh_load(os.fsencode(path)))

    def save(self, path):
        if lib().rth_bvh_save(self._h, os.fsencode(path)):
            raise IOError(f"cannot write {path}")


def camera_params(w, h, aabb_min, aabb_max, light_pos=(-23.0, 200.0, 3.0), light_color=(1.0, 1.0, 1.0),
                  d_radius=0.0, d_alpha=0.0, d_beta=0.0):
    """Reference `Camera()` default pose (+ optional add_radius/add_rotate) -> 128-byte Params block."""
    f3 = lambda v: np.ascontiguousarray(v, dtype=np.float32)
    lp, lc, mn, mx = f3(light_pos), f3(light_color), f3(aabb_min), f3(aabb_max)
    out = np.zeros(32, dtype=np.float32)
    eye = np.zeros(3, dtype=np.float32)
    lib().rth_camera_params(d_radius, d_alpha, d_beta, w, h, lp.ctypes.data, lc.ctypes.data, mn.ctypes.data,
                            mx.ctypes.data, out.ctypes.data, eye.ctypes.data)
    return out, eye
