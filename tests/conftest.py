import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = ["test0", "test1", "ico2", "terrain12", "sticks150", "mix", "cubes2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the shared libraries exist (the driver normally ran __graft_entry__.build())."""
    import rtb200

    missing = [p for p in (rtb200.device.LIB_PATH, rtb200.hostlib._LIB_PATH) if not os.path.exists(p)]
    if missing or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__

        __graft_entry__.build()


def load_scene(name):
    z = np.load(os.path.join(GOLDEN, f"scene_{name}.npz"))
    return {k: z[k] for k in z.files}


def mesh_dict(g):
    return {k: g[k] for k in ("verts", "indices", "normals", "normal_indices", "materials", "tri_to_material", "aabb_min", "aabb_max")}


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_hits_identical(got, want, what=""):
    for k in ("idx", "t", "u", "v"):
        if not same_bits(got[k], want[k]):
            bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
            raise AssertionError(f"{what}: field {k} differs on {bad.size}/{got.size} rays, first at {bad[:5]}: "
                                 f"got {got[k][bad[:5]]} want {want[k][bad[:5]]}")


def channel_diff(img_a, img_b):
    sh = np.array([0, 8, 16])
    return np.abs(((img_a[..., None] >> sh) & 255).astype(np.int32) - ((img_b[..., None] >> sh) & 255).astype(np.int32))


def gpu_context(device=0):
    """A context whose kernels run on torch's current stream. The tests hand torch tensors to the device entry points; a
    context on its own (non-blocking) stream would race torch's asynchronous fills of those tensors (torch.zeros is a
    kernel on torch's stream): a 66 MB zero-fill once landed AFTER the primary pass had written the rays. A caller that
    mixes frameworks orders the streams itself -- that is what rt_set_stream is for (include/rtb200.h)."""
    import torch

    import rtb200

    c = rtb200.Context(device)
    c.set_stream(torch.cuda.current_stream(device).cuda_stream)
    return c
