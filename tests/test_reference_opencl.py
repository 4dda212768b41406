"""The reference's OWN OpenCL kernel (unmodified volumeRender.cl, embedded in oracle/_ref/libref_cl.so at build time),
run through NVIDIA's OpenCL ICD on the GPU box, against the CPU oracle and the CUDA path. This is what pins the oracle's
traversal + shading to reference-produced output. Tolerance: the OpenCL compiler contracts to FMA and uses its own
normalize/pow/division, so single pixels may differ; coverage (which pixels hit anything) must be identical."""
import numpy as np
import pytest
from conftest import SCENES, channel_diff, load_scene, mesh_dict

import rtb200
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refcl_ok():
    try:
        O.refcl()
    except O.RefCLUnavailable as e:
        pytest.skip(f"reference OpenCL program cannot run here: {e}")
    return True


@pytest.mark.parametrize("name", SCENES)
def test_reference_kernel_frame_vs_oracle_and_cuda(refcl_ok, name):
    g = load_scene(name)
    md = mesh_dict(g)
    w, h = (int(v) for v in g["wh"])
    ref = O.RefCLScene(md, g["ref_nodes"], g["ref_tri_indices"])
    img, _ms = ref.render_frame(g["params"], w, h)
    # (1) the oracle's committed frame
    d = channel_diff(img, g["frame"])
    assert np.array_equal(img != 0, g["frame"] != 0), "coverage differs from the reference kernel"
    assert (d.max(axis=-1) > 1).mean() <= 0.002, f"{(d.max(axis=-1) > 1).sum()} pixels differ by more than 1 LSB"
    assert (d.max(axis=-1) > 0).mean() <= 0.01
    # (2) the CUDA path through the C ABI
    ctx = rtb200.Context(0)
    ctx.upload_scene(md, g["ref_nodes"], g["ref_tri_indices"])
    ctx.set_params(g["params"])
    d2 = channel_diff(img, ctx.render_frame(w, h))
    ctx.close()
    assert (d2.max(axis=-1) > 1).mean() <= 0.002 and (d2.max(axis=-1) > 0).mean() <= 0.01


def test_reference_kernel_on_a_larger_scene(refcl_ok):
    """~56 K triangles, 320x240, grazing light (order-dependent shadows): the reference kernel and the oracle agree on
    coverage everywhere and on colour in >= 99.5 % of the pixels"""
    m = rtb200.Mesh().terrain(160, 100.0).icosphere(4, 25.0, (20.0, 30.0, -10.0)).finish(diffuse=(0.6, 0.7, 0.8))
    A = m.arrays()
    b = rtb200.FlatBVH.build(m)
    w, h = 320, 240
    params, _ = rtb200.camera_params(w, h, A["aabb_min"], A["aabb_max"], light_pos=(-150.0, 25.0, 3.0))
    img, _ = O.RefCLScene(A, b.nodes, b.tri_indices).render_frame(params, w, h)
    want, _ = O.OracleScene(A, b.nodes, b.tri_indices).render_frame(params, w, h)
    d = channel_diff(img, want)
    assert np.array_equal(img != 0, want != 0)
    assert (d.max(axis=-1) > 1).mean() <= 0.005
