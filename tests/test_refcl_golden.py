"""CPU pin of the oracle's whole-frame path (traversal + shading + shadows + reflections) to frames the reference's OWN
OpenCL kernel produced: tests/golden/refcl_frames.npz was rendered on a B200 by the unmodified volumeRender.cl through
NVIDIA's OpenCL ICD (tests/golden/make_refcl_golden.py). Tolerance as in test_reference_opencl.py: the OpenCL compiler
contracts to FMA and has its own normalize/pow/division, so single pixels may differ by an LSB; coverage is exact."""
import os

import numpy as np
import pytest
from conftest import GOLDEN, SCENES, channel_diff, load_scene, mesh_dict

from oracle import oracle_py as O

GRAZING_LIGHT = np.float32([-150, 25, 3, 1])


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLDEN, "refcl_frames.npz"))


def test_fixture_provenance(frames):
    assert "B200" in str(frames["device"])
    assert len([k for k in frames.files if k not in ("device", "build_note")]) == 2 * len(SCENES)


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("grazing", [False, True])
def test_oracle_frame_matches_reference_kernel_frame(frames, name, grazing):
    g = load_scene(name)
    w, h = (int(v) for v in g["wh"])
    params = g["params"].copy()
    if grazing:  # order-dependent any-hit shadows at a grazing light (make_refcl_golden.py)
        params[16:20] = GRAZING_LIGHT
    got, _ = O.OracleScene(mesh_dict(g), g["ref_nodes"], g["ref_tri_indices"]).render_frame(params, w, h)
    ref = frames[name + ("_grazing" if grazing else "")]
    assert np.array_equal(got != 0, ref != 0), "coverage differs from the reference kernel"
    d = channel_diff(got, ref).max(axis=-1)
    assert (d > 1).mean() <= 0.002, f"{(d > 1).sum()} pixels differ by more than 1 LSB"
    assert (d > 0).mean() <= 0.01
