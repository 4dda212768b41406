"""The culled tile queue (csrc/rtb200.cu cull_setup) rests on one geometric claim: every pixel whose ray passes the scene-AABB
gate (vR.cl:1156-1196) lies inside the bounding rectangle of the box corners projected through the eye onto the image plane.
Checked here on the CPU tier against the ORACLE's gate, pixel by pixel, for random orbit poses, boxes and frame sizes --
including eyes inside / beside the box (no bound: whole frame) and boxes off screen."""
import numpy as np
import pytest

import rtb200
from oracle import oracle_py as O


def _gate(params, w, h):
    _rays, gate = O.primary_rays(params, w, h)
    return gate.reshape(h, w).astype(bool)


@pytest.mark.parametrize("seed", range(6))
def test_every_gate_passing_pixel_is_inside_the_rectangle(seed):
    rng = np.random.default_rng(100 + seed)
    tight = whole = empty = 0
    for _ in range(60):
        w, h = int(rng.integers(3, 60)) * 8 + int(rng.integers(0, 8)), int(rng.integers(3, 50)) * 4 + int(rng.integers(0, 4))
        lo = rng.uniform(-120, -5, 3).astype(np.float32)
        hi = (lo + rng.uniform(1, 230, 3)).astype(np.float32)
        pose = dict(d_radius=float(rng.uniform(-199.0, 900.0)), d_alpha=float(rng.uniform(-3.2, 3.2)), d_beta=float(rng.uniform(-0.78, 0.78)))
        params, _ = rtb200.camera_params(w, h, lo, hi, **pose)
        if rng.random() < 0.25:  # look past the box
            params = params.copy()
            params[8:11] += params[0:3] * float(rng.uniform(-2.5, 2.5)) + params[4:7] * float(rng.uniform(-2.5, 2.5))
        x0, x1, y0, y1 = rtb200.device.cull_rect(params, w, h)
        g = _gate(params, w, h)
        ys, xs = np.nonzero(g)
        if xs.size:
            assert x0 <= xs.min() and xs.max() <= x1 and y0 <= ys.min() and ys.max() <= y1, (pose, (w, h), (x0, x1, y0, y1), (xs.min(), xs.max(), ys.min(), ys.max()))
        if (x0, x1, y0, y1) == (0, w - 1, 0, h - 1):
            whole += 1
        elif x0 > x1 or y0 > y1:
            empty += 1
            assert not g.any()
        else:
            tight += 1  # (a thin tip of the projected box may reach a few pixels past the last pixel centre it covers)
    assert tight >= 10 and whole >= 1


def test_rectangle_of_the_bench_frame():
    """the bench frame (default camera over the 200 x 200 terrain box): the rectangle is a proper part of the frame"""
    w, h = 1920, 1080
    params, _ = rtb200.camera_params(w, h, (-100.0, -7.5, -100.0), (100.0, 7.5, 100.0))
    x0, x1, y0, y1 = rtb200.device.cull_rect(params, w, h)
    g = _gate(params, w, h)
    ys, xs = np.nonzero(g)
    assert x0 <= xs.min() <= x0 + 4 and x1 - 4 <= xs.max() <= x1 and y0 <= ys.min() <= y0 + 4 and y1 - 4 <= ys.max() <= y1
    assert (x1 - x0 + 1) * (y1 - y0 + 1) < 0.8 * w * h


def test_eye_inside_the_box_gives_the_whole_frame():
    w, h = 320, 200
    params, _ = rtb200.camera_params(w, h, (-300.0, -300.0, -300.0), (300.0, 300.0, 300.0))
    assert rtb200.device.cull_rect(params, w, h) == (0, w - 1, 0, h - 1)
