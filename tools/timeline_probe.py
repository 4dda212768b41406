"""Where does a launch's time go when a GPU holds only 1/N of a frame (BASELINE config 5, strong scaling)?
Needs the tools build of the library (per-batch timestamps):
    make -C real-time-opencl-raytracer_b200/csrc timeline
    RTB200_LIB=real-time-opencl-raytracer_b200/csrc/build/librtb200_timeline.so python tools/timeline_probe.py [out.json]
One GPU plays rank `part` of n_parts (the kernel a rank runs at N GPUs is exactly this launch): 3840x2160, C2 terrain, fused
primary + shadow pass and the shaded frame. Per launch: event time, first tile start, queue exhausted (last tile
fetched), last tile end, and the tile-duration distribution of tiles that trace (> 2 us) vs gated-out tiles."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import device as D

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_probe.json"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
RAW_DIR = sys.argv[4] if len(sys.argv) > 4 else None  # also dump the raw per-tile stamps there
PARTS = tuple(int(v) for v in os.environ.get("RTB_PROBE_PARTS", "1,2,4,8").split(","))
rtb200.hostlib.set_num_threads(os.cpu_count() or 1)
mesh = rtb200.Mesh().terrain(707, 100.0).finish()
A = mesh.arrays()
bvh = rtb200.FlatBVH.build(mesh)
params, _ = rtb200.camera_params(W, H, A["aabb_min"], A["aabb_max"])
ctx = rtb200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(A, bvh.nodes, bvh.tri_indices)
ctx.set_params(params)
for opt in ("tile_hints", "hint_heavy_pct", "hint_light_pct", "hint_split_pct", "hint_keep_pct"):  # A/B from the environment, e.g. RTB_TILE_HINTS=0
    if ("RTB_" + opt.upper()) in os.environ:
        ctx.set_option(opt, int(os.environ["RTB_" + opt.upper()]))
L = D.lib()
have_tl = hasattr(L, "rt_debug_timeline")
if have_tl:
    L.rt_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frame = torch.zeros((H, W), dtype=torch.int32, device="cuda")
tiles_x = W // 8


def launch(kind, part, n_parts, band_rows=16):
    if kind == "primary_shadow":
        ctx.primary_shadow_device(W, H, None, None, frame, part=part, n_parts=n_parts, band_rows=band_rows)
    elif kind == "primary":
        ctx.primary_gather_device(W, H, None, frame, part=part, n_parts=n_parts, band_rows=band_rows)
    else:
        ctx.render_frame_device(W, H, frame, part=part, n_parts=n_parts, band_rows=band_rows)


def timed(fn, iters=7, do_flush=True):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def n_batches(part, n_parts, band_rows):
    tile_rows = H // 4
    btr = band_rows // 4
    bands = (tile_rows + btr - 1) // btr
    mine = sum(min(btr, tile_rows - b * btr) for b in range(part, bands, n_parts))
    return mine * tiles_x


def timeline(kind, part, n_parts, band_rows=16, do_flush=True):
    nb = n_batches(part, n_parts, band_rows)
    buf = torch.zeros((nb, 2), dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(4):  # the tile hints settle after two launches of a geometry
            launch(kind, part, n_parts, band_rows)
    torch.cuda.synchronize()
    if do_flush:
        flush.fill_(1)
    torch.cuda.synchronize()
    L.rt_debug_timeline(ctx._h, C.c_void_p(buf.data_ptr()))
    with torch.cuda.stream(stream):
        launch(kind, part, n_parts, band_rows)
    torch.cuda.synchronize()
    L.rt_debug_timeline(ctx._h, None)
    t = buf.cpu().numpy().astype(np.uint64)
    ok = t[:, 0] > 0
    t0 = (t[:, 0] & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.int64)  # the end stamp carries the SM id in its top byte
    t1 = (t[:, 1] & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.int64)
    sm = (t[:, 1] >> np.uint64(56)).astype(np.int64)
    if not ok.any():  # kernel without timeline stamps (render_kernel)
        return None
    if RAW_DIR:
        np.savez_compressed(os.path.join(RAW_DIR, f"timeline_{kind}_n{n_parts}_part{part}.npz"), t0=t0, t1=t1, sm=sm, ok=ok,
                            w=W, h=H, part=part, n_parts=n_parts, band_rows=band_rows)
    t0, t1, sm = t0[ok], t1[ok], sm[ok]
    base = t0.min()
    dur = (t1 - t0) / 1e3
    busy = dur > 2.0
    end = (t1.max() - base) / 1e3
    # SM-occupancy curve: number of tiles in flight at 10 us steps
    grid = np.arange(0, end, 10.0)
    inflight = [int(((t0 - base) / 1e3 <= g).sum() - ((t1 - base) / 1e3 <= g).sum()) for g in grid]
    work_us = float(dur[busy].sum())
    return {
        "batches": int(nb), "tiles_tracing": int(busy.sum()), "kernel_span_us": float(end),
        "queue_exhausted_at_us": float((t0.max() - base) / 1e3),
        "tail_us": float(end - (t0.max() - base) / 1e3),
        "tracing_tile_us": {"median": float(np.median(dur[busy])), "p90": float(np.percentile(dur[busy], 90)), "max": float(dur[busy].max()),
                            "sum_ms": work_us / 1e3},
        "gated_tile_us": {"median": float(np.median(dur[~busy])) if (~busy).any() else None, "count": int((~busy).sum()),
                          "sum_ms": float(dur[~busy].sum() / 1e3)},
        "warp_slots": 148 * 32,
        "mean_tiles_in_flight": float(np.mean(inflight)) if inflight else 0.0,
        "tiles_in_flight_every_10us": inflight,
        "sms_used": int(np.unique(sm).size),
    }


res = {"frame": [W, H], "scene": "C2 terrain (999 698 triangles), default camera + light", "timeline_build": bool(have_tl)}
for kind in ("primary_shadow", "primary", "shaded_frame"):
    r = {}
    for n_parts in PARTS:
        per_part = [timed(lambda p=p: launch(kind, p, n_parts)) for p in range(n_parts)]
        r[f"n{n_parts}"] = {"max_ms": max(per_part), "mean_ms": float(np.mean(per_part)), "per_part_ms": per_part}
    if 8 in PARTS:
        r["n8_no_flush_ms"] = max(timed(lambda p=p: launch(kind, p, 8), do_flush=False) for p in range(8))
    for n_parts in PARTS:
        r[f"speedup_n{n_parts}"] = r["n1"]["max_ms"] / r[f"n{n_parts}"]["max_ms"]
    if have_tl:
        r["timeline_n8_part3"] = timeline(kind, 3, 8)
        r["timeline_n1"] = timeline(kind, 0, 1)
    launch(kind, 3, 8)
    r["hint_stats_n8_part3"] = ctx.tile_hint_stats()
    res[kind] = r
    print(kind, json.dumps({k: (v if not isinstance(v, dict) or "per_part_ms" not in v else {"max_ms": v["max_ms"], "mean_ms": v["mean_ms"]})
                            for k, v in r.items() if not k.startswith("timeline")}), flush=True)
    if have_tl:
        for k in ("timeline_n8_part3", "timeline_n1"):
            if r[k]:
                print(" ", k, json.dumps(r[k]), flush=True)

# fixed cost: the same launch with a camera that looks away from the scene (every pixel gated out)
p2 = params.copy()
p2[24:27] += 1e6  # aabb_min far away: nothing passes the gate
p2[28:31] += 1e6
ctx.set_params(p2)
res["all_gated"] = {f"n{n}": timed(lambda: launch("primary_shadow", 0, n)) for n in (1, 8)}
print("all gated", res["all_gated"])
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
