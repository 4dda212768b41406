#!/usr/bin/env python
"""bench.py -- headline benchmark: Mrays/s of primary rays (fused camera ray generation + closest-hit
SBVH traversal + ray/triangle intersection) on BASELINE.json configs[1]:
1920x1080 primary rays vs the ~1M-triangle displaced-grid terrain, SBVH, one B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

A step = one pass of the hot path over one frame of primary rays. Rays counted = traversals completed
(pixels that pass the scene-AABB gate call traverse_bvh once; the others do not, as in the reference).
N > 1: weak scaling -- the frame grows with N at fixed aspect (N = 4 is exactly the 3840x2160 frame
of configs[4]) so every rank keeps ~2.07 M pixels, partitioned in interleaved 16-row bands; the scene
is built on rank 0's host and broadcast once as one device buffer (NCCL); every step ends with the
gather of the 4-byte/pixel framebuffer (hit triangle index), the only data-path collective.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s primary (1080p, 1M-tri terrain, SBVH)"
UNIT = "Mrays/s"
TERRAIN_QUADS = 707  # 999 698 triangles
BAND_ROWS = 16


def frame_size(n_gpus):
    """~2.07 M pixels per GPU at 16:9; w % 8 == 0, h % 4 == 0 (the kernel's 8x4 warp tile)."""
    s = float(n_gpus) ** 0.5
    return int(round(1920 * s / 8)) * 8, int(round(1080 * s / 4)) * 4


def build_scene():
    import rtb200

    rtb200.hostlib.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1; only rank 0 builds
    t0 = time.time()
    mesh = rtb200.Mesh().terrain(TERRAIN_QUADS, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
    arrays = mesh.arrays()
    bvh = rtb200.FlatBVH.build(mesh)
    return mesh, arrays, bvh, time.time() - t0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """samples taken from now on are the ones reported (the load window)"""
        self.first = len(self.rows)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[getattr(self, "first", 0):]:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pass(arrays, bvh, params, w, h, repeats, threads=None):
    """CPU restatement of the reference kernel (oracle/) on the same frame: seconds per pass + visit counts."""
    from oracle import oracle_py as O

    if threads:
        O.lib().orc_set_num_threads(threads)
    sc = O.OracleScene(arrays, bvh.nodes, bvh.tri_indices)
    rays, gate = O.primary_rays(params, w, h)
    live = np.ascontiguousarray(rays[gate.astype(bool)])
    times, cnt = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _hits, cnt = sc.trace(0, live)
        times.append(time.perf_counter() - t0)
    return times, cnt, live.shape[0], O.lib().orc_num_threads()


def reference_opencl_frame(arrays, bvh, params, w, h, repeats=5):
    """The reference's OWN OpenCL kernel (unmodified, oracle/_ref/libref_cl.so) on this box's OpenCL device -- on the GPU
    boxes that is the same B200 through NVIDIA's ICD. It renders the whole frame (primary + shadow + reflections + shading):
    returns the median kernel time, or None where no OpenCL device / no prebuilt library exists."""
    try:
        from oracle import oracle_py as O

        ref = O.RefCLScene(arrays, bvh.nodes, bvh.tri_indices)
        ms = [ref.render_frame(params, w, h)[1] for _ in range(repeats + 1)][1:]
        return {"device": ref.device_name(), "kernel": "raytracer_bvh (volumeRender.cl, unmodified)", "frame": [w, h],
                "frame_kernel_ms": float(np.median(ms)), "build_note": ref.build_note()}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": str(exc)[:200]}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path. pocl/OpenCL do not exist in
    this image, so this is the oracle port (oracle/oracle.c, line-by-line restatement of volumeRender.cl)
    on all host threads, same frame, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rtb200

    _mesh, arrays, bvh, _ = build_scene()
    w, h = frame_size(1)
    params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"])
    times, cnt, nrays, threads = oracle_pass(arrays, bvh, params, w, h, args.warmup + args.steps)
    timed = times[args.warmup:]
    sec = float(np.sum(timed))
    value = nrays * len(timed) / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / len(timed) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w}x{h} primary rays vs {arrays['indices'].size // 3}-triangle displaced-grid terrain, SBVH "
                               f"(BASELINE configs[1]); CPU restatement of the reference OpenCL kernel (pocl unavailable in image)",
                   "rays_per_step": nrays},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"the full {w}x{h} frame ({nrays} traversed rays) per step, {len(timed)} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_opencl": reference_opencl_frame(arrays, bvh, params, w, h),
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: gather the framebuffer with NCCL all_gather instead of peer stores")
    ap.add_argument("--extra", action="store_true", help="also time shadow / diffuse / frame passes (N=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import rtb200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, h = frame_size(world)
    n_pix = w * h
    ctx = rtb200.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- scene: built once on rank 0's host, broadcast as ONE device buffer -----------------------
    arrays = bvh = None
    build_s = bcast_ms = 0.0
    if rank == 0:
        _mesh, arrays, bvh, build_s = build_scene()
        ctx.upload_scene(arrays, bvh.nodes, bvh.tri_indices)
        params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"])
    else:
        params = np.zeros(32, dtype=np.float32)
    if world > 1:
        from rtb200 import tiling

        pt = torch.from_numpy(params).cuda()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blob = tiling.broadcast_scene(dist, ctx, rank, "cuda")  # NCCL over NVLink: the packed BVH + triangles, one buffer
        dist.broadcast(pt, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        params = pt.cpu().numpy()
    ctx.set_params(params)
    ctx_params[0] = params

    # ---- buffers ---------------------------------------------------------------------------------
    d_hits = torch.zeros((n_pix, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    if world > 1:
        # per-rank band buffer (4 B/pixel hit index) -> all_gather -> frame
        owned_rows = list(tiling.owned_rows(rank, world, h, BAND_ROWS))
        rows_per_rank = tiling.rows_per_rank(world, h, BAND_ROWS)
        band_idx = tiling.band_index(rank, world, h, BAND_ROWS, device="cuda")
        d_band = torch.empty((rows_per_rank, w), dtype=torch.int32, device="cuda")
        d_gather = torch.empty((world, rows_per_rank, w), dtype=torch.int32, device="cuda")
        # preferred: gather fused into the kernel's store (peer-mapped framebuffer on rank 0, NVLink);
        # fallback if CUDA IPC is unavailable: band pack + NCCL all_gather
        shared_frame = None
        if not args.nccl_gather:
            try:
                shared_frame = tiling.open_shared_frame(dist, ctx, rank, n_pix * 4)
            except Exception as exc:  # noqa: BLE001
                shared_frame = None
                if rank == 0:
                    print(f"bench: CUDA IPC framebuffer unavailable ({exc}); using NCCL all_gather", file=sys.stderr)
            flag = torch.tensor([1 if shared_frame else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()):
                shared_frame = None
    else:
        owned_rows = list(range(h))
        shared_frame = None

    def step():
        if shared_frame:
            ctx.primary_gather_device(w, h, d_hits, shared_frame, part=rank, n_parts=world, band_rows=BAND_ROWS)
            return
        ctx.primary_device(w, h, d_hits, None, part=rank, n_parts=world, band_rows=BAND_ROWS)
        if world > 1:
            with torch.cuda.stream(stream):
                torch.index_select(d_hits.view(torch.int32).view(h, w, 4)[:, :, 0], 0, band_idx, out=d_band)
                tiling.gather_bands(dist, d_band, out=d_gather)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    sampler.wait_first()
    # keep the GPU under the same load until the sampler has started delivering, then open the window
    sampler.mark()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ------
    ctx.reset_counters()
    step_ms = []
    barrier()
    for _ in range(args.steps):
        if not args.no_flush:
            flush.fill_(1)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = ctx.counters()["kernel_launches"]
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # rays per step = pixels passing the scene gate (traversals), summed over ranks
    rays_local = int(count_gate_pass(params, w, h, owned_rows))
    if world > 1:
        t = torch.tensor([rays_local], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        rays_total = int(t.item())
    else:
        rays_total = rays_local
    value = rays_total * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: the same pass through the host-buffer C-ABI call (params in, hit records out) --------
    pinned = torch.empty((len(owned_rows) * w if world > 1 else n_pix, 4), dtype=torch.float32).pin_memory()
    e2e = None
    if world == 1:
        ctx.set_stream(None)
        for _ in range(3):
            ctx.set_params(params)
            ctx.primary(w, h, pinned)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.set_params(params)          # host -> device: the 128-byte Params block of this frame
            ctx.primary(w, h, pinned)       # device -> host: w*h 16-byte hit records (pinned memory)
        e2e_s = time.perf_counter() - t0
        e2e = {"value": rays_total * args.steps / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 128,
               "d2h_bytes_per_step": n_pix * 16, "call": "rt_set_params + rt_primary (host buffers, pinned)"}
        # the same pass with the result the N > 1 runs deliver: the 4-byte/pixel hit-index framebuffer (the size of the
        # reference's own per-frame readback, RayTracer.cpp:343), stored by the kernel straight into pinned host memory
        idx_frame = torch.empty((h, w), dtype=torch.int32).pin_memory()
        for _ in range(3):
            ctx.primary_gather_device(w, h, None, idx_frame, part=0, n_parts=1, band_rows=BAND_ROWS)
            ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.set_params(params)
            ctx.primary_gather_device(w, h, None, idx_frame, part=0, n_parts=1, band_rows=BAND_ROWS)
            ctx.synchronize()
        e2e_idx_s = time.perf_counter() - t0
        if not np.array_equal(idx_frame.numpy().reshape(-1), pinned.numpy().view(np.int32)[:, 0]):
            raise SystemExit("bench: hit-index framebuffer differs from the hit records")
        e2e["index_frame"] = {"value": rays_total * args.steps / e2e_idx_s / 1e6, "unit": UNIT, "d2h_bytes_per_step": n_pix * 4,
                              "call": "rt_set_params + rt_primary_gather_device (4-byte/pixel hit-index frame in pinned host memory): the result the N > 1 runs deliver"}
        ctx.set_stream(stream.cuda_stream)
    else:
        # N > 1: params from host each step, own bands traced, framebuffer gathered, rank 0 reads the frame back
        # preferred e2e path: the framebuffer is POSIX shared memory that every rank page-locks; each rank's kernel stores
        # its bands straight into rank 0's HOST memory over its own PCIe link. Fallbacks: peer-mapped device frame + D2H
        # on rank 0, then NCCL all_gather + D2H.
        host_frame = None
        if not args.nccl_gather:
            try:
                host_frame = tiling.HostFrame(dist, ctx, rank, h, w)
            except Exception as exc:  # noqa: BLE001
                host_frame = None
                if rank == 0:
                    print(f"bench: shared-memory host framebuffer unavailable ({exc})", file=sys.stderr)
            flag = torch.tensor([1 if host_frame else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()) and host_frame:
                host_frame.close()
                host_frame = None
        pinned_frame = (torch.empty((h, w), dtype=torch.int32) if shared_frame else torch.empty((world, rows_per_rank, w), dtype=torch.int32)).pin_memory()
        done = torch.zeros(1, device="cuda")

        def e2e_step():
            ctx.set_params(params)  # host -> device: this frame's 128-byte Params block, on every rank
            with torch.cuda.stream(stream):
                if host_frame:
                    ctx.primary_gather_device(w, h, None, host_frame.device_alias, part=rank, n_parts=world, band_rows=BAND_ROWS)
                else:
                    step()
                    if shared_frame:
                        dist.all_reduce(done)  # frame-complete signal: every rank's stores have landed on rank 0
                        if rank == 0:
                            ctx.memcpy_to_host(pinned_frame, shared_frame, n_pix * 4)
                    elif rank == 0:
                        pinned_frame.copy_(d_gather, non_blocking=True)
            stream.synchronize()
            if host_frame:
                dist.barrier()  # every rank's kernel has finished: the frame in rank 0's host memory is complete

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if host_frame:
            call = "rt_set_params + rt_primary_gather_device(bands) storing into a shared-memory host framebuffer page-locked by every rank (each GPU's own PCIe link) + barrier"
            if rank == 0:  # sanity: the host frame equals the device-side gathered frame
                with torch.cuda.stream(stream):
                    step()
                torch.cuda.synchronize()
            dist.barrier()
            if rank == 0 and shared_frame:
                ctx.memcpy_to_host(pinned_frame, shared_frame, n_pix * 4)
                if not np.array_equal(pinned_frame.numpy(), host_frame.array):
                    raise SystemExit("bench: host framebuffer differs from the device framebuffer")
        elif shared_frame:
            call = "rt_set_params + rt_primary_gather_device(bands, peer-mapped frame on rank 0) + 1-element all_reduce + frame D2H on rank 0"
        else:
            call = "rt_set_params + rt_primary_device(bands) + NCCL all_gather + frame D2H on rank 0"
        e2e = {"value": rays_total * args.steps / float(t.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 128 * world,
               "d2h_bytes_per_step": int(n_pix * 4), "call": call}
        if host_frame:
            host_frame.close()

    # clocks: the sampling window covers the K timed steps and the e2e loop; if the timed steps were shorter than
    # two sampling periods, extend the window with more (untimed) steps of the same load
    with torch.cuda.stream(stream):
        t_end = time.time() + 0.4
        while time.time() < t_end:
            step()
            stream.synchronize()
    clocks = sampler.stop()

    # ---- rank 0: CPU baseline + roofline, then the JSON line ---------------------------------------
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        cpu = roof = None
        w1, h1 = frame_size(1)
        if world == 1:
            times, cnt, nrays, threads = oracle_pass(arrays, bvh, params, w1, h1, repeats=5)
            cpu_val = nrays / float(np.median(times)) / 1e6
            cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"the full {w1}x{h1} frame ({nrays} traversed rays), median of 5 passes; CPU restatement of the reference OpenCL kernel (pocl unavailable)"}
            # algorithmic bytes per launch: 64 B per inner-node visit + 48 B per triangle test (counted by the
            # oracle for this exact ray set) + 16 B hit record per pixel
            algo_bytes = 64 * cnt["inner"] + 48 * cnt["tris"] + 16 * n_pix
            kernel_ms = float(np.mean(step_ms))
            achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": load_traffic(), "peak_source": peak_src, "kernel": "trace_kernel<SRC_PRIMARY,closest>",
                    "algorithmic_bytes_per_launch": int(algo_bytes), "inner_visits_per_ray": cnt["inner"] / nrays,
                    "tri_tests_per_ray": cnt["tris"] / nrays,
                    "note": "working set (~80 MB touched) is L2-resident after first touch: the real limiter is L2/L1 latency and SIMT divergence, not HBM"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{w}x{h} primary rays ({rays_total} traversed/step) vs 999698-triangle displaced-grid terrain, SBVH via SplitBVHBuilder "
                                   f"(BASELINE configs[1]{'' if world == 1 else '; frame scaled with N at 16:9, interleaved 16-row bands, NCCL scene broadcast + framebuffer all_gather'})",
                       "frame": [w, h], "rays_per_step": rays_total, "pixels_per_step": n_pix,
                       "mpixels_per_s": n_pix * args.steps / (total_ms * 1e-3) / 1e6, "l2": "not flushed" if args.no_flush else "flushed between timed steps (256 MiB fill)",
                       "scene_build_s": build_s, "scene_broadcast_ms": bcast_ms, "band_rows": BAND_ROWS,
                       "gather": "n/a (1 GPU)" if world == 1 else ("fused into the kernel store: peer-mapped framebuffer on rank 0 over NVLink (CUDA IPC)" if shared_frame else "NCCL all_gather of 4-byte/pixel bands")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        }
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        if world == 1:
            # same-hardware comparison on the reference's own unit of work (one shaded frame): its kernel vs rt_render_frame
            ref_cl = reference_opencl_frame(arrays, bvh, params, w, h)
            d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            ts = []
            for i in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record()
                    ctx.render_frame_device(w, h, d_img)
                    e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ref_cl["ours_frame_kernel_ms"] = float(np.median(ts))
            line["reference_opencl"] = ref_cl
        if args.extra and world == 1:
            line["extra"] = extra_passes(ctx, torch, rtb200, w, h, stream)
        print(json.dumps(line))
    if world > 1:
        if shared_frame:
            tiling.close_shared_frame(dist, ctx, rank, shared_frame)
        dist.barrier()
        dist.destroy_process_group()


def count_gate_pass(params, w, h, rows):
    """Pixels whose primary ray passes the scene-AABB gate (numpy restatement of the gate, used only to COUNT rays)."""
    p = np.asarray(params, dtype=np.float32).reshape(8, 4)[:, :3]
    a, b, c, campos, bmin, bmax = p[0], p[1], p[2], p[3], p[6], p[7]
    ys = np.asarray(rows, dtype=np.float32)
    xf = ((np.arange(w, dtype=np.float32) - np.float32(0.5)) / np.float32(w)).astype(np.float32)
    yf = ((ys - np.float32(0.5)) / np.float32(h)).astype(np.float32)
    pos = c[None, None, :] + a[None, None, :] * xf[None, :, None] + b[None, None, :] * yf[:, None, None]
    d = pos - campos[None, None, :]
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        l1, l2 = (bmin - pos) * inv, (bmax - pos) * inv
        tmin = np.fmax.reduce(np.fmin(l1, l2), axis=-1)
        tmax = np.fmin.reduce(np.fmax(l1, l2), axis=-1)
    return int(np.count_nonzero((tmax >= tmin) & (tmax >= 0)))


def load_traffic():
    """dram bytes per launch of the primary kernel from the committed ncu capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "primary_kernel_traffic.json")))["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        return None


ctx_params = [None]


def extra_passes(ctx, torch, rtb200, w, h, stream):
    """shadow / diffuse / frame throughput on the same scene (informational; not the headline metric)."""
    n = w * h
    out = {}

    def timed(fn, iters=10):
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                fn()
                e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    d_hits = torch.zeros((n, 4), device="cuda")
    d_rays = torch.zeros((n, 8), device="cuda")
    d_sh = torch.zeros((n, 4), device="cuda")
    ctx.primary_device(w, h, d_hits, d_rays)
    torch.cuda.synchronize()
    nhit = int((d_hits.view(torch.int32)[:, 0] >= 0).sum().item())
    ms = timed(lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh))
    out["shadow_anyhit_mrays_s"] = nhit / ms / 1e3
    d_dr = torch.zeros((n * 4, 8), device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.diffuse_rays_device(n, d_rays, d_hits, 4, 0x5EED, d_dr, d_cnt)
    torch.cuda.synchronize()
    nd = int(d_cnt.item())
    d_dh = torch.zeros((nd, 4), device="cuda")
    ms = timed(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh))
    out["diffuse_4spp_mrays_s_on_1M_scene"] = nd / ms / 1e3
    d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    ms = timed(lambda: ctx.render_frame_device(w, h, d_img), iters=5)
    out["full_frame_ms"] = ms  # default: the megakernel (frame_mode 0)
    ctx.set_option("frame_mode", 1)
    out["full_frame_wavefront_ms"] = timed(lambda: ctx.render_frame_device(w, h, d_img), iters=5)
    ctx.set_option("frame_mode", 0)
    # the frame through the host entry points (pinned frame buffers, wall clock over 50 frames, params set every frame)
    import time
    bufs = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
    nf = 50
    for _ in range(3):
        ctx.render_frame(w, h, bufs[0])
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(ctx_params[0])
        ctx.render_frame(w, h, bufs[0])
    out["frame_e2e_ms_frame_by_frame"] = (time.perf_counter() - t0) / nf * 1e3
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(ctx_params[0])
        ctx.render_frame_begin(w, h, bufs[k & 1], k & 1)
        if k:
            ctx.render_frame_end((k - 1) & 1)
    ctx.render_frame_end((nf - 1) & 1)
    out["frame_e2e_ms_two_in_flight"] = (time.perf_counter() - t0) / nf * 1e3
    # NOT the shipped default: closest-hit traversal with the reciprocal-multiply box test (not bit-exact)
    ctx.set_option("fast_box", 1)
    ms = timed(lambda: ctx.primary_device(w, h, d_hits))
    ctx.set_option("fast_box", 0)
    ntrav = count_gate_pass(ctx_params[0], w, h, list(range(h)))
    out["primary_mrays_s_with_inexact_fast_box_option"] = ntrav / ms / 1e3
    return out


if __name__ == "__main__":
    main()
