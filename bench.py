#!/usr/bin/env python
"""bench.py -- headline benchmark: Mrays/s of primary rays (fused camera ray generation + closest-hit
SBVH traversal + ray/triangle intersection) on BASELINE.json configs[1]:
1920x1080 primary rays vs the ~1M-triangle displaced-grid terrain, SBVH, one B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extra]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

A step = one pass of the hot path over one frame of primary rays. Rays counted = traversals completed
(pixels that pass the scene-AABB gate call traverse_bvh once; the others do not, as in the reference); the count is
the kernels' own (RT_CNT_RAYS_TRACED).
N > 1: weak scaling -- the frame grows with N at fixed aspect (N = 4 is exactly the 3840x2160 frame
of configs[4]) so every rank keeps ~2.07 M pixels, partitioned in interleaved 16-row bands; the scene
is built on rank 0's host and broadcast once as one device buffer (NCCL); every step ends with the
gather of the 4-byte/pixel framebuffer (hit triangle index) -- fused into the kernels' stores (peer-mapped frame on
rank 0), the only data-path exchange.

Besides the headline line (unchanged metric), rank 0 adds, outside the timed headline region:
  N = 1: "extra" -- the rest of BASELINE's metric, each with an oracle-counted roofline and an oracle parity check:
         configs[2] shadow rays (default + grazing light, two-pass and fused primary+shadow), configs[3] incoherent
         diffuse rays (4 spp) on the 10 M-triangle scene, configs[0] the 81 920-triangle sphere through a .dae file.
  N > 1: "config5_strong" -- configs[4]: ONE fixed 3840x2160 frame (fused primary+shadow pass, and the shaded frame)
         split over the N GPUs: ms (max over ranks), speed-up over the same frame on one GPU, per-rank kernel ms.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s primary (1080p, 1M-tri terrain, SBVH)"
UNIT = "Mrays/s"
TERRAIN_QUADS = 707  # 999 698 triangles
TERRAIN_TRIS = 2 * TERRAIN_QUADS * TERRAIN_QUADS
BAND_ROWS = 16
GRAZING_LIGHT = (-150.0, 25.0, 3.0)


def frame_size(n_gpus):
    """~2.07 M pixels per GPU at 16:9; w % 8 == 0, h % 4 == 0 (the kernel's 8x4 warp tile)."""
    s = float(n_gpus) ** 0.5
    return int(round(1920 * s / 8)) * 8, int(round(1080 * s / 4)) * 4


def workload_string(w, h):
    """config.workload -- the SAME sentence in both arms (the driver compares it)"""
    return (f"{w}x{h} primary rays vs {TERRAIN_TRIS}-triangle displaced-grid terrain, SBVH via SplitBVHBuilder "
            f"(BASELINE configs[1]; frame = 1920x1080 x sqrt(n_gpus) at 16:9)")


def build_scene():
    import rtb200

    rtb200.hostlib.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1; only rank 0 builds
    t0 = time.time()
    mesh = rtb200.Mesh().terrain(TERRAIN_QUADS, 100.0).finish(diffuse=(0.7, 0.7, 0.7))
    arrays = mesh.arrays()
    bvh = rtb200.FlatBVH.build(mesh)
    return mesh, arrays, bvh, time.time() - t0


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md) through NVML, in process, once per
    timed step right after the step's kernel has finished -- inside the region, but not while a timed kernel runs. (A
    concurrent `nvidia-smi -lms 50` loop, the first implementation, cost the 0.2 ms kernel 3-4 %: 0.206 -> 0.213-0.217 ms,
    tools/headline_probe.py. It is kept as the fallback when NVML cannot be loaded, at the recipe's 200 ms.)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.nvml, self.handle, self.first = gpu_index, [], None, None, None, 0

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            import torch

            try:  # CUDA_VISIBLE_DEVICES may renumber the devices: go through the PCI address of torch's device
                pr = torch.cuda.get_device_properties(self.gpu)
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = pynvml
            self.sample()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def sample(self):
        """one NVML sample (no-op in the nvidia-smi fallback, which polls by itself)"""
        if not self.nvml:
            return
        n = self.nvml
        try:
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            flag = lambda name: "Active" if (mask & getattr(n, name, 0)) else "Not Active"  # noqa: E731
            self.rows.append([str(self.gpu), str(sm), str(mx), "", "", flag("nvmlClocksEventReasonHwSlowdown"),
                              flag("nvmlClocksEventReasonHwThermalSlowdown"), flag("nvmlClocksEventReasonSwThermalSlowdown"),
                              flag("nvmlClocksEventReasonSwPowerCap")])
        except Exception:  # noqa: BLE001
            pass

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
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "how": "NVML, one sample per timed step right after its kernel" if self.nvml else "nvidia-smi -lms 200 during the timed region"}


def oracle_pass(arrays, bvh, params, w, h, repeats, threads=None):
    """CPU restatement of the reference kernel (oracle/) on the same frame: seconds per pass + visit counts."""
    from oracle import oracle_py as O

    if threads is None:  # every core this process may run on; torchrun exports OMP_NUM_THREADS=1, which would leave one
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
    w, h = frame_size(max(1, args.gpus))
    params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"])
    times, cnt, nrays, threads = oracle_pass(arrays, bvh, params, w, h, args.warmup + args.steps)
    timed = times[args.warmup:]
    sec = float(np.sum(timed))
    value = nrays * len(timed) / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / len(timed) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(w, h), "frame": [w, h], "rays_per_step": nrays,
                   "implementation": "CPU restatement of the reference OpenCL kernel (oracle/oracle.c; pocl unavailable in the image), all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"the full {w}x{h} frame ({nrays} traversed rays) per step, {len(timed)} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_opencl": reference_opencl_frame(arrays, bvh, params, w, h) if args.gpus <= 1 else None,
    }
    print(json.dumps(line))


class Timer:
    """CUDA-event timing of fn() on `stream`; L2 flushed (256 MiB fill) before every timed launch unless flush is None."""

    def __init__(self, torch, stream, flush):
        self.torch, self.stream, self.flush = torch, stream, flush

    def __call__(self, fn, iters=10, warm=3):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warm):
                fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            if self.flush is not None:
                self.flush.fill_(1)
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self.stream):
                e0.record()
                fn()
                e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return ts


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_of(cnt, io_bytes, ms, kernel, resident, capture=None):
    """algorithmic bytes = 64 B x inner-node visits + 48 B x triangle tests (both counted by the ORACLE for
    this exact ray set) + ray / pixel I/O; achieved = bytes / measured kernel time; peak = measured HBM copy bandwidth."""
    peak, src = peak_hbm()
    algo = 64 * cnt["inner"] + 48 * cnt["tris"] + io_bytes
    achieved = algo / (ms * 1e-3) / 1e9
    cap = load_capture(capture) if capture else {}
    return {"bound": "l2-resident (latency)" if resident else "hbm-resident scene (L2-miss latency + divergence; DRAM bandwidth 7-12 % of peak)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": cap.get("dram_bytes_per_launch"), "l2_bytes_per_launch": cap.get("l2_bytes_per_launch"),
            "l1_bytes_per_launch": cap.get("l1_bytes_per_launch"), "capture": cap.get("source"),
            "peak_source": src, "kernel": kernel, "algorithmic_bytes_per_launch": int(algo)}


def load_capture(name):
    """per-launch DRAM / L2 / L1 bytes of a kernel from the committed ncu capture summary (profiles/<name>.json), or {}"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name + ".json")))
    except (OSError, ValueError):
        return {}


def same_hits(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in ("idx", "t", "u", "v"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: gather the framebuffer with NCCL all_gather instead of peer stores")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra / config5_strong blocks (quick runs, ncu)")
    ap.add_argument("--skip-c4", action="store_true", help="extra block without the 10 M-triangle scene")
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
    blob_bytes = 0
    if rank == 0:
        _mesh, arrays, bvh, build_s = build_scene()
        ctx.upload_scene(arrays, bvh.nodes, bvh.tri_indices)
        params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"])
        aabb = np.concatenate([arrays["aabb_min"], arrays["aabb_max"]]).astype(np.float32)
    else:
        params = np.zeros(32, dtype=np.float32)
        aabb = np.zeros(6, dtype=np.float32)
    if world > 1:
        from rtb200 import tiling

        pt = torch.from_numpy(np.concatenate([params, aabb])).cuda()
        dist.broadcast(pt, src=0)
        both = pt.cpu().numpy()
        params, aabb = both[:32].copy(), both[32:].copy()
        # the NCCL broadcast of the packed scene, timed alone (buffer allocated and sized beforehand)
        meta = [ctx.scene_blob()[1] if rank == 0 else 0]
        dist.broadcast_object_list(meta, src=0)
        blob_bytes = int(meta[0])
        blob = torch.empty(blob_bytes, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ctx.copy_scene_blob(blob, blob_bytes)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(blob, src=0)  # NCCL over NVLink: node pairs + triangles + shading arrays, one buffer
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bcast_ms = float(t.item())
        if rank != 0:
            ctx.adopt_scene_blob(blob.data_ptr(), blob_bytes)
    ctx.set_params(params)

    # ---- buffers ---------------------------------------------------------------------------------
    d_hits = torch.zeros((n_pix, 4), dtype=torch.float32, device="cuda")
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    shared_frame = None
    if world > 1:
        # per-rank band buffer (4 B/pixel hit index) -> all_gather -> frame
        rows_per_rank = tiling.rows_per_rank(world, h, BAND_ROWS)
        band_idx = tiling.band_index(rank, world, h, BAND_ROWS, device="cuda")
        d_band = torch.empty((rows_per_rank, w), dtype=torch.int32, device="cuda")
        d_gather = torch.empty((world, rows_per_rank, w), dtype=torch.int32, device="cuda")
        # preferred: gather fused into the kernel's store (peer-mapped framebuffer on rank 0, NVLink);
        # fallback if CUDA IPC is unavailable: band pack + NCCL all_gather
        if not args.nccl_gather:
            try:
                shared_frame = tiling.open_shared_frame(dist, ctx, rank, max(n_pix, 3840 * 2160) * 4)
            except Exception as exc:  # noqa: BLE001
                shared_frame = None
                if rank == 0:
                    print(f"bench: CUDA IPC framebuffer unavailable ({exc}); using NCCL all_gather", file=sys.stderr)
            flag = torch.tensor([1 if shared_frame else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()):
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

    def allsum(v):
        if world == 1:
            return int(v)
        t = torch.tensor([int(v)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):  # warm-up steps run like the timed ones (L2 flushed before each): the tile hints a
        if flush is not None:               # launch records are then those of the conditions that are timed
            flush.fill_(1)
            torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            step()
    barrier()
    # rays per step = traversals the kernels start (their own counter), summed over ranks
    ctx.reset_counters()
    if flush is not None:
        flush.fill_(1)
        torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        step()
    rays_total = allsum(ctx.counters()["rays_traced"])
    sampler.wait_first()
    # keep the GPU under the same load until the sampler has started delivering, then open the window
    sampler.mark()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ------
    ctx.reset_counters()
    step_ms = []
    barrier()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        sampler.sample()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = ctx.counters()["kernel_launches"]
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = rays_total * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: the same pass through the host-buffer C-ABI call (params in, results out) ------------
    e2e = None
    if world == 1:
        pinned = torch.empty((n_pix, 4), dtype=torch.float32).pin_memory()
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
        # in assembled 512-byte rows
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
                              "ms_per_step": e2e_idx_s / args.steps * 1e3,
                              "call": "rt_set_params + rt_primary_gather_device (4-byte/pixel hit-index frame, rows assembled to 512-byte stores into pinned host memory) + rt_synchronize: the result the N > 1 runs deliver"}
        ctx.set_stream(stream.cuda_stream)
    else:
        # N > 1: params from host each step, own bands traced, every rank's kernel stores its bands (assembled 512-byte rows)
        # straight into a shared-memory HOST frame that all ranks page-lock: each GPU uses its own PCIe link, rank 0 needs no
        # device->host copy. Completion: one sequence word per rank (rt_signal behind the kernel) polled by rank 0 -- no
        # barrier per frame; two frame buffers are cycled and rank 0 acknowledges the frames it has seen complete.
        host_frame = None
        if not args.nccl_gather:
            try:
                host_frame = tiling.HostFrame(dist, ctx, rank, h, w, n_frames=2, band_rows=BAND_ROWS)
            except Exception as exc:  # noqa: BLE001
                host_frame = None
                if rank == 0:
                    print(f"bench: shared-memory host framebuffer unavailable ({exc})", file=sys.stderr)
            flag = torch.tensor([1 if host_frame else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not int(flag.item()) and host_frame:
                host_frame.close()
                host_frame = None
        if host_frame:
            seq = [0]

            def e2e_frames(k):
                for _ in range(k):
                    seq[0] += 1
                    s = seq[0]
                    if s > 2:
                        host_frame.wait_ack(s - 2)   # the buffer about to be overwritten has been taken by the consumer
                    ctx.set_params(params)           # host -> device: this frame's 128-byte Params block, on every rank
                    ctx.primary_gather_device(w, h, None, host_frame.frame_alias(s), part=rank, n_parts=world, band_rows=BAND_ROWS)
                    host_frame.signal(s)
                    if rank == 0:
                        host_frame.wait_frame(s)     # the complete frame is in rank 0's host memory
                if rank != 0:
                    host_frame.wait_ack(seq[0])

            e2e_frames(3)
            barrier()
            t0 = time.perf_counter()
            e2e_frames(args.steps)
            dt = time.perf_counter() - t0
            barrier()
            call = ("rt_set_params + rt_primary_gather_device(bands) storing assembled 512-byte rows into a shared-memory host frame "
                    "page-locked by every rank (each GPU's own PCIe link) + rt_signal; rank 0 polls one sequence word per rank, no barrier per frame")
            # sanity: the host frame equals the device-side gathered frame
            with torch.cuda.stream(stream):
                step()
            barrier()
            if rank == 0 and shared_frame:
                check = torch.empty((h, w), dtype=torch.int32).pin_memory()
                ctx.memcpy_to_host(check, shared_frame, n_pix * 4)
                if not np.array_equal(check.numpy(), host_frame.frames[seq[0] % 2]):
                    raise SystemExit("bench: host framebuffer differs from the device framebuffer")
        else:
            pinned_frame = (torch.empty((h, w), dtype=torch.int32) if shared_frame else torch.empty((world, rows_per_rank, w), dtype=torch.int32)).pin_memory()
            done = torch.zeros(1, device="cuda")

            def e2e_step():
                ctx.set_params(params)
                with torch.cuda.stream(stream):
                    step()
                    if shared_frame:
                        dist.all_reduce(done)  # frame-complete signal: every rank's stores have landed on rank 0
                        if rank == 0:
                            ctx.memcpy_to_host(pinned_frame, shared_frame, n_pix * 4)
                    elif rank == 0:
                        pinned_frame.copy_(d_gather, non_blocking=True)
                stream.synchronize()

            for _ in range(3):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            dt = time.perf_counter() - t0
            barrier()
            call = ("rt_set_params + rt_primary_gather_device(bands, peer-mapped frame on rank 0) + 1-element all_reduce + frame D2H on rank 0"
                    if shared_frame else "rt_set_params + rt_primary_device(bands) + NCCL all_gather + frame D2H on rank 0")
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": rays_total * args.steps / float(t.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 128 * world,
               "d2h_bytes_per_step": int(n_pix * 4), "ms_per_step": float(t.item()) / args.steps * 1e3, "call": call}
        if host_frame:
            numa = [host_frame.numa]
            dist.broadcast_object_list(numa, src=0)
            e2e["host_frame_numa"] = numa[0]
            host_frame.close()

    # clocks: the sampling window covers the K timed steps and the e2e loop; if the timed steps were shorter than
    # two sampling periods, extend the window with more (untimed) steps of the same load
    if not sampler.nvml:  # (the nvidia-smi fallback polls every 200 ms; NVML samples were taken step by step)
        with torch.cuda.stream(stream):
            t_end = time.time() + 0.8
            while time.time() < t_end:
                step()
                stream.synchronize()
    clocks = sampler.stop()

    timer = Timer(torch, stream, flush)
    strong = None
    if world > 1 and not args.no_extra:
        strong = config5_strong(ctx, torch, dist, rtb200, rank, world, timer, shared_frame, aabb)

    # ---- rank 0: CPU baseline + roofline, then the JSON line ---------------------------------------
    if rank == 0:
        cpu = roof = None
        w1, h1 = frame_size(1)
        if world == 1:
            times, cnt, nrays, threads = oracle_pass(arrays, bvh, params, w1, h1, repeats=5)
            cpu_val = nrays / float(np.median(times)) / 1e6
            cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"the full {w1}x{h1} frame ({nrays} traversed rays), median of 5 passes; CPU restatement of the reference OpenCL kernel (pocl unavailable)"}
            if nrays != rays_total:
                raise SystemExit(f"bench: the kernels counted {rays_total} traversals, the oracle's gate passes {nrays}")
            roof = roofline_of(cnt, 16 * n_pix, float(np.mean(step_ms)), "trace_kernel<SRC_PRIMARY,closest>", True, "primary_kernel_traffic")
            roof.update({"inner_visits_per_ray": cnt["inner"] / nrays, "tri_tests_per_ray": cnt["tris"] / nrays,
                         "note": "the scene's traversal working set (~71 MB) is L2-resident after first touch: the algorithmic bytes are served by "
                                 "L1 and L2, DRAM sees `traffic`; the limiter is L1/L2 latency and SIMT divergence, not HBM. `frac` is the "
                                 "algorithmic-byte rate over the measured HBM copy bandwidth, as SURVEY 8(d) defines it."})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(w, h), "frame": [w, h], "rays_per_step": rays_total, "pixels_per_step": n_pix,
                       "mpixels_per_s": n_pix * args.steps / (total_ms * 1e-3) / 1e6, "l2": "not flushed" if args.no_flush else "flushed between timed steps (256 MiB fill)",
                       "step_ms_min_median_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))],
                       "tile_scheduler": "temporal tile hints + scene-box culling of the tile queue (DESIGN 3.6); warm-up steps run under the timed conditions",
                       "scene_build_s": build_s, "scene_broadcast_ms": bcast_ms, "scene_blob_bytes": blob_bytes,
                       "scene_broadcast": "n/a (1 GPU)" if world == 1 else "one NCCL broadcast of the packed scene buffer, timed alone with CUDA events (max over ranks)",
                       "partition": "n/a (1 GPU)" if world == 1 else f"interleaved {BAND_ROWS}-row bands, one process per GPU",
                       "gather": "n/a (1 GPU)" if world == 1 else ("fused into the kernel store: peer-mapped framebuffer on rank 0 over NVLink (CUDA IPC), rows assembled to 128-byte stores"
                                                                    if shared_frame else "band pack + NCCL all_gather of 4-byte/pixel bands")},
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
            ts = timer(lambda: ctx.render_frame_device(w, h, d_img), iters=5)
            ref_cl["ours_frame_kernel_ms"] = float(np.median(ts))
            line["reference_opencl"] = ref_cl
            if not args.no_extra:
                line["extra"] = extra_block(ctx, torch, rtb200, timer, arrays, bvh, w, h, args)
        if strong:
            line["config5_strong"] = strong
        print(json.dumps(line))
    if world > 1:
        if shared_frame:
            tiling.close_shared_frame(dist, ctx, rank, shared_frame)
        dist.barrier()
        dist.destroy_process_group()


# ======================================================================================================================
# N > 1: BASELINE configs[4] -- ONE fixed 3840x2160 frame split over the N GPUs (strong scaling)
# ======================================================================================================================
def config5_strong(ctx, torch, dist, rtb200, rank, world, timer, shared_frame, aabb):
    W, H = 3840, 2160
    n = W * H
    params, _ = rtb200.camera_params(W, H, aabb[:3], aabb[3:])
    ctx.set_params(params)
    stream, flush = timer.stream, timer.flush
    local = torch.zeros((H, W), dtype=torch.int32, device="cuda")       # rank 0: the same frame on ONE GPU
    own_gather = None if shared_frame else torch.zeros((H, W), dtype=torch.int32, device="cuda")
    target = shared_frame if shared_frame else own_gather

    def launch(kind, dst, part, n_parts, band_rows):
        if kind == "primary_shadow":
            ctx.primary_shadow_device(W, H, None, None, dst, part=part, n_parts=n_parts, band_rows=band_rows)
        else:
            ctx.render_frame_device(W, H, dst, part=part, n_parts=n_parts, band_rows=band_rows)

    def gathered(vals):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])  # (world, iters)

    out = {"frame": [W, H], "scene": "the bench scene (999 698-triangle terrain), default camera and light",
           "timing": "CUDA events around each rank's launch, L2 flushed before every launch, median over 9 launches of the max over ranks; "
                     "speedup = the same frame on one GPU (rank 0 alone) / that",
           "output": "4-byte/pixel frame, every rank storing its bands into rank 0's frame over NVLink (peer mapping)" if shared_frame else
                     "4-byte/pixel frame, band stores into a local frame (no peer mapping available)"}
    for kind in ("primary_shadow", "shaded_frame"):
        # one GPU, whole frame
        n1 = None
        rays = 0
        if rank == 0:
            ctx.reset_counters()
            n1 = float(np.median(timer(lambda: launch(kind, local, 0, 1, 16), iters=7)))
            rays = ctx.counters()["rays_traced"] // 10  # 3 warm-up + 7 timed launches
        dist.barrier()
        res = {"one_gpu_ms": n1, "rays_per_frame": rays}
        for band_rows in (16, 4):
            ms = gathered(timer(lambda: launch(kind, target, rank, world, band_rows), iters=9))
            per_iter_max = ms.max(axis=0)
            dist.barrier()
            entry = {"ms": float(np.median(per_iter_max)), "per_rank_ms": [float(v) for v in np.median(ms, axis=1)]}
            if rank == 0:
                entry["speedup"] = n1 / entry["ms"]
                entry["mrays_s"] = rays / entry["ms"] / 1e3
                if shared_frame:
                    chk = torch.empty((H, W), dtype=torch.int32).pin_memory()
                    ctx.memcpy_to_host(chk, shared_frame, n * 4)
                    entry["gathered_frame_identical_to_one_gpu"] = bool(np.array_equal(chk.numpy(), local.cpu().numpy()))
                    if not entry["gathered_frame_identical_to_one_gpu"]:
                        raise SystemExit(f"bench: config5 {kind}: the gathered frame differs from the one-GPU frame")
            dist.barrier()
            res[f"bands_of_{band_rows}_rows"] = entry
        # sustained: 20 frames back to back on every rank, no synchronisation between frames (disjoint rows)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            for _ in range(20):
                launch(kind, target, rank, world, 16)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["back_to_back_ms_per_frame"] = float(t.item()) / 20 * 1e3
        out[kind] = res
    best = min(("bands_of_16_rows", "bands_of_4_rows"), key=lambda k: out["primary_shadow"][k]["ms"])
    out["summary"] = {"primary_shadow_ms": out["primary_shadow"][best]["ms"], "primary_shadow_speedup": out["primary_shadow"][best].get("speedup"),
                      "shaded_frame_ms": min(out["shaded_frame"][k]["ms"] for k in ("bands_of_16_rows", "bands_of_4_rows")),
                      "shaded_frame_speedup": max((out["shaded_frame"][k].get("speedup") or 0) for k in ("bands_of_16_rows", "bands_of_4_rows")) or None,
                      "partition": best}
    return out


# ======================================================================================================================
# N = 1: the rest of BASELINE's metric (shadow, diffuse, C1, C4), each with an oracle-counted roofline + parity check
# ======================================================================================================================
def extra_block(ctx, torch, rtb200, timer, arrays, bvh, w, h, args):
    from oracle import oracle_py as O

    HIT = rtb200.HIT_DTYPE
    out = {"note": "same timing discipline as the headline (CUDA events, L2 flushed before every timed launch, median of 10); rays = traversals "
                   "started (RT_CNT_RAYS_TRACED); parity = bitwise idx/t/u/v against the CPU oracle on the rays named; roofline bytes counted by the oracle"}
    med = lambda ts: float(np.median(ts))

    def passes(tag, arrays, bvh, w, h, light, d_radius=0.0, diffuse_spp=0, check_stride=1, resident=True, frame=True, note=None, cap=None):
        """primary / shadow / fused primary+shadow (/ diffuse / shaded frame) on the scene the context holds"""
        r = {}
        if note:
            r["note"] = note
        params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"], light_pos=light, d_radius=d_radius)
        ctx.set_params(params)
        n = w * h
        d_hits = torch.zeros((n, 4), device="cuda")
        d_rays = torch.zeros((n, 8), device="cuda")
        d_sh = torch.zeros((n, 4), device="cuda")
        d_sr = torch.zeros((n, 8), device="cuda")
        ctx.primary_device(w, h, d_hits, d_rays)
        ctx.shadow_device(n, d_rays, d_hits, d_sh, d_sr)
        torch.cuda.synchronize()
        hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
        sh = d_sh.cpu().numpy().view(HIT).reshape(-1)
        srays = d_sr.cpu().numpy()
        sc = O.OracleScene(arrays, bvh.nodes, bvh.tri_indices)
        orays, gate = O.primary_rays(params, w, h)
        g = gate.astype(bool)
        ntrav, nhit = int(g.sum()), int((hits["idx"] >= 0).sum())
        info = ctx.scene_info()
        r["scene"] = {"triangles": int(arrays["indices"].size // 3), "blob_MB": info["blob_bytes"] / 1e6, "frame": [w, h], "light": list(light)}
        # ---- primary ----
        sel = np.flatnonzero(g)[::check_stride]
        t0 = time.perf_counter()
        want, cnt = sc.trace(0, np.ascontiguousarray(orays[sel]))
        cpu_s = time.perf_counter() - t0
        ok = same_hits(hits[sel], want) and bool(np.all(hits["idx"][~g] == -1))
        ms = med(timer(lambda: ctx.primary_device(w, h, d_hits)))
        scale = ntrav / max(1, sel.size)
        cnt_full = {"inner": cnt["inner"] * scale, "tris": cnt["tris"] * scale}
        r["primary"] = {"rays": ntrav, "ms": ms, "mrays_s": ntrav / ms / 1e3, "cpu_oracle_mrays_s": sel.size / cpu_s / 1e6,
                        "parity": {"rays_checked": int(sel.size), "bit_identical": ok},
                        "roofline": roofline_of(cnt_full, 16 * n, ms, "trace_kernel<SRC_PRIMARY,closest>", resident, cap and f"traffic_{cap}_primary")}
        # ---- shadow (any-hit, buffer-fed: rays + closest hits in, records out) ----
        hsel = np.flatnonzero(hits["idx"] >= 0)[::check_stride]
        if hsel.size:
            t0 = time.perf_counter()
            want_s, cnt_s = sc.trace(1, np.ascontiguousarray(srays[hsel]))
            cpu_s = time.perf_counter() - t0
            ok_s = same_hits(sh[hsel], want_s)
            ms = med(timer(lambda: ctx.shadow_device(n, d_rays, d_hits, d_sh)))
            scale = nhit / hsel.size
            cnt_sf = {"inner": cnt_s["inner"] * scale, "tris": cnt_s["tris"] * scale}
            r["shadow_anyhit"] = {"rays": nhit, "ms": ms, "mrays_s": nhit / ms / 1e3, "occluded_fraction": float((want_s["idx"] >= 0).mean()),
                                  "cpu_oracle_mrays_s": hsel.size / cpu_s / 1e6, "parity": {"rays_checked": int(hsel.size), "bit_identical": ok_s},
                                  "roofline": roofline_of(cnt_sf, 64 * n, ms, "trace_kernel<SRC_SHADOW,any>", resident, cap and f"traffic_{cap}_shadow")}
            # ---- fused: primary + shadow in one launch ----
            d_vis = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            ctx.primary_shadow_device(w, h, None, None, d_vis)
            torch.cuda.synchronize()
            occl = (sh["idx"] >= 0) & (sh["t"] > np.float32(0.025))
            want_vis = np.where(hits["idx"] >= 0, hits["idx"] + occl, -1).astype(np.int32)
            ok_f = bool(np.array_equal(d_vis.cpu().numpy().reshape(-1), want_vis))
            ctx.reset_counters()
            ms = med(timer(lambda: ctx.primary_shadow_device(w, h, None, None, d_vis)))
            rays_f = ctx.counters()["rays_traced"] // 13
            cnt_f = {"inner": cnt_full["inner"] + cnt_sf["inner"], "tris": cnt_full["tris"] + cnt_sf["tris"]}
            r["primary_plus_shadow_fused"] = {"rays": rays_f, "ms": ms, "mrays_s": rays_f / ms / 1e3,
                                              "parity": {"pixels_checked": n, "vis_frame_identical_to_two_pass": ok_f},
                                              "roofline": roofline_of(cnt_f, 4 * n, ms, "primary_shadow_kernel", resident, cap and f"traffic_{cap}_fused")}
            ok = ok and ok_s and ok_f
        # ---- incoherent diffuse rays ----
        if diffuse_spp and nhit:
            d_dr = torch.zeros((nhit * diffuse_spp, 8), device="cuda")
            d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.diffuse_rays_device(n, d_rays, d_hits, diffuse_spp, 0x5EED, d_dr, d_cnt)
            torch.cuda.synchronize()
            nd = int(d_cnt.item())
            d_dh = torch.zeros((nd, 4), device="cuda")
            ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)
            torch.cuda.synchronize()
            dh = d_dh.cpu().numpy().view(HIT).reshape(-1)
            drays = d_dr.cpu().numpy()
            dsel = np.arange(0, nd, check_stride)
            t0 = time.perf_counter()
            want_d, cnt_d = sc.trace(0, np.ascontiguousarray(drays[dsel]))
            cpu_s = time.perf_counter() - t0
            ok_d = same_hits(dh[dsel], want_d)
            ms = med(timer(lambda: ctx.trace_device(rtb200.CLOSEST, nd, d_dr, d_dh)))
            scale = nd / dsel.size
            r[f"diffuse_{diffuse_spp}spp"] = {"rays": nd, "ms": ms, "mrays_s": nd / ms / 1e3, "hit_fraction": float((want_d["idx"] >= 0).mean()),
                                               "cpu_oracle_mrays_s": dsel.size / cpu_s / 1e6,
                                               "parity": {"rays_checked": int(dsel.size), "bit_identical": ok_d},
                                               "inner_visits_per_ray": cnt_d["inner"] / dsel.size, "tri_tests_per_ray": cnt_d["tris"] / dsel.size,
                                               "roofline": roofline_of({"inner": cnt_d["inner"] * scale, "tris": cnt_d["tris"] * scale}, 48 * nd, ms,
                                                                       "trace_lanes_kernel<SRC_BUFFER,closest>", resident, cap and f"traffic_{cap}_diffuse")}
            ok = ok and ok_d
            del d_dr, d_dh
        if frame:
            d_img = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            ms = med(timer(lambda: ctx.render_frame_device(w, h, d_img), iters=5))
            r["shaded_frame"] = {"ms": ms, "fps": 1e3 / ms}
            if n <= 640 * 480:
                ref_img, _ = sc.render_frame(params, w, h)
                img = d_img.cpu().numpy().view(np.uint32)
                sh8 = np.array([0, 8, 16])
                d = np.abs(((img[..., None] >> sh8) & 255).astype(np.int32) - ((ref_img[..., None] >> sh8) & 255).astype(np.int32))
                r["shaded_frame"]["parity"] = {"pixels_checked": n, "max_channel_diff_lsb": int(d.max()), "pixels_differing": int((d.max(-1) > 0).sum())}
                ok = ok and int(d.max()) <= 1
        if not ok:
            raise SystemExit(f"bench: extra block '{tag}': the GPU results differ from the oracle: {json.dumps(r)[:2000]}")
        out[tag] = r
        del d_hits, d_rays, d_sh, d_sr
        torch.cuda.empty_cache()

    # configs[2]: the bench scene is still uploaded
    passes("c3_terrain_1M_default_light", arrays, bvh, w, h, (-23.0, 200.0, 3.0), diffuse_spp=4, frame=True, cap="c2")
    passes("c3_terrain_1M_grazing_light", arrays, bvh, w, h, GRAZING_LIGHT, frame=False)
    # frames through the host entry points (pinned frame buffers, wall clock over 50 frames, params set every frame)
    params, _ = rtb200.camera_params(w, h, arrays["aabb_min"], arrays["aabb_max"])
    ctx.set_params(params)
    bufs = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
    nf = 50
    for _ in range(3):
        ctx.render_frame(w, h, bufs[0])
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(params)
        ctx.render_frame(w, h, bufs[0])
    fe = {"frame_by_frame_ms": (time.perf_counter() - t0) / nf * 1e3}
    for k in range(4):  # first use of a frame slot allocates its row-assembly scratch: not part of a steady frame rate
        ctx.render_frame_begin(w, h, bufs[k & 1], k & 1)
        ctx.render_frame_end(k & 1)
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(params)
        ctx.render_frame_begin(w, h, bufs[k & 1], k & 1)
        if k:
            ctx.render_frame_end((k - 1) & 1)
    ctx.render_frame_end((nf - 1) & 1)
    fe["two_in_flight_ms"] = (time.perf_counter() - t0) / nf * 1e3
    vis = torch.zeros((h, w), dtype=torch.int32).pin_memory()
    for _ in range(3):
        ctx.primary_shadow(w, h, vis)
    t0 = time.perf_counter()
    for k in range(nf):
        ctx.set_params(params)
        ctx.primary_shadow(w, h, vis)
    fe["primary_shadow_vis_frame_ms"] = (time.perf_counter() - t0) / nf * 1e3
    fe["call"] = "rt_set_params + rt_render_frame / rt_render_frame_begin+_end / rt_primary_shadow, 1080p, pinned host frames, wall clock per frame"
    out["host_entry_points_1080p"] = fe

    # configs[0]: 81 920-triangle icosphere through a COLLADA file (ColladaLoader -> Mesh::init), 640x480
    with tempfile.TemporaryDirectory() as td:
        src = rtb200.Mesh().icosphere(6, 50.0).finish(diffuse=(0.8, 0.3, 0.2))
        path = os.path.join(td, "c1.dae")
        src.write_dae(path)
        t0 = time.time()
        m1 = rtb200.Mesh().load_dae(path)
        load_s = time.time() - t0
    a1 = m1.arrays()
    b1 = rtb200.FlatBVH.build(m1)
    ctx.upload_scene(a1, b1.nodes, b1.tri_indices)
    passes("c1_sphere_81920_dae_640x480", a1, b1, 640, 480, (-23.0, 200.0, 3.0), diffuse_spp=4,
           cap="c1", note=f"COLLADA file parsed + baked in {load_s:.2f} s; 75 K traversed rays = half a wave of 8x4 tiles: launch/latency bound "
                          "(ncu: SMs active 47 % of the launch, issue active 32 %)")

    # configs[3]: incoherent diffuse bounce rays, 4 spp, 10 M-triangle instanced-sphere scene (1.14 GB blob: HBM-resident)
    if not args.skip_c4:
        t0 = time.time()
        field = rtb200.Mesh().sphere_field(11, 120.0, 6, 50.0).finish(diffuse=(0.6, 0.6, 0.8))
        a4 = field.arrays()
        b4 = rtb200.FlatBVH.build(field)
        build_s = time.time() - t0
        ctx.upload_scene(a4, b4.nodes, b4.tri_indices)
        passes("c4_sphere_field_10M_1080p", a4, b4, 1920, 1080, (-23.0, 200.0, 3.0), d_radius=1320.0, diffuse_spp=4, check_stride=4,
               resident=False, cap="c4",
               note=f"scene generated + SBVH built in {build_s:.1f} s on {os.cpu_count()} host threads; parity on every 4th ray. ncu: DRAM traffic is 0.15-0.25 x "
                    "the algorithmic bytes (L1 + L2 serve the rest), DRAM throughput 7-12 % of peak: bound by the latency of L2 misses "
                    "(long-scoreboard stalls, L2 hit 39-61 %) and by SIMT divergence (6-15 of 32 lanes active), not by HBM bandwidth")
        del field, a4, b4
    # leave the context on the bench scene
    ctx.upload_scene(arrays, bvh.nodes, bvh.tri_indices)
    ctx.set_params(params)
    return out


if __name__ == "__main__":
    main()
