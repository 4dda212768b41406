"""Multi-GPU plumbing (one process per GPU, torch.distributed): screen-space partition, the one-off
scene broadcast, and the per-frame framebuffer gather -- the only collectives of the path.

Partition = interleaved bands of `band_rows` rows (a multiple of the kernel's 4-row warp tile): band b
belongs to rank b % world, which balances sky against terrain. Every rank holds the whole packed scene.
The reference has no multi-device code; this is new.

Everything here is backend agnostic: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import ctypes
import glob
import os
from multiprocessing import shared_memory

import numpy as np
import torch


def owned_rows(rank, world, h, band_rows):
    """Row indices (ascending) that `rank` renders."""
    if band_rows < 4 or band_rows % 4:
        raise ValueError("band_rows must be a positive multiple of 4")
    ys = np.arange(h)
    return ys[(ys // band_rows) % world == rank]


def rows_per_rank(world, h, band_rows):
    """Rows of the largest share: every rank's band buffer is padded to this so the gather is uniform."""
    return max(len(owned_rows(r, world, h, band_rows)) for r in range(world))


def band_index(rank, world, h, band_rows, device=None):
    """Index tensor that picks this rank's rows out of a full (h, w) frame, padded by repeating the last row."""
    rows = owned_rows(rank, world, h, band_rows)
    n = rows_per_rank(world, h, band_rows)
    if len(rows) == 0:
        rows = np.zeros(1, dtype=np.int64)
    padded = np.concatenate([rows, np.full(n - len(rows), rows[-1])])
    return torch.as_tensor(padded, dtype=torch.int64, device=device)


def gather_bands(dist, band, out=None):
    """all_gather of the per-rank band buffers (rows_per_rank, w) -> (world, rows_per_rank, w)."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world,) + tuple(band.shape), dtype=band.dtype, device=band.device)
    # a list of views into one contiguous buffer: one collective on NCCL, and accepted by gloo as well
    dist.all_gather(list(out.unbind(0)), band.contiguous())
    return out


def assemble_frame(gathered, h, band_rows):
    """(world, rows_per_rank, w) gathered bands -> (h, w) frame (drops the padding rows)."""
    world, _n, w = gathered.shape
    frame = torch.empty((h, w), dtype=gathered.dtype, device=gathered.device)
    for r in range(world):
        rows = owned_rows(r, world, h, band_rows)
        if len(rows):
            frame[torch.as_tensor(rows, device=gathered.device)] = gathered[r, : len(rows)]
    return frame


def broadcast_scene(dist, ctx, rank, device):
    """Rank 0 has uploaded the scene: broadcast its packed blob as ONE buffer and adopt it everywhere.
    Returns the tensor that owns the blob on this rank (keep it alive as long as the context uses it)."""
    meta = [ctx.scene_blob()[1] if rank == 0 else 0]
    dist.broadcast_object_list(meta, src=0)
    blob = torch.empty(meta[0], dtype=torch.uint8, device=device)
    if rank == 0:
        ctx.copy_scene_blob(blob, blob.numel())
    dist.broadcast(blob, src=0)
    if blob.is_cuda:
        # the collective runs on the framework's stream; the context reads the blob on its own stream
        torch.cuda.synchronize(blob.device)
    ctx.adopt_scene_blob(blob.data_ptr(), blob.numel())
    return blob


def open_shared_frame(dist, ctx, rank, nbytes):
    """Store-fused gather: rank 0 allocates the framebuffer, every other rank maps it through CUDA IPC and its
    trace kernel writes pixels straight into it over NVLink (rt_primary_gather_device). Returns the device
    pointer valid on THIS rank; release with close_shared_frame."""
    box = [None]
    ptr = None
    if rank == 0:
        try:
            ptr, handle = ctx.ipc_alloc(nbytes)
            box = [handle]
        except Exception as exc:  # noqa: BLE001  -- tell the peers instead of leaving them in the broadcast
            box = [exc]
    dist.broadcast_object_list(box, src=0)
    if isinstance(box[0], Exception):
        raise RuntimeError(f"rank 0 could not create the shared framebuffer: {box[0]}")
    if rank != 0:
        ptr = ctx.ipc_open(box[0])
    return ptr


def close_shared_frame(dist, ctx, rank, ptr):
    dist.barrier()
    if rank == 0:
        ctx.ipc_free(ptr)
    else:
        ctx.ipc_close(ptr)


# ---- NUMA placement of the shared host frame ---------------------------------------------------------------------------
# Eight GPUs storing into ONE host buffer measured 106 GB/s in total (13 GB/s per GPU against ~30 GB/s alone): all pages sit
# on the NUMA node of the process that touched them first, and half of the GPUs reach it through the other socket. The pages
# of a band are therefore bound to the node the band's GPU hangs off (mbind on the tmpfs object, before anybody faults the
# pages in), or, with RTB_NUMA=interleave, spread round-robin over the nodes.
_MPOL_BIND, _MPOL_INTERLEAVE, _SYS_MBIND = 2, 3, 237  # x86-64 Linux


def numa_nodes():
    return sorted(int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` (from its PCI address), or -1"""
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
        if bus is None:
            import subprocess

            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception:  # noqa: BLE001
        return -1


def _mbind(addr, nbytes, mode, nodes):
    libc = ctypes.CDLL(None, use_errno=True)
    mask = (ctypes.c_ulong * 16)()
    for n in nodes:
        mask[n // 64] |= 1 << (n % 64)
    rc = libc.syscall(_SYS_MBIND, ctypes.c_void_p(addr), ctypes.c_ulong(nbytes), ctypes.c_int(mode), mask, ctypes.c_ulong(16 * 64 + 1), ctypes.c_uint(0))
    return rc == 0


def place_frame_pages(addr, frame_bytes, n_frames, frame_stride, h, w, itemsize, band_rows, rank_nodes, policy):
    """Apply `policy` ("bands" | "interleave" | "off") to the frames that start at `addr`; returns what was done."""
    nodes = numa_nodes()
    if policy == "off" or len(nodes) < 2:
        return f"off ({len(nodes)} NUMA node(s))"
    page = 4096
    if policy == "interleave" or any(n < 0 for n in rank_nodes):
        ok = _mbind(addr, frame_stride * n_frames, _MPOL_INTERLEAVE, nodes)
        return f"interleave over nodes {nodes}: {'ok' if ok else 'mbind refused'}"
    world, row_bytes, done = len(rank_nodes), w * itemsize, 0
    for f in range(n_frames):
        base = addr + f * frame_stride
        for b in range((h + band_rows - 1) // band_rows):
            lo = base + b * band_rows * row_bytes
            hi = min(base + (b + 1) * band_rows * row_bytes, base + frame_bytes)
            lo_p, hi_p = (lo + page - 1) // page * page, hi // page * page  # whole pages inside the band
            if hi_p > lo_p and _mbind(lo_p, hi_p - lo_p, _MPOL_BIND, [rank_nodes[b % world]]):
                done += 1
    return f"bands bound to their GPU's node {rank_nodes}: {done} band ranges"


class HostFrame:
    """Framebuffer(s) in POSIX shared memory that every rank page-locks and maps into its GPU: each rank's trace kernel
    stores its bands straight into the consumer's HOST memory over its own PCIe link (rt_host_register). A frame is
    complete on rank 0 -- without any device->host copy -- once every rank's kernel has finished; the ranks say so through
    one sequence word each (rt_signal, stream-ordered behind the kernel) that rank 0 polls, instead of a barrier per
    frame, and rank 0 publishes the last frame it has seen complete in an `ack` word so that nobody overwrites a frame
    the consumer has not taken yet (`n_frames` buffers are cycled)."""

    CTRL_BYTES = 4096  # one 64-byte line per rank for its sequence word, the ack word in the last line

    def __init__(self, dist, ctx, rank, h, w, dtype=np.int32, n_frames=1, band_rows=16, numa=None):
        self.dist, self.ctx, self.rank, self.world = dist, ctx, rank, dist.get_world_size()
        numa = numa or os.environ.get("RTB_NUMA", "bands")
        my_node = [gpu_numa_node(torch.cuda.current_device()) if torch.cuda.is_available() else -1]
        all_nodes = [None] * self.world
        dist.all_gather_object(all_nodes, my_node[0])
        self.numa = "off"
        self.frame_bytes = int(h) * int(w) * np.dtype(dtype).itemsize
        self.frame_stride = (self.frame_bytes + 4095) & ~4095
        self.n_frames = n_frames
        nbytes = self.frame_stride * n_frames + self.CTRL_BYTES
        box = [None]
        if rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
                try:  # before anybody touches (page-locks) the frames
                    addr = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
                    self.numa = place_frame_pages(addr, self.frame_bytes, n_frames, self.frame_stride, int(h), int(w),
                                                  np.dtype(dtype).itemsize, band_rows, all_nodes, numa)
                except Exception as exc:  # noqa: BLE001
                    self.numa = f"off ({exc})"
                box = [self.shm.name]
            except Exception as exc:  # noqa: BLE001  (e.g. /dev/shm too small) -- tell the peers
                box = [exc]
        dist.broadcast_object_list(box, src=0)
        if isinstance(box[0], Exception):
            raise RuntimeError(f"rank 0 could not create the shared-memory framebuffer: {box[0]}")
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:  # only the creator may unlink; keep Python's resource tracker from doing it when a peer exits
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.frames = [np.ndarray((h, w), dtype=dtype, buffer=self.shm.buf, offset=k * self.frame_stride) for k in range(n_frames)]
        self.array = self.frames[0]
        ctrl = self.frame_stride * n_frames
        self.words = np.ndarray((self.CTRL_BYTES // 4,), dtype=np.uint32, buffer=self.shm.buf, offset=ctrl)
        if rank == 0:
            self.words[:] = 0
        self.host_ptr = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
        self.nbytes = nbytes
        self.device_alias = ctx.host_register(self.host_ptr, nbytes)
        self._ctrl_alias = self.device_alias + ctrl
        dist.barrier()

    def frame_alias(self, k):
        return self.device_alias + (k % self.n_frames) * self.frame_stride

    def signal(self, seq):
        """enqueue on the context stream: this rank's sequence word <- seq once its part of the frame has landed"""
        self.ctx.signal(self._ctrl_alias + 64 * self.rank, seq)

    @staticmethod
    def _spin(words, index, seq, what, timeout_s=60.0):
        import time

        spins, t0 = 0, None
        while words[index] < seq:
            spins += 1
            if (spins & 0xFFFF) == 0:  # a dead peer must not hang the job
                t0 = t0 or time.time()
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"{what}: sequence word {index // 16} stayed at {int(words[index])}, waiting for {seq}")

    def wait_frame(self, seq):
        """rank 0: spin until every rank has signalled frame `seq`, then acknowledge it"""
        for r in range(self.world):
            self._spin(self.words, 16 * r, seq, "wait_frame")
        self.words[16 * (self.CTRL_BYTES // 64 - 1)] = seq

    def wait_ack(self, seq):
        """any rank: spin until rank 0 has taken frame `seq`"""
        self._spin(self.words, 16 * (self.CTRL_BYTES // 64 - 1), seq, "wait_ack")

    def close(self):
        self.ctx.synchronize()
        self.ctx.host_unregister(self.host_ptr)
        self.dist.barrier()
        self.array = self.frames = self.words = None
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            self.shm.unlink()
