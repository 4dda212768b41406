"""Calibration for the N-GPU end-to-end number: how fast can the GPUs of this box write into host memory AT ONCE?
torchrun --nproc-per-node N tools/d2h_ceiling.py: every rank copies a device buffer into (a) its own pinned buffer with
cudaMemcpyAsync and (b) its slice of ONE shared-memory buffer registered by all ranks, first alone (rank by rank), then all
ranks at the same time. Prints aggregate GB/s; the bench's host-frame e2e cannot exceed the concurrent figure."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src = torch.empty(MB << 20, dtype=torch.uint8, device="cuda").fill_(rank + 1)
dst = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
stream = torch.cuda.Stream()


def copy_ms(iters=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
    stream.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


copy_ms(3)
res = {"MB_per_rank": MB, "ranks": world}
alone = []
for r in range(world):
    dist.barrier()
    ms = copy_ms() if r == rank else 0.0
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    alone.append(MB * 1.048576 / float(t.item()))
res["one_rank_at_a_time_GBps"] = [round(v, 1) for v in alone]
dist.barrier()
ms = copy_ms()
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
res["all_ranks_at_once_aggregate_GBps"] = round(world * MB * 1.048576 / float(t.item()), 1)
res["all_ranks_at_once_ms"] = float(t.item())
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
