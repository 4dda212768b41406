"""bench.py's contract on the CPU tier: the reference arm prints ONE JSON line with the keys the driver reads (same metric /
unit / workload sentence as our arm), and our arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == "Mrays/s primary (1080p, 1M-tri terrain, SBVH)" and d["value"] > 0
    for key in ("ms_per_step", "steps", "warmup", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the workload sentence is the one our arm prints (the driver compares the two)
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"]["workload"] == bench.workload_string(1920, 1080)
    assert d["config"]["rays_per_step"] == 752541  # pixels of the frame that pass the scene gate (oracle-counted)


def test_our_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return  # (the GPU tier runs the real thing)
    p = _run("--steps", "1", "--warmup", "0", "--no-extra", timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_numa_helpers_degrade_quietly():
    """tiling.place_frame_pages on whatever this host is: never raises; one node -> 'off'"""
    import ctypes
    import mmap

    sys.path.insert(0, ROOT)
    import rtb200
    from rtb200 import tiling

    nodes = tiling.numa_nodes()
    assert nodes == sorted(nodes) and all(isinstance(n, int) for n in nodes)
    m = mmap.mmap(-1, 1 << 20)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    for policy in ("off", "bands", "interleave"):
        msg = tiling.place_frame_pages(addr, 1 << 20, 1, 1 << 20, 64, 4096, 4, 16, [0, 0], policy)
        assert isinstance(msg, str) and msg
        if len(nodes) < 2 or policy == "off":
            assert msg.startswith("off")
    assert tiling.gpu_numa_node(0) >= -1
    assert rtb200.device.cull_rect(np.zeros(32, dtype=np.float32), 64, 64) == (0, 63, 0, 63)  # degenerate params: whole frame
