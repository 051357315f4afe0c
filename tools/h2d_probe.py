#!/usr/bin/env python
"""H2D bandwidth of one step's frame features (103 MB fp32) from pinned host memory: default pinned, write-combined
pinned, and the default split over two copy streams.   python tools/h2d_probe.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from svol_b200 import comm
torch.cuda.set_device(0)
print("numa bound:", comm.bind_to_gpu_numa_node(0))
n = 32 * 1568 * 512
dev = torch.empty(n, dtype=torch.float32, device="cuda")
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n * 4), ctypes.c_uint(flags))
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((ctypes.c_float * n).from_address(p.value))
    arr[:] = 1.0
    return torch.from_numpy(arr)
bufs = {"default pinned (torch)": torch.ones(n).pin_memory(), "cudaHostAlloc default": host_alloc(0), "write-combined": host_alloc(4),
        "portable|mapped": host_alloc(1 | 2)}
def bw(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for name, h in bufs.items():
    print(f"{name:26s} pinned={h.is_pinned()}  {bw(lambda: dev.copy_(h, non_blocking=True)):6.1f} GB/s")
h = bufs["default pinned (torch)"]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
def two():
    with torch.cuda.stream(s1): dev[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): dev[half:].copy_(h[half:], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print(f"{'two streams':26s} {bw(two):6.1f} GB/s")
