"""Data-parallel plumbing, mirror of lib/utils/comm.py:5-25 plus the pair sharding the reference leaves to its
launch scripts (train_quickdraw.sh:34-37: one process per GPU; svol_dataloader.py:70: videos_per_gpu = bs // gpus).

Sketch-video pairs are independent through forward, matching and per-pair loss terms, so the hot path needs NO
data-path collective: every rank processes its own shard.  The only collectives are the ones the reference has
(loss averaging for logging, comm.py:21-25; the gradient all-reduce of a training step, train.py:124) and the
max-over-ranks of device timings in bench.py.  Works on any
torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:                       # comm.py:5-10
    return dist.get_rank() if is_distributed() else 0


def get_world_size() -> int:                 # comm.py:13-18
    return dist.get_world_size() if is_distributed() else 1


def reduce_tensor(tensor: torch.Tensor, world_size: int = None) -> torch.Tensor:
    """clone -> all_reduce(SUM) -> / world_size (comm.py:21-25)."""
    world_size = world_size or get_world_size()
    rt = tensor.clone()
    if is_distributed():
        dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= world_size
    return rt


def shard_range(n_items: int, rank: int = None, world_size: int = None) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank's items: the first n % world ranks get one extra.  Every item is
    owned by exactly one rank, ranks keep global order (so concatenating rank results restores batch order)."""
    rank = get_rank() if rank is None else rank
    world_size = get_world_size() if world_size is None else world_size
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank's value (device timings are reported as the max over ranks)."""
    if not is_distributed():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_loss_dict(loss_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Averages every entry over ranks in ONE collective (train.py:240 reduces the summed loss per iteration;
    folding all entries into one flat buffer keeps it to a single small all-reduce)."""
    if not is_distributed() or not loss_dict:
        return dict(loss_dict)
    keys = sorted(loss_dict)
    flat = torch.stack([loss_dict[k].detach().float().reshape(()) for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= get_world_size()
    return {k: flat[i] for i, k in enumerate(keys)}


def allreduce_gradients(flat_grad: torch.Tensor, async_op: bool = False):
    """Sums the flat fp32 gradient buffer of the head (``TrainEngine.grad_flat``, 27.7 MB at two layers) over the
    ranks in ONE collective -- the data-parallel exchange of a training step (apex DDP in the reference, train.py:124;
    SURVEY.md section 8e).  Returns ``(grad_scale, work)``: the optimizer multiplies the summed gradient by
    ``grad_scale = 1 / world_size`` inside its fused update (the reference averages), ``work`` is the async handle
    (or None).  The per-rank loss means are over the rank's own pairs (loss.py:55,94,102), so sum / world_size is
    exactly the gradient of the mean of the rank losses."""
    if not is_distributed():
        return 1.0, None
    work = dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)
    return 1.0 / get_world_size(), (work if async_op else None)


def all_gather_indices(local: torch.Tensor) -> torch.Tensor:
    """Concatenates variable-length 1-D int64 tensors of all ranks in rank order (gathering matched indices of
    the rank shards back into batch order for evaluation)."""
    if not is_distributed():
        return local
    world = get_world_size()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)])


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pins the calling process to the CPU cores NVML reports as local to the GPU (its NUMA node / PCIe root), so that
    host buffers it allocates afterwards -- the pinned staging memory of the H2D input copies -- are first-touched on
    that node.  With 8 ranks uploading ~100 MB per step each, memory that all sits on one socket caps the whole box at
    that socket's DRAM bandwidth.  Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False

