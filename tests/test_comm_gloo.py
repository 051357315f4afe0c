"""N > 1 host logic on CPU: world_size-2 gloo process group (the GPU box runs the same code over NCCL).
Pairs are sharded across ranks with no data-path collective; the oracle plays the role of each rank's device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svol_b200 import comm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import svol_oracle as orc
        from svol_b200 import synth
        out = {}
        assert comm.get_rank() == rank and comm.get_world_size() == world
        # 1. shard 5 pairs over 2 ranks; every rank matches its own shard (no collective on the data path)
        cfg = synth.CONFIGS["tiny"]
        B = 5
        begin, end = comm.shard_range(B)
        out["range"] = (begin, end)
        rng = np.random.RandomState(7)
        logits = rng.standard_normal((B, cfg.num_queries, 2)).astype(np.float32)
        boxes = rng.uniform(0.2, 0.8, (B, cfg.num_queries, 4)).astype(np.float32)
        targets = synth.make_targets(cfg, B, seed=3)
        idx = orc.per_frame_matcher(logits[begin:end], boxes[begin:end], targets[begin:end], cfg.num_frames, cfg.num_queries_per_frame)
        flat = torch.from_numpy(np.concatenate([np.asarray(p, np.int64) + (begin + i) * cfg.num_queries
                                                for i, (p, _) in enumerate(idx)]))
        out["gathered"] = comm.all_gather_indices(flat).tolist()
        # 2. the reference's logging reduction and bench.py's timing reduction
        out["mean"] = float(comm.reduce_tensor(torch.tensor([float(rank + 1)]))[0])
        out["max"] = comm.max_over_ranks(10.0 * (rank + 1))
        red = comm.reduce_loss_dict({"loss_bbox": torch.tensor(float(rank)), "loss_giou": torch.tensor(2.0 * rank)})
        out["loss"] = {k: float(v) for k, v in red.items()}
        # 3. training step: one all-reduce over the flat gradient buffer, averaged inside the optimizer update
        g = torch.full((1000,), float(rank + 1))
        scale, _ = comm.allreduce_gradients(g)
        out["grad"] = (float(g[0]), float(g[-1]), scale)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_every_item_once():
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [comm.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_matching_equals_single_process():
    from oracle import svol_oracle as orc
    from svol_b200 import synth
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["range"] == (0, 3) and res[1]["range"] == (3, 5)
    # single-process matching of the whole batch == rank-order concatenation of the shards
    cfg = synth.CONFIGS["tiny"]
    rng = np.random.RandomState(7)
    logits = rng.standard_normal((5, cfg.num_queries, 2)).astype(np.float32)
    boxes = rng.uniform(0.2, 0.8, (5, cfg.num_queries, 4)).astype(np.float32)
    idx = orc.per_frame_matcher(logits, boxes, synth.make_targets(cfg, 5, seed=3), cfg.num_frames, cfg.num_queries_per_frame)
    want = np.concatenate([np.asarray(p, np.int64) + i * cfg.num_queries for i, (p, _) in enumerate(idx)]).tolist()
    assert res[0]["gathered"] == want and res[1]["gathered"] == want
    for r in range(world):
        assert res[r]["mean"] == 1.5 and res[r]["max"] == 20.0
        assert res[r]["loss"] == {"loss_bbox": 0.5, "loss_giou": 1.0}
        assert res[r]["grad"] == (3.0, 3.0, 0.5)          # summed over the two ranks; the optimizer applies 1 / world
