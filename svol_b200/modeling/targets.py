"""Host-side marshalling of the reference's nested ``targets`` structure into flat device arrays.

The reference re-walks the nested dicts once per decoder layer in the matcher (matcher.py:62-70) and
again in the box loss (loss.py:78-85).  Here the walk happens once per batch; the result is cached on
the batch object's identity so the criterion, the matcher and repeated layers share it.
Flatten order = video -> frame (dict insertion order) -> instance, exactly as the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch


@dataclass
class FlatTargets:
    B: int
    S: int                       # total target boxes
    K: int                       # total matched pairs per layer
    P: int                       # assignment problems per layer
    problems_per_video: int
    rows_per_problem: int
    max_cols: int
    cost_total: int              # floats of cost workspace per layer
    tgt_boxes: torch.Tensor      # [S,4] f32 (device)
    tgt_off: torch.Tensor        # [P+1] i32
    match_off: torch.Tensor      # [P+1] i32
    cost_off: torch.Tensor       # [P+1] i64
    video_tgt_off: torch.Tensor  # [B+1] i32
    video_match_off: torch.Tensor  # [B+1] i32
    match_video: torch.Tensor    # [K] i32
    h_video_match_off: np.ndarray
    per_frame: bool
    meta: torch.Tensor = None    # [4] i32 on the device: K, S, max_cols, P
    order: torch.Tensor = None   # [P] i32: problems sorted by number of targets, descending (launch order of svol_match)
    packed: torch.Tensor = None  # the whole packed buffer (uint8, device); [0, n_static) = everything but match_video
    n_fixed: int = 0             # bytes in front of the boxes: a function of (P, B) only
    n_static: int = 0            # n_fixed + 16 * S


def _walk(targets: Sequence[dict]):
    """One pass over the nested structure; the ~B*T*n leaves are gathered with a single stack instead of one
    tensor -> numpy conversion per box (measured: 2.3 ms -> 0.6 ms per batch of 32 videos / ~1000 boxes)."""
    leaves, per_video, per_frame_counts = [], [], []
    for t in targets:
        per_frame_counts.extend(t["num_boxes_per_frame"])
        n0 = len(leaves)
        for frame in t["bboxes"].values():
            for inst in frame:
                leaves.append(inst["bbox"])
        per_video.append(len(leaves) - n0)
    if not leaves:
        raise ValueError("targets contain no boxes (the dataset guarantees >= 1 per video, svol_dataset.py:272)")
    if isinstance(leaves[0], torch.Tensor):
        boxes = torch.stack(leaves).detach().to(device="cpu", dtype=torch.float32).numpy()
    else:
        boxes = np.asarray(leaves, dtype=np.float32)
    return np.ascontiguousarray(boxes.reshape(-1, 4)), per_video, [int(n) for n in per_frame_counts]


class PackedTargets(list):
    """The reference's ``batched_targets`` list (svol_dataset.py:310-312) plus its flat packing, built where the list
    is built -- in the DataLoader worker's ``collate_fn`` -- so that the training / evaluation process
    only issues one H2D copy per batch and never walks the nested dicts.  Behaves as the plain list for every other
    consumer (``test.py`` reads ``targets[i]['bboxes']`` etc.)."""

    def __init__(self, targets, host: torch.Tensor, meta: dict):
        super().__init__(targets)
        self.host, self.meta = host, meta


def _pack_host(targets: Sequence[dict], per_frame: bool, num_frames: int, num_queries: int, q_per_frame: int):
    """Nested targets -> (packed uint8 numpy buffer, meta).  Layout: [cost_off i64 | tgt_off, match_off, video_tgt_off,
    video_match_off, (K, S, max_cols, P), order i32 | boxes f32 | match_video i32], sections 16-byte aligned.  Everything in
    front of the boxes has a size that depends on (P, B) only, so a consumer with a static device copy of this buffer
    (the criterion's captured launch sequence, loss.py) finds every array at a fixed address for every batch."""
    boxes, per_video, per_frame_counts = _walk(targets)
    B = len(targets)
    if per_frame:
        if len(per_frame_counts) != B * num_frames:
            raise ValueError("'num_boxes_per_frame' must hold exactly num_frames entries per video (svol_dataset.py:267-271)")
        if sum(per_frame_counts) != boxes.shape[0]:
            raise ValueError("'num_boxes_per_frame' does not add up to the number of boxes under 'bboxes'")
        cols = np.asarray(per_frame_counts, np.int64)
        rows, ppv = q_per_frame, num_frames
    else:
        cols = np.asarray(per_video, np.int64)
        rows, ppv = num_queries, 1
    P = cols.shape[0]
    tgt_off = np.zeros(P + 1, np.int64); tgt_off[1:] = np.cumsum(cols)
    matched = np.minimum(cols, rows)
    match_off = np.zeros(P + 1, np.int64); match_off[1:] = np.cumsum(matched)
    cost_off = np.zeros(P + 1, np.int64); cost_off[1:] = np.cumsum(cols * rows)
    video_tgt_off = np.zeros(B + 1, np.int64); video_tgt_off[1:] = np.cumsum(per_video)
    video_match_off = match_off[::ppv].copy()
    K = int(match_off[-1])
    match_video = np.repeat(np.arange(B, dtype=np.int32), np.diff(video_match_off))
    S = int(boxes.shape[0])
    max_cols = int(cols.max())
    order = np.argsort(-cols, kind="stable")                 # largest problems first (svol_match's launch order)
    ints = np.concatenate([tgt_off, match_off, video_tgt_off, video_match_off, [K, S, max_cols, P], order]).astype(np.int32)
    n_cost = ((P + 1) * 8 + 15) // 16 * 16                                           # sections stay 16-byte aligned
    n_fixed = n_cost + (ints.shape[0] * 4 + 15) // 16 * 16
    n_box = S * 16
    total = n_fixed + n_box + K * 4
    hb = np.zeros(total, np.uint8)
    hb[:(P + 1) * 8].view(np.int64)[:] = cost_off
    hb[n_cost:n_cost + ints.shape[0] * 4].view(np.int32)[:] = ints
    hb[n_fixed:n_fixed + n_box].view(np.float32)[:] = boxes.reshape(-1)
    hb[n_fixed + n_box:total].view(np.int32)[:] = match_video
    meta = dict(B=B, S=S, K=K, P=P, ppv=ppv, rows=rows, max_cols=max_cols, cost_total=int(cost_off[-1]),
                n_cost=n_cost, n_fixed=n_fixed, n_box=n_box, total=total, video_match_off=video_match_off, per_frame=per_frame,
                key=(per_frame, num_frames, num_queries, q_per_frame))
    return hb, meta


def pack_targets(targets: Sequence[dict], per_frame: bool, num_frames: int, num_queries: int, q_per_frame: int,
                 pin: bool = False) -> PackedTargets:
    """Collate-time packing.  ``pin=True`` page-locks the buffer (only in a process that owns a CUDA context; a
    DataLoader with ``pin_memory=True`` does not look inside list subclasses, so the main process pins on upload)."""
    hb, meta = _pack_host(targets, per_frame, num_frames, num_queries, q_per_frame)
    host = torch.from_numpy(hb)
    return PackedTargets(targets, host.pin_memory() if pin else host, meta)


def make_collate_fn(base_collate, matcher, num_queries: int):
    """Wraps the reference's ``collate_fn`` (svol_dataset.py:310-319): the returned ``batched_targets`` is a
    ``PackedTargets`` for the configuration of ``matcher`` (= ``build_matcher(args)``) and ``args.num_queries``."""
    if hasattr(matcher, "num_queries_per_frame"):          # PerFrameMatcher
        args = (True, matcher.num_frames, num_queries, matcher.num_queries_per_frame)
    else:                                                   # HungarianMatcher (one problem per video)
        args = (False, 0, num_queries, 0)

    def collate(batch):
        inputs, targets = base_collate(batch)
        return inputs, pack_targets(targets, *args)
    return collate


def flatten_targets(targets: Sequence[dict], device, per_frame: bool, num_frames: int, num_queries: int,
                    q_per_frame: int) -> FlatTargets:
    if isinstance(targets, PackedTargets) and targets.meta["key"] == (per_frame, num_frames, num_queries, q_per_frame):
        hb, meta = targets.host.numpy(), targets.meta              # packed at collate time: no walk here
    else:
        hb, meta = _pack_host(targets, per_frame, num_frames, num_queries, q_per_frame)
    B, S, K, P, total = meta["B"], meta["S"], meta["K"], meta["P"], meta["total"]
    n_cost, n_fixed, n_box = meta["n_cost"], meta["n_fixed"], meta["n_box"]
    # one packed host buffer -> one H2D copy; the pinned staging buffers are recycled (ring + event) because
    # cudaHostAlloc per batch costs more than the copy itself
    if isinstance(targets, PackedTargets) and targets.host.is_pinned():
        host, done = targets.host, None
    elif device.type == "cuda":
        host, done = _staging(total)
        host.numpy()[:total] = hb
    else:
        host, done = torch.from_numpy(hb), None
    dbuf = host[:total].to(device, non_blocking=True)
    if done is not None:
        done.record()
    d_cost = dbuf[:(P + 1) * 8].view(torch.int64)
    d_int = dbuf[n_cost:n_fixed].view(torch.int32)
    d_box = dbuf[n_fixed:n_fixed + n_box].view(torch.float32).view(S, 4)
    n1, n2, n3, n4 = P + 1, 2 * (P + 1), 2 * (P + 1) + B + 1, 2 * (P + 1) + 2 * (B + 1)
    return FlatTargets(
        B=B, S=S, K=K, P=P, problems_per_video=meta["ppv"], rows_per_problem=meta["rows"],
        max_cols=meta["max_cols"], cost_total=meta["cost_total"],
        tgt_boxes=d_box, tgt_off=d_int[:n1], match_off=d_int[n1:n2],
        cost_off=d_cost, video_tgt_off=d_int[n2:n3], video_match_off=d_int[n3:n4],
        match_video=dbuf[n_fixed + n_box:].view(torch.int32), h_video_match_off=meta["video_match_off"],
        per_frame=meta["per_frame"], meta=d_int[n4:n4 + 4], order=d_int[n4 + 4:n4 + 4 + P], packed=dbuf, n_fixed=n_fixed,
        n_static=n_fixed + n_box)


def static_views(buf: torch.Tensor, P: int, B: int):
    """The arrays of a packed target buffer that sit at addresses fixed by (P, B): a static device copy ``buf`` (uint8,
    at least FlatTargets.n_static bytes) keeps them valid for every batch copied into it."""
    n_cost = ((P + 1) * 8 + 15) // 16 * 16
    n_int = 2 * (P + 1) + 2 * (B + 1) + 4 + P
    n_fixed = n_cost + (n_int * 4 + 15) // 16 * 16
    ints = buf[n_cost:n_fixed].view(torch.int32)
    n1, n2, n3, n4 = P + 1, 2 * (P + 1), 2 * (P + 1) + B + 1, 2 * (P + 1) + 2 * (B + 1)
    return dict(cost_off=buf[:(P + 1) * 8].view(torch.int64), tgt_off=ints[:n1], match_off=ints[n1:n2],
                video_tgt_off=ints[n2:n3], video_match_off=ints[n3:n4], meta=ints[n4:n4 + 4], order=ints[n4 + 4:n4 + 4 + P],
                tgt_boxes=buf[n_fixed:], n_fixed=n_fixed)


_STAGING = {"bufs": [], "next": 0}


def _staging(nbytes: int):
    """A pinned uint8 buffer of at least ``nbytes`` whose previous H2D copy has completed, and the event to record
    after the next copy.  Ring of 4 buffers, grown on demand."""
    st = _STAGING
    if len(st["bufs"]) < 4:
        st["bufs"].append([torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory(), torch.cuda.Event()])
        slot = st["bufs"][-1]
    else:
        slot = st["bufs"][st["next"] % 4]
        st["next"] += 1
        slot[1].synchronize()
        if slot[0].numel() < nbytes:
            slot[0] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    return slot[0], slot[1]


class TargetsCache:
    """Remembers the flattening of the most recently used ``targets`` lists (keyed by object identity, length and
    device; the list object is kept alive so the id cannot be recycled) -- the criterion calls into it once per
    forward, the standalone matcher once per call; a loop that cycles through a few resident batches hits it."""

    CAPACITY = 4

    def __init__(self):
        self._entries = []          # [(key, targets, FlatTargets)], most recent last
        self._key = None            # key of the most recent entry; set to None to force a re-flatten

    def get(self, targets, device, per_frame, num_frames, num_queries, q_per_frame) -> FlatTargets:
        key = (id(targets), len(targets), str(device), per_frame, num_frames, num_queries, q_per_frame)
        if self._key is None:
            self._entries = []
        for i, (k, ref, val) in enumerate(self._entries):
            if k == key and ref is targets:
                if i != len(self._entries) - 1:
                    self._entries.append(self._entries.pop(i))
                self._key = key
                return val
        val = flatten_targets(targets, device, per_frame, num_frames, num_queries, q_per_frame)
        self._entries.append((key, targets, val))
        if len(self._entries) > self.CAPACITY:
            self._entries.pop(0)
        self._key = key
        return val
