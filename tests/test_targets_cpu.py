"""Host logic of the target marshalling (svol_b200/modeling/targets.py): the flat arrays follow the reference's
flatten order (video -> frame -> instance, matcher.py:62-70 / loss.py:78-85), and packing at collate time
 gives the same flat targets as walking the nested dicts in the training process."""
import numpy as np
import torch

from oracle import svol_oracle as orc
from svol_b200 import synth
from svol_b200.modeling import build_matcher
from svol_b200.modeling import targets as T


def _fields(flat):
    return {k: getattr(flat, k).numpy() for k in ("tgt_boxes", "tgt_off", "match_off", "cost_off", "video_tgt_off",
                                                  "video_match_off", "match_video")}


def test_flat_targets_follow_reference_order():
    cfg = synth.CONFIGS["C1b"]
    tn = synth.make_targets(cfg, 5, seed=4)
    targets = synth.targets_to_torch(tn)
    flat = T.flatten_targets(targets, torch.device("cpu"), True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
    boxes, per_video = orc.flatten_targets(tn)[:2]
    assert np.array_equal(flat.tgt_boxes.numpy(), np.asarray(boxes, np.float32).reshape(-1, 4))
    assert flat.B == 5 and flat.P == 5 * cfg.num_frames and flat.S == flat.tgt_boxes.shape[0]
    counts = np.concatenate([t["num_boxes_per_frame"] for t in tn])
    assert np.array_equal(np.diff(flat.tgt_off.numpy()), counts)
    assert np.array_equal(np.diff(flat.match_off.numpy()), np.minimum(counts, cfg.num_queries_per_frame))
    assert np.array_equal(np.diff(flat.video_tgt_off.numpy()), [t["total_boxes"] for t in tn])


def test_collate_time_packing_equals_in_process_flattening():
    for matcher_name in ("per_frame_matcher", "video_matcher"):
        cfg = synth.CONFIGS["C1b"]
        ns = cfg.to_namespace()
        ns.matcher = matcher_name
        matcher = build_matcher(ns)
        targets = synth.targets_to_torch(synth.make_targets(cfg, 6, seed=9))
        collate = T.make_collate_fn(lambda batch: ({"x": 1}, [d["targets"] for d in batch]), matcher, cfg.num_queries)
        inputs, packed = collate([{"targets": t} for t in targets])
        assert isinstance(packed, T.PackedTargets) and len(packed) == 6 and packed[2] is targets[2]      # still the plain list
        a = matcher._flat(targets, torch.device("cpu"), cfg.num_queries)
        b = matcher._flat(packed, torch.device("cpu"), cfg.num_queries)
        fa, fb = _fields(a), _fields(b)
        for k in fa:
            assert np.array_equal(fa[k], fb[k]), (matcher_name, k)
        assert (a.K, a.P, a.S, a.max_cols, a.cost_total, a.per_frame) == (b.K, b.P, b.S, b.max_cols, b.cost_total, b.per_frame)
        # a packing made for another configuration is ignored (falls back to the walk), never misused
        other = T.pack_targets(targets, True, cfg.num_frames, cfg.num_queries + 10, cfg.num_queries_per_frame + 1)
        if matcher_name == "per_frame_matcher":
            c = matcher._flat(other, torch.device("cpu"), cfg.num_queries)
            assert np.array_equal(_fields(c)["match_off"], fa["match_off"])


def test_static_views_match_the_flat_packing():
    """The criterion's static device workspace (modeling/loss.py) addresses a COPY of the packed target buffer through
    ``static_views``: every array in front of the boxes must sit where ``flatten_targets`` put it, for any batch with the
    same (P, B) -- and ``order`` lists the problems by number of targets, largest first (svol_match's launch order)."""
    cfg = synth.CONFIGS["C2"]
    for seed, mpf in ((0, 2), (1, 5), (2, 9)):
        tn = synth.make_targets(cfg, 3, seed, max_per_frame=mpf)
        targets = synth.targets_to_torch(tn)
        flat = T.flatten_targets(targets, torch.device("cpu"), True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
        cap = flat.n_fixed + 16 * (flat.S + 7)
        buf = torch.zeros(cap, dtype=torch.uint8)
        buf[:flat.n_static].copy_(flat.packed[:flat.n_static])
        v = T.static_views(buf, flat.P, flat.B)
        assert v["n_fixed"] == flat.n_fixed and flat.n_static == flat.n_fixed + 16 * flat.S
        for k in ("cost_off", "tgt_off", "match_off", "video_tgt_off", "video_match_off", "meta", "order"):
            assert torch.equal(v[k], getattr(flat, k)), k
        assert v["meta"].tolist() == [flat.K, flat.S, flat.max_cols, flat.P]
        assert torch.equal(v["tgt_boxes"][:16 * flat.S].view(torch.float32).view(-1, 4), flat.tgt_boxes)
        cols = (flat.tgt_off[1:] - flat.tgt_off[:-1]).numpy()
        order = flat.order.numpy()
        assert sorted(order.tolist()) == list(range(flat.P)) and (np.diff(cols[order]) <= 0).all()
        # same (P, B) -> same fixed layout, whatever the number of boxes
        assert v["n_fixed"] == T.static_views(torch.zeros(0, dtype=torch.uint8), flat.P, flat.B)["n_fixed"]
