"""Pins the CPU oracle to the reference: every oracle function is compared with golden
vectors that tests/golden/make_golden.py recorded from the reference's own modules
(lib/modeling/{svanet,cross_modal_transformer,matcher,loss}.py, torch 2.11.0 CPU) and from
scipy 1.18.1's linear_sum_assignment.  CPU only."""
import os
from dataclasses import replace

import numpy as np
import pytest

from oracle import lsap, svol_oracle as orc
from svol_b200 import synth

C = synth.CONFIGS


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


# ------------------------------------------------------------------------------ LSAP
@pytest.mark.parametrize("impl", ["c", "python"])
def test_lsap_matches_scipy_golden(golden_dir, impl):
    g = _load(golden_dir, "lsap_cases")
    solve = lsap.lsap_c if impl == "c" else lsap.lsap_python
    if impl == "c" and lsap._load() is None:
        pytest.skip("C oracle not built")
    coff = ooff = 0
    for (nr, nc) in g["shapes"]:
        if impl == "python" and nr * nc > 2000:
            coff += nr * nc; ooff += min(nr, nc)
            continue
        c = g["cost"][coff:coff + nr * nc].reshape(nr, nc)
        k = min(nr, nc)
        r, col = solve(c)
        assert np.array_equal(r, g["rows"][ooff:ooff + k]), (nr, nc)
        assert np.array_equal(col, g["cols"][ooff:ooff + k]), (nr, nc)
        coff += nr * nc
        ooff += k


def test_lsap_rejects_nan():
    c = np.ones((3, 3)); c[1, 1] = np.nan
    with pytest.raises(ValueError):
        lsap.linear_sum_assignment(c)
    with pytest.raises(ValueError):
        lsap.lsap_python(c)


# ------------------------------------------------------------------------------ head forward
HEAD_CASES = [("tiny", "tiny", False), ("tiny_pad", "tiny", True), ("C1a", "C1a", False),
              ("C1b", "C1b", False), ("C1b_pad", "C1b", True)]


@pytest.mark.parametrize("case,cfgname,padded", HEAD_CASES)
def test_head_forward_matches_reference(golden_dir, case, cfgname, padded):
    g = _load(golden_dir, "head_" + case)
    cfg = C[cfgname]
    batch, seed = int(g["batch"]), int(g["seed"])
    sd = synth.random_state_dict(cfg, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    # fp64 oracle vs fp64 reference: restatement error only.  The tolerance is set by the one
    # place the reference itself stays in fp32 even when run in fp64 -- the sine table
    # (position_encoding.py:57,63) -- where numpy and torch sin/pow differ by an fp32 ulp.
    out = orc.svanet_forward(sd, inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"],
                             inp["src_video_mask"], nheads=cfg.nheads, dtype=np.float64,
                             return_intermediates=True)
    n = cfg.num_layers
    # golden order: aux layers 0..n-2 then the last layer == natural layer order
    assert np.abs(out["all_logits"] - g["logits_f64"]).max() < 5e-6
    assert np.abs(out["all_boxes"] - g["boxes_f64"]).max() < 5e-6
    assert np.abs(out["hs"][:, :, ::7, ::5] - g["hs_sample_f64"]).max() < 5e-6
    assert len(out["aux_outputs"]) == n - 1
    # fp32 oracle vs fp32 reference: both carry fp32 rounding noise
    if cfgname != "C1b" or not padded:
        out32 = orc.svanet_forward(sd, inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"],
                                   inp["src_video_mask"], nheads=cfg.nheads, dtype=np.float32,
                                   return_intermediates=True)
        assert out32["all_logits"].dtype == np.float32
        assert np.abs(out32["all_logits"] - g["logits_f32"]).max() < 2e-5
        assert np.abs(out32["all_boxes"] - g["boxes_f32"]).max() < 2e-5


# ------------------------------------------------------------------------------ matcher / criterion
CRIT_CASES = [("tiny", C["tiny"], 2, None), ("C2_b4", C["C2"], 2, None), ("C2_b32", C["C2"], 2, None),
              ("C2n4_b2", C["C2n4"], 2, None),
              ("C2_video", replace(C["C2"], matcher="video_matcher"), 2, None),
              ("C5_b2", C["C5"], 50, None),
              ("C5_video", replace(C["C2"], matcher="video_matcher"), 10, None)]


def _check_crit(g, cfg, outputs, targets):
    losses, all_idx = orc.set_criterion(outputs, targets, cfg, return_indices=True)
    for li, idx in enumerate(all_idx):
        pred = np.concatenate([p for p, _ in idx])
        tgt = np.concatenate([t for _, t in idx])
        counts = np.array([len(p) for p, _ in idx])
        assert np.array_equal(counts, g[f"counts_{li}"])
        assert np.array_equal(pred, g[f"pred_idx_{li}"]), f"layer slot {li}"
        assert np.array_equal(tgt, g[f"tgt_idx_{li}"]), f"layer slot {li}"
    keys = [k[5:] for k in g.files if k.startswith("loss/")]
    assert sorted(keys) == sorted(losses.keys())
    for k in keys:
        ref = float(g["loss/" + k])
        assert abs(float(losses[k]) - ref) <= 2e-5 * max(1.0, abs(ref)), k
    wd = orc.weight_dict(cfg)
    assert sorted(wd.keys()) == list(g["weight_keys"])
    assert np.allclose([wd[k] for k in sorted(wd)], g["weight_vals"])


@pytest.mark.parametrize("case,cfg,mpf,_", CRIT_CASES)
def test_matcher_and_criterion_match_reference(golden_dir, case, cfg, mpf, _):
    g = _load(golden_dir, "crit_" + case)
    batch, seed = int(g["batch"]), int(g["seed"])
    targets = synth.make_targets(cfg, batch, seed, max_per_frame=int(g["max_per_frame"]))
    logits, boxes = synth.make_predictions(cfg, batch, seed)
    outputs = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
               "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
    _check_crit(g, cfg, outputs, targets)


@pytest.mark.parametrize("case,seed,padded", [("C1b_e2e", 0, False), ("C1b_pad_e2e", 1, True)])
def test_forward_then_criterion_matches_reference(golden_dir, case, seed, padded):
    """End-to-end chain on the reference's fp32 head outputs (taken from the golden file so the
    comparison isolates matcher+criterion)."""
    g = _load(golden_dir, "crit_" + case)
    h = _load(golden_dir, "head_C1b_pad" if padded else "head_C1b")
    cfg = C["C1b"]
    fm = synth.make_inputs(cfg, 2, seed, padded=padded)["frame_mask"] if padded else None
    targets = synth.make_targets(cfg, 2, seed, frame_mask=fm)
    lg, bx = h["logits_f32"], h["boxes_f32"]
    outputs = {"pred_logits": lg[-1], "pred_boxes": bx[-1],
               "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(lg[:-1], bx[:-1])]}
    _check_crit(g, cfg, outputs, targets)


def test_postprocess_matches_reference(golden_dir):
    g = _load(golden_dir, "post_C2_b4")
    cfg = C["C2"]
    logits, boxes = synth.make_predictions(cfg, int(g["batch"]), int(g["seed"]), layers=1)
    srt, order = orc.postprocess(logits[0], boxes[0], cfg.num_frames)
    assert np.array_equal(order, g["order"])
    assert np.abs(srt - g["sorted"]).max() < 1e-6


# ------------------------------------------------------------------------------ torch restatement (CPU baseline)
@pytest.mark.parametrize("case,cfgname,padded", [("tiny", "tiny", False), ("tiny_pad", "tiny", True), ("C1b_pad", "C1b", True)])
def test_torch_port_head_matches_reference(golden_dir, case, cfgname, padded):
    """oracle/torch_port.py (the reference's CPU path restated with the reference's own ATen calls; what
    bench.py times as the CPU baseline) against the reference's fp32 golden outputs."""
    import torch
    from oracle import torch_port as tp
    g = _load(golden_dir, "head_" + case)
    cfg = C[cfgname]
    batch, seed = int(g["batch"]), int(g["seed"])
    sd = tp.state_dict_to_torch(synth.random_state_dict(cfg, seed))
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    t = lambda k: torch.from_numpy(inp[k])
    out = tp.svanet_forward(sd, t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"), nheads=cfg.nheads)
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]]).numpy()
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]]).numpy()
    assert np.abs(logits - g["logits_f32"]).max() < 2e-5
    assert np.abs(boxes - g["boxes_f32"]).max() < 2e-5


@pytest.mark.parametrize("case,cfgname,spread", [("C2n4_b4", "C2n4", 0.0), ("C2_b4_spread", "C2", 1.0), ("C1b_spread", "C1b", 1.0),
                                                 ("C4_b1", "C4", 0.0)])
def test_torch_port_head_matches_reference_full_size(golden_dir, case, cfgname, spread):
    """Round-2 goldens (tests/golden/make_golden_r2.py: four layers, the spread box head, the long clip with a masked
    tail): the ATen restatement against the reference's fp64 outputs, and -- where the golden carries a matching record
    -- its PerFrameMatcher against the reference's indices on those outputs."""
    import torch
    from oracle import torch_port as tp
    g = _load(golden_dir, "head_" + case)
    cfg = C[cfgname]
    batch, seed = int(g["batch"]), int(g["seed"])
    sd = tp.state_dict_to_torch(synth.random_state_dict(cfg, seed, box_spread=spread))
    inp = synth.make_inputs(cfg, batch, seed, padded=bool(g["padded"]))
    tail = int(g["mask_tail_frames"])
    if tail:
        inp["src_video_mask"][-1, -tail * cfg.tokens_per_frame:] = 0
        inp["frame_mask"][-1, -tail:] = 0
    t = lambda k: torch.from_numpy(inp[k])
    out = tp.svanet_forward(sd, t("src_sketch"), t("src_sketch_mask"), t("src_video"), t("src_video_mask"), nheads=cfg.nheads)
    logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    assert np.abs(logits.numpy() - g["logits_f64"]).max() < 5e-5
    assert np.abs(boxes.numpy() - g["boxes_f64"]).max() < 5e-5
    if "frame_gap" in g.files:
        targets = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
        ref_l, ref_b = torch.from_numpy(g["logits_f32"]), torch.from_numpy(g["boxes_f32"])
        for li in range(cfg.num_layers):
            idx = tp.per_frame_matcher(ref_l[li], ref_b[li], targets, cfg.num_frames, cfg.num_queries_per_frame,
                                       cfg.set_cost_class, cfg.set_cost_bbox, cfg.set_cost_giou)
            assert np.array_equal(torch.cat([p for p, _ in idx]).numpy(), g[f"pred_idx_{li}"])
            assert np.array_equal(torch.cat([t for _, t in idx]).numpy(), g[f"tgt_idx_{li}"])
        gaps = g["frame_gap"]
        assert np.isfinite(gaps).sum() > 0 and (gaps[np.isfinite(gaps)] >= 0).all()


@pytest.mark.parametrize("case,cfg,mpf", [("tiny", C["tiny"], 2), ("C2_b4", C["C2"], 2),
                                          ("C2_video", replace(C["C2"], matcher="video_matcher"), 2)])
def test_torch_port_criterion_matches_reference(golden_dir, case, cfg, mpf):
    import torch
    from oracle import torch_port as tp
    g = _load(golden_dir, "crit_" + case)
    batch, seed = int(g["batch"]), int(g["seed"])
    targets = synth.make_targets(cfg, batch, seed, max_per_frame=int(g["max_per_frame"]))
    logits, boxes = synth.make_predictions(cfg, batch, seed)
    tt = torch.from_numpy
    outputs = {"pred_logits": tt(logits[-1]), "pred_boxes": tt(boxes[-1]),
               "aux_outputs": [{"pred_logits": tt(a), "pred_boxes": tt(b)} for a, b in zip(logits[:-1], boxes[:-1])]}
    losses, all_idx = tp.set_criterion(outputs, targets, cfg)
    for li, idx in enumerate(all_idx):
        assert np.array_equal(torch.cat([p for p, _ in idx]).numpy(), g[f"pred_idx_{li}"])
        assert np.array_equal(torch.cat([t for _, t in idx]).numpy(), g[f"tgt_idx_{li}"])
    for k in [k[5:] for k in g.files if k.startswith("loss/")]:
        ref = float(g["loss/" + k])
        assert abs(float(losses[k]) - ref) <= 2e-5 * max(1.0, abs(ref)), k


# ------------------------------------------------------------------------------------------ training step (autograd)
def test_oracle_head_gradients_match_reference_autograd(golden_dir):
    """oracle/torch_port.head_gradients (autograd over the ATen restatement) against the gradients recorded from the
    reference's own SVANet in train mode (tests/golden/make_golden_grads.py), config C1a, batch 2."""
    from dataclasses import replace
    import torch
    from oracle import torch_port as tp
    cfg = replace(synth.CONFIGS["C1a"], input_dropout=0.0)
    gold = np.load(os.path.join(golden_dir, "grads_C1a_b2.npz"))
    batch, seed = int(gold["batch"]), int(gold["seed"])
    sd = synth.random_state_dict(cfg, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=bool(gold["padded"]))
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    grads, _, _ = tp.head_gradients(tp.state_dict_to_torch(sd), inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"],
                                    inp["src_video_mask"], gl, gb, nheads=cfg.nheads)
    names = [k[len("head_f32/norm/"):] for k in gold.files if k.startswith("head_f32/norm/")]
    assert set(names) == set(grads), "same set of parameters receives a gradient"
    scale = max(float(gold["head_f64/norm/" + k]) for k in names)
    stride = int(gold["stride"])
    for k in names:
        norm = float(gold["head_f32/norm/" + k])
        got = grads[k].double()
        assert abs(float(got.norm()) - norm) <= 1e-3 * max(norm, 1e-4 * scale), k
        sample = gold["head_f32/sample/" + k].astype(np.float64)
        assert np.abs(got.reshape(-1)[::stride].numpy() - sample).max() <= 1e-3 * max(np.abs(sample).max(), 1e-5 * scale), k


# ------------------------------------------------------------------------------------------ evaluation metrics
def _eval_arrays(cfg, batch, seed):
    """Flat evaluation inputs from the synthetic predictions, through the oracle's post-processing."""
    from svol_b200.evaluate import flatten_eval_targets
    logits, boxes = synth.make_eval_predictions(cfg, batch, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=True)
    targets = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
    sorted_rows, _ = orc.postprocess(logits, boxes, cfg.num_frames)
    post = np.asarray(sorted_rows, np.float32).reshape(batch * cfg.num_frames, cfg.num_queries_per_frame, 5)
    gt, gt_off, frame_off, frame_index = flatten_eval_targets(targets, cfg.num_frames)
    return post, gt, gt_off, frame_off, frame_index


@pytest.mark.parametrize("name", ["C2_b4", "C2_b3"])
def test_eval_oracle_matches_reference_eval_svol(name, golden_dir):
    """oracle/eval_oracle.py against the metric dictionary the reference's own eval_svol produced."""
    import json
    from oracle import eval_oracle as ev
    gold = np.load(os.path.join(golden_dir, f"eval_{name}.npz"))
    cfg = synth.CONFIGS["C2"]
    post, gt, gt_off, frame_off, frame_index = _eval_arrays(cfg, int(gold["batch"]), int(gold["seed"]))
    got = ev.eval_svol(post[frame_index], gt, gt_off, frame_off)
    ref = json.loads(str(gold["metrics"]))
    assert got == ref
