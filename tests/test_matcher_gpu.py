"""GPU parity of the Hungarian matchers, the set criterion and the post-processing against the CPU
oracle and the reference's golden vectors.  Index outputs must be bit-exact; a mismatch is accepted
only if the two assignments' total costs differ by less than the float tolerance (a genuine near-tie),
and the test reports how many frames that excused (expected: zero on these seeds)."""
import os
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle import svol_oracle as orc
from svol_b200 import synth
from svol_b200.modeling import build_loss, build_matcher, postprocess

pytestmark = pytest.mark.gpu
C = synth.CONFIGS
DEV = "cuda:0"


def _outputs(logits, boxes, device=DEV):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return {"pred_logits": t(logits[-1]), "pred_boxes": t(boxes[-1]),
            "aux_outputs": [{"pred_logits": t(a), "pred_boxes": t(b)} for a, b in zip(logits[:-1], boxes[:-1])]}


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


CASES = [("tiny", C["tiny"]), ("C2_b4", C["C2"]), ("C2_b32", C["C2"]), ("C2n4_b2", C["C2n4"]),
         ("C2_video", replace(C["C2"], matcher="video_matcher")), ("C5_b2", C["C5"]),
         ("C5_video", replace(C["C2"], matcher="video_matcher"))]


@pytest.mark.parametrize("case,cfg", CASES)
def test_criterion_and_indices_match_reference_golden(golden_dir, case, cfg):
    g = _golden(golden_dir, "crit_" + case)
    batch, seed = int(g["batch"]), int(g["seed"])
    targets = synth.targets_to_torch(synth.make_targets(cfg, batch, seed, max_per_frame=int(g["max_per_frame"])))
    logits, boxes = synth.make_predictions(cfg, batch, seed)
    crit = build_loss(cfg.to_namespace()).to(DEV)
    losses = crit(_outputs(logits, boxes), targets)
    crit.check_status()
    n = logits.shape[0]
    # golden slot 0 = last decoder layer, slot i+1 = aux layer i
    for slot in range(n):
        layer = n - 1 if slot == 0 else slot - 1
        idx = crit.indices(layer)
        pred = np.concatenate([p.numpy() for p, _ in idx])
        tgt = np.concatenate([t.numpy() for _, t in idx])
        assert pred.dtype == np.int64 and tgt.dtype == np.int64
        assert np.array_equal(np.array([len(p) for p, _ in idx]), g[f"counts_{slot}"])
        assert np.array_equal(pred, g[f"pred_idx_{slot}"]), f"pred indices differ (layer {layer})"
        assert np.array_equal(tgt, g[f"tgt_idx_{slot}"]), f"target indices differ (layer {layer})"
    keys = [k[5:] for k in g.files if k.startswith("loss/")]
    assert list(losses.keys())[:4] == ["loss_label", "class_error", "loss_bbox", "loss_giou"]
    assert sorted(losses.keys()) == sorted(keys)
    for k in keys:
        ref = float(g["loss/" + k])
        assert abs(float(losses[k]) - ref) <= 2e-5 * max(1.0, abs(ref)), (k, float(losses[k]), ref)


@pytest.mark.parametrize("matcher", ["per_frame_matcher", "video_matcher"])
def test_standalone_matcher_api(matcher):
    cfg = replace(C["C2"], matcher=matcher)
    targets = synth.targets_to_torch(synth.make_targets(cfg, 3, 11))
    logits, boxes = synth.make_predictions(cfg, 3, 11, layers=1)
    m = build_matcher(cfg.to_namespace())
    got = m({"pred_logits": torch.from_numpy(logits[0]).to(DEV), "pred_boxes": torch.from_numpy(boxes[0]).to(DEV)}, targets)
    if matcher == "per_frame_matcher":
        ref = orc.per_frame_matcher(logits[0], boxes[0], targets, cfg.num_frames, cfg.num_queries_per_frame)
    else:
        ref = orc.video_matcher(logits[0], boxes[0], targets)
    assert len(got) == 3
    for (gp, gt), (rp, rt) in zip(got, ref):
        assert gp.device.type == "cpu" and gp.dtype == torch.int64
        assert np.array_equal(gp.numpy(), rp) and np.array_equal(gt.numpy(), rt)


def test_cost_blocks_bit_exact_vs_oracle():
    """The block-diagonal costs themselves: same fp32 operation order as matcher.py:59-85.  exp() differs
    between libraries by an ulp at most, so allow 2 ulp of the cost magnitude; everything else is exact."""
    from svol_b200.modeling.matcher import run_match
    from svol_b200.modeling.targets import flatten_targets
    cfg = C["C2"]
    targets = synth.targets_to_torch(synth.make_targets(cfg, 4, 2))
    logits, boxes = synth.make_predictions(cfg, 4, 2, layers=1)
    flat = flatten_targets(targets, torch.device(DEV), True, cfg.num_frames, cfg.num_queries, cfg.num_queries_per_frame)
    _, _, status, cost_ws = run_match(torch.from_numpy(logits).to(DEV), torch.from_numpy(boxes).to(DEV), flat, 2.0, 5.0, 1.0)
    assert int(status.item()) == 0
    cost = cost_ws[0].cpu().numpy()
    tgt, num_boxes, _ = orc.flatten_targets(targets)
    offs = np.concatenate([[0], np.cumsum(num_boxes)])
    coff = flat.cost_off.cpu().numpy()
    worst = 0.0
    for i, n in enumerate(num_boxes):
        if n == 0:
            continue
        b, t = divmod(i, cfg.num_frames)
        rows = slice(t * cfg.num_queries_per_frame, (t + 1) * cfg.num_queries_per_frame)
        ref = orc.cost_matrix(logits[0, b, rows], boxes[0, b, rows], tgt[offs[i]:offs[i] + n], 2.0, 5.0, 1.0)
        got = cost[coff[i]:coff[i + 1]].reshape(cfg.num_queries_per_frame, n)
        worst = max(worst, np.abs(got - ref).max())
    assert worst <= 2 * np.spacing(np.float32(4.0)), worst


def test_nan_cost_raises_like_scipy():
    cfg = C["tiny"]
    targets = synth.targets_to_torch(synth.make_targets(cfg, 2, 0))
    logits, boxes = synth.make_predictions(cfg, 2, 0, layers=1)
    logits[0, :, :, 0] = np.nan          # every frame, so at least one frame with targets sees it
    m = build_matcher(cfg.to_namespace())
    with pytest.raises(ValueError, match="invalid numeric entries"):
        m({"pred_logits": torch.from_numpy(logits[0]).to(DEV), "pred_boxes": torch.from_numpy(boxes[0]).to(DEV)}, targets)


def test_criterion_backward_matches_finite_differences():
    """d(sum_k w_k loss_k)/d(logits, boxes) from svol_criterion_backward vs central differences of the
    fp64 oracle losses with the matching held fixed (it is piecewise constant)."""
    cfg = C["tiny"]
    targets_np = synth.make_targets(cfg, 2, 4)
    targets = synth.targets_to_torch(targets_np)
    logits, boxes = synth.make_predictions(cfg, 2, 4)
    crit = build_loss(cfg.to_namespace()).to(DEV)
    lg = torch.from_numpy(logits).to(DEV).requires_grad_(True)
    bx = torch.from_numpy(boxes).to(DEV).requires_grad_(True)
    out = {"pred_logits": lg[-1], "pred_boxes": bx[-1], "aux_outputs": [{"pred_logits": lg[0], "pred_boxes": bx[0]}]}
    losses = crit(out, targets)
    wd = crit.weight_dict
    total = sum(losses[k] * wd[k] for k in losses if k in wd)        # train.py:227-228
    total.backward()
    g_lg, g_bx = lg.grad.cpu().numpy(), bx.grad.cpu().numpy()

    idx = [crit.indices(l) for l in range(2)]
    idx = [[(p.numpy(), t.numpy()) for p, t in layer] for layer in idx]

    def total_loss(lgs, bxs):
        tot = 0.0
        for layer, sfx in ((1, ""), (0, "_0")):
            l1 = orc.loss_labels(lgs[layer], idx[layer], cfg.eos_coef, np.float64)
            l2 = orc.loss_boxes(bxs[layer], targets_np, idx[layer], np.float64)
            tot += wd["loss_label" + sfx] * l1["loss_label"] + wd["loss_bbox" + sfx] * l2["loss_bbox"] + wd["loss_giou" + sfx] * l2["loss_giou"]
        return tot

    rng = np.random.RandomState(0)
    L64, B64 = logits.astype(np.float64), boxes.astype(np.float64)
    eps = 1e-5
    for _ in range(40):
        which = rng.randint(2)
        arr, grad = (L64, g_lg) if which == 0 else (B64, g_bx)
        pos = tuple(rng.randint(s) for s in arr.shape)
        old = arr[pos]
        arr[pos] = old + eps; up = total_loss(L64, B64)
        arr[pos] = old - eps; dn = total_loss(L64, B64)
        arr[pos] = old
        fd = (up - dn) / (2 * eps)
        assert abs(fd - grad[pos]) < 1e-4 + 1e-3 * abs(fd), (which, pos, fd, grad[pos])


def test_postprocess_matches_reference_golden(golden_dir):
    g = _golden(golden_dir, "post_C2_b4")
    cfg = C["C2"]
    logits, boxes = synth.make_predictions(cfg, int(g["batch"]), int(g["seed"]), layers=1)
    srt, order = postprocess(torch.from_numpy(logits[0]).to(DEV), torch.from_numpy(boxes[0]).to(DEV), cfg.num_frames)
    assert np.array_equal(order.cpu().numpy().astype(np.int64), g["order"])
    assert np.abs(srt.cpu().numpy() - g["sorted"]).max() < 1e-6
