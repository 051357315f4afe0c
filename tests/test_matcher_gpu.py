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
    _, _, status, cost_ws = run_match(torch.from_numpy(logits).to(DEV), torch.from_numpy(boxes).to(DEV), flat, 2.0, 5.0, 1.0,
                                      export_cost=True)
    assert status.cpu().tolist() == [0, 0]
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


def _lsap_gpu(costs, solver):
    """Runs a list of 2-D cost matrices through svol_lsap_f32 (the solver of svol_match on caller-supplied costs)."""
    from svol_b200 import _lib
    shapes = np.array([c.shape for c in costs], np.int32)
    sizes = np.array([c.size for c in costs], np.int64)
    cost_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    nout = shapes.min(axis=1).astype(np.int64)
    out_off = np.concatenate([[0], np.cumsum(nout)]).astype(np.int64)
    flat = np.concatenate([np.asarray(c, np.float32).ravel() for c in costs] + [np.zeros(1, np.float32)])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    d_cost, d_off, d_shape, d_out_off = t(flat), t(cost_off), t(shapes), t(out_off)
    rows = torch.full((max(int(out_off[-1]), 1),), -7, device=DEV, dtype=torch.int64)
    cols = torch.full_like(rows, -7)
    status = torch.full((len(costs),), -1, device=DEV, dtype=torch.int32)
    _lib.check(_lib.get_lib().svol_lsap_f32(d_cost.data_ptr(), d_off.data_ptr(), d_shape.data_ptr(), len(costs),
                                            int(shapes.min(axis=1).max()), int(shapes.max(axis=1).max()),
                                            int((shapes[:, 0] * shapes[:, 1]).max()), rows.data_ptr(), cols.data_ptr(),
                                            d_out_off.data_ptr(), status.data_ptr(), solver, _lib.stream_ptr()), "lsap")
    torch.cuda.synchronize()
    return rows.cpu().numpy(), cols.cpu().numpy(), status.cpu().numpy(), out_off


@pytest.mark.parametrize("solver", [0, 1])
def test_gpu_lsap_matches_scipy_edge_cases(golden_dir, solver):
    """scipy.optimize.linear_sum_assignment's own outputs (tests/golden/lsap_cases.npz, generated by calling scipy:
    heavy ties, constant matrices, +inf entries, tall / wide / one-column problems, 320 x 64, 64 x 320, 100 x 50)
    through the GPU solver -- both the register-state solver and the shared-memory-state one must reproduce the
    assignment bit for bit, ties included (matcher.py:93,158)."""
    g = _golden(golden_dir, "lsap_cases")
    shapes, cost = g["shapes"], g["cost"]
    costs, expect, off, ro = [], [], 0, 0
    for nr, nc in shapes:
        n = int(nr) * int(nc)
        k = int(min(nr, nc))
        if n > 0:
            c = cost[off:off + n].reshape(int(nr), int(nc))
            assert np.array_equal(c.astype(np.float32).astype(np.float64), c) or np.isinf(c).any(), "case not exact in fp32"
            costs.append(c)
            expect.append((g["rows"][ro:ro + k], g["cols"][ro:ro + k]))
        off += n
        ro += k
    rows, cols, status, out_off = _lsap_gpu(costs, solver)
    assert (status == 0).all(), np.nonzero(status)[0]
    bad = [i for i, (er, ec) in enumerate(expect)
           if not (np.array_equal(rows[out_off[i]:out_off[i + 1]], er) and np.array_equal(cols[out_off[i]:out_off[i + 1]], ec))]
    assert not bad, f"{len(bad)} of {len(expect)} scipy cases differ, first: {bad[:5]} shape {costs[bad[0]].shape}"


@pytest.mark.parametrize("solver", [0, 1])
def test_gpu_lsap_invalid_and_infeasible(solver):
    """scipy raises "matrix contains invalid numeric entries" for NaN / -inf and "cost matrix is infeasible" when a row
    has only +inf entries; the GPU solver reports them per problem (status 1 / 2) and solves its neighbours."""
    rng = np.random.RandomState(3)
    ok = rng.rand(6, 9).astype(np.float32)
    nan = ok.copy(); nan[2, 3] = np.nan
    ninf = ok.copy(); ninf[0, 0] = -np.inf
    infeasible = ok.copy(); infeasible[4, :] = np.inf
    _, _, status, _ = _lsap_gpu([ok, nan, ninf, infeasible, ok.T.copy()], solver)
    assert status.tolist() == [0, 1, 1, 2, 0]


@pytest.mark.parametrize("case,cfg", [("C5_b2", C["C5"]), ("C5_video", replace(C["C2"], matcher="video_matcher")),
                                      ("C2_b4", C["C2"])])
def test_match_solver_variants_and_split_modes_agree(golden_dir, case, cfg):
    """svol_match with the shared-memory-state solver, and the cost-only + solve-only pair of launches (the two halves
    that are profiled separately), give the indices of the default fused launch -- which the golden test pins to the
    reference."""
    from svol_b200.modeling.matcher import run_match
    g = _golden(golden_dir, "crit_" + case)
    batch, seed = int(g["batch"]), int(g["seed"])
    targets = synth.targets_to_torch(synth.make_targets(cfg, batch, seed, max_per_frame=int(g["max_per_frame"])))
    logits, boxes = synth.make_predictions(cfg, batch, seed)
    m = build_matcher(cfg.to_namespace())
    lg, bx = torch.from_numpy(logits).to(DEV), torch.from_numpy(boxes).to(DEV)
    flat = m._flat(targets, lg.device, lg.shape[2])
    w = (m.cost_class, m.cost_bbox, m.cost_giou)
    p0, t0, st0, _ = run_match(lg, bx, flat, *w)
    p1, t1, st1, _ = run_match(lg, bx, flat, *w, solver=1)
    _, _, _, ws = run_match(lg, bx, flat, *w, mode=1)
    p2, t2, st2, _ = run_match(lg, bx, flat, *w, mode=2, cost_ws=ws)
    assert st0.cpu().tolist() == [0, 0] and st1.cpu().tolist() == [0, 0] and st2.cpu().tolist() == [0, 0]
    assert torch.equal(p0, p1) and torch.equal(t0, t1) and torch.equal(p0, p2) and torch.equal(t0, t2)
    n = logits.shape[0]
    assert np.array_equal(p0[n - 1].cpu().numpy(), g["pred_idx_0"]) and np.array_equal(t0[n - 1].cpu().numpy(), g["tgt_idx_0"])


def test_large_cost_blocks_solved_from_the_workspace():
    """Problems whose cost block exceeds the shared-memory budget (video matcher, 320 queries x up to 320 boxes = 400 KB)
    are solved from the global workspace, tall and wide ones alike."""
    cfg = replace(C["C2"], matcher="video_matcher")
    targets_np = synth.make_targets(cfg, 3, 3, max_per_frame=20)           # n_v = 341, 326, 302: both orientations occur
    targets = synth.targets_to_torch(targets_np)
    logits, boxes = synth.make_predictions(cfg, 3, 3, layers=1)
    m = build_matcher(cfg.to_namespace())
    got = m({"pred_logits": torch.from_numpy(logits[0]).to(DEV), "pred_boxes": torch.from_numpy(boxes[0]).to(DEV)}, targets)
    ref = orc.video_matcher(logits[0], boxes[0], targets)
    sizes = [sum(len(f) for f in t["bboxes"].values()) for t in targets_np]
    assert max(sizes) * cfg.num_queries * 4 > 96 * 1024 and min(sizes) < cfg.num_queries < max(sizes), sizes
    for (gp, gt), (rp, rt) in zip(got, ref):
        assert np.array_equal(gp.numpy(), rp) and np.array_equal(gt.numpy(), rt)


def test_nan_prediction_leaves_valid_indices_and_a_status():
    """A NaN logit (a diverged step) must not become an out-of-bounds access in the criterion: the matcher writes safe
    indices for the unsolvable problems and raises its status; check_status() raises scipy's error; the next call on
    the same workspace is clean (the status word is published and cleared on the device)."""
    cfg = C["tiny"]
    targets = synth.targets_to_torch(synth.make_targets(cfg, 2, 0))
    logits, boxes = synth.make_predictions(cfg, 2, 0)
    crit = build_loss(cfg.to_namespace()).to(DEV)
    bad = logits.copy()
    bad[-1, :, :, 0] = np.nan
    with torch.no_grad():
        crit(_outputs(bad, boxes), targets)
        torch.cuda.synchronize()                      # no illegal address
        pred, tgt, flat = crit.last_indices
        assert int(pred.min()) >= 0 and int(pred.max()) < cfg.num_queries and int(tgt.min()) >= 0 and int(tgt.max()) < flat.S
        with pytest.raises(ValueError, match="invalid numeric entries"):
            crit.check_status()
        losses = crit(_outputs(logits, boxes), targets)
        crit.check_status()
    ref, _ = orc.set_criterion({"pred_logits": logits[-1], "pred_boxes": boxes[-1],
                                "aux_outputs": [{"pred_logits": logits[0], "pred_boxes": boxes[0]}]},
                               synth.make_targets(cfg, 2, 0), cfg, return_indices=True)
    for k, v in ref.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-5 * max(1.0, abs(float(v))), k


def test_criterion_static_workspace_follows_changing_targets(golden_dir):
    """The criterion's launches run on a static workspace (targets copied into a fixed device buffer, sizes read on the
    device): alternating batches with different numbers of boxes -- and one that outgrows the capacity -- must each give
    the reference's indices and losses."""
    cfg = C["C2"]
    crit = build_loss(cfg.to_namespace()).to(DEV)
    g = _golden(golden_dir, "crit_C2_b4")
    batch, seed = int(g["batch"]), int(g["seed"])
    logits, boxes = synth.make_predictions(cfg, batch, seed)
    out = _outputs(logits, boxes)
    t_a = synth.targets_to_torch(synth.make_targets(cfg, batch, seed, max_per_frame=2))
    t_b_np = synth.make_targets(cfg, batch, seed + 1, max_per_frame=1)
    t_c_np = synth.make_targets(cfg, batch, seed + 2, max_per_frame=9)          # 4.5 boxes per frame: outgrows the first workspace
    with torch.no_grad():
        for rep in range(2):
            for tg_np, tg in ((None, t_a), (t_b_np, synth.targets_to_torch(t_b_np)), (t_c_np, synth.targets_to_torch(t_c_np))):
                losses = {k: float(v) for k, v in crit(out, tg).items()}
                crit.check_status()
                if tg_np is None:
                    for k in [k[5:] for k in g.files if k.startswith("loss/")]:
                        assert abs(losses[k] - float(g["loss/" + k])) <= 2e-5 * max(1.0, abs(float(g["loss/" + k]))), k
                    idx = crit.indices(-1)
                    assert np.array_equal(np.concatenate([p.numpy() for p, _ in idx]), g["pred_idx_0"])
                    assert np.array_equal(np.concatenate([t.numpy() for _, t in idx]), g["tgt_idx_0"])
                else:
                    ref, ref_idx = orc.set_criterion({"pred_logits": logits[-1], "pred_boxes": boxes[-1],
                                                      "aux_outputs": [{"pred_logits": logits[0], "pred_boxes": boxes[0]}]},
                                                     tg_np, cfg, return_indices=True)
                    for k, v in ref.items():
                        assert abs(losses[k] - float(v)) <= 2e-5 * max(1.0, abs(float(v))), k
                    for (gp, gt), (rp, rt) in zip(crit.indices(-1), ref_idx[0]):
                        assert np.array_equal(gp.numpy(), rp) and np.array_equal(gt.numpy(), rt)


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
