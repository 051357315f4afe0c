"""Generates tests/golden/grads_<case>.npz by running the REFERENCE's own SVANet / SetCriterion in train mode
under torch.autograd (build container only; needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_grads.py

Two kinds of record per case (weights / inputs / targets are regenerated from seeds by svol_b200.synth):

  head   parameter gradients of SVANet.forward for SYNTHETIC upstream gradients w.r.t. the stacked logits / boxes
         (svol_b200.synth.make_upstream_grads): isolates the head's backward from the matcher.
  step   train.py:222-229 -- outputs = model(...); loss_dict = criterion(outputs, targets);
         sum(loss_dict[k] * weight_dict[k]).backward() -- losses and parameter gradients.

The model is built with input_dropout = 0 (the CUDA training path implements the deterministic network).
A parameter's gradient is stored as its L2 norm, its sum and a strided sample (every STRIDE-th element), fp32 and
fp64 runs, so the fixtures stay small.
"""
import os
import sys
from dataclasses import replace

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.argv = ["x"]

import numpy as np
import scipy
import torch

from lib.modeling.svanet import build_svanet            # noqa: E402  (reference)
from lib.modeling.loss import build_loss                # noqa: E402  (reference)

from svol_b200 import synth                             # noqa: E402

VERSIONS = np.array([f"torch={torch.__version__}", f"scipy={scipy.__version__}", f"numpy={np.__version__}"])
STRIDE = 997
torch.set_num_threads(8)


def summarise(rec, tag, grads):
    for k, g in grads.items():
        g = g.detach().double().numpy().ravel()
        rec[f"{tag}/norm/{k}"] = np.float64(np.linalg.norm(g))
        rec[f"{tag}/sum/{k}"] = np.float64(g.sum())
        rec[f"{tag}/sample/{k}"] = g[::STRIDE].astype(np.float32)


def case(name, cfg_name, batch, seed, padded):
    cfg = replace(synth.CONFIGS[cfg_name], input_dropout=0.0)
    sd = synth.random_state_dict(cfg, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    targets = synth.targets_to_torch(synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"]))
    rec = {"versions": VERSIONS, "batch": batch, "seed": seed, "padded": padded, "stride": STRIDE}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ns = cfg.to_namespace()
        model = build_svanet(ns)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        model = model.to(dt).train()
        t = lambda a: torch.from_numpy(a).to(dt)
        out = model(t(inp["src_sketch"]), t(inp["src_sketch_mask"]), t(inp["src_video"]), t(inp["src_video_mask"]))
        logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
        boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
        # ---- head: synthetic upstream gradients
        model.zero_grad(set_to_none=True)
        torch.autograd.backward([logits, boxes], [t(gl), t(gb)], retain_graph=True)
        summarise(rec, f"head_{tag}", {k: p.grad for k, p in model.named_parameters() if p.grad is not None})
        # ---- full step (criterion on fp32 outputs, as the reference runs it)
        if tag == "f32":
            model.zero_grad(set_to_none=True)
            criterion = build_loss(ns).train()
            loss_dict = criterion(out, targets)
            wd = criterion.weight_dict
            total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)          # train.py:227-228
            total.backward()
            summarise(rec, "step_f32", {k: p.grad for k, p in model.named_parameters() if p.grad is not None})
            for k, v in loss_dict.items():
                rec["loss/" + k] = np.float32(float(v))
            rec["loss_total"] = np.float32(float(total))
    np.savez_compressed(os.path.join(HERE, f"grads_{name}.npz"), **rec)
    n = sum(1 for k in rec if k.startswith("head_f32/norm/"))
    err = max(abs(rec[k] - rec[k.replace("f32", "f64")]) / (rec[k.replace("f32", "f64")] + 1e-30)
              for k in rec if k.startswith("head_f32/norm/"))
    print("grads", name, n, "parameters; max rel |norm_f32 - norm_f64| =", float(err), "loss", float(rec["loss_total"]))


if __name__ == "__main__":
    case("C1a_b2", "C1a", 2, 0, True)
    case("C1b_b2", "C1b", 2, 1, True)
