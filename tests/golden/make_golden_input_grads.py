"""Generates tests/golden/grads_inputs_<case>.npz: gradients of the REFERENCE's SVANet.forward (train mode, input_dropout 0)
with respect to its INPUT FEATURES src_video / src_sketch under torch.autograd, for the synthetic upstream gradients of
svol_b200.synth.make_upstream_grads (build container only; needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_input_grads.py

train.py:72 optimises backbone + head, so the head's backward has to hand d(loss)/d(features) back to the backbone
(model.py:18-28).  Stored: L2 norm, sum and a strided sample of each input gradient, fp32 and fp64 runs.
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

import numpy as np                                       # noqa: E402
import torch                                             # noqa: E402

from lib.modeling.svanet import build_svanet             # noqa: E402  (reference)

from svol_b200 import synth                              # noqa: E402

STRIDE = 97
torch.set_num_threads(8)


def case(name, cfg_name, batch, seed, padded):
    cfg = replace(synth.CONFIGS[cfg_name], input_dropout=0.0)
    sd = synth.random_state_dict(cfg, seed)
    inp = synth.make_inputs(cfg, batch, seed, padded=padded)
    gl, gb = synth.make_upstream_grads(cfg, batch, seed)
    rec = {"versions": np.array([f"torch={torch.__version__}"]), "batch": batch, "seed": seed, "padded": padded, "stride": STRIDE}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        model = build_svanet(cfg.to_namespace())
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        model = model.to(dt).train()
        t = lambda a: torch.from_numpy(a).to(dt)
        vid, sk = t(inp["src_video"]).requires_grad_(True), t(inp["src_sketch"]).requires_grad_(True)
        out = model(sk, t(inp["src_sketch_mask"]), vid, t(inp["src_video_mask"]))
        logits = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
        boxes = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
        torch.autograd.backward([logits, boxes], [t(gl), t(gb)])
        for key, g in (("src_video", vid.grad), ("src_sketch", sk.grad)):
            g = g.detach().double().numpy().ravel()
            rec[f"{tag}/norm/{key}"] = np.float64(np.linalg.norm(g))
            rec[f"{tag}/sum/{key}"] = np.float64(g.sum())
            rec[f"{tag}/sample/{key}"] = g[::STRIDE].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f"grads_inputs_{name}.npz"), **rec)
    print("input grads", name, {k: float(v) for k, v in rec.items() if "/norm/" in k})


if __name__ == "__main__":
    case("C1b_b2", "C1b", 2, 1, True)
    case("C1a_b2", "C1a", 2, 0, True)
