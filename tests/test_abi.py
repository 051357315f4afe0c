"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
that include/svol_b200.h declares, its structures match the ctypes mirrors, and the host-side
modules keep the reference's state_dict keys, builders and targets schema."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from svol_b200 import _lib, synth
from svol_b200.modeling import build_loss, build_matcher, build_svanet
from svol_b200.modeling.targets import flatten_targets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "svol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svol_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.get_lib()
    declared = _header_functions()
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.svol_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match():
    lib = _lib.get_lib()
    for which, st in enumerate((_lib.GemmArgs, _lib.AttnArgs, _lib.MatchArgs, _lib.CriterionArgs, _lib.GemmEpilogue)):
        assert lib.svol_sizeof_args(which) == ctypes.sizeof(st)


def test_no_cpu_fallback():
    """Without a CUDA device every product entry point must refuse to run rather than fall back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = synth.CONFIGS["tiny"]
    model = build_svanet(cfg.to_namespace()).eval()
    inp = synth.make_inputs(cfg, 2, 0)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            model(*(torch.from_numpy(inp[k]) for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")))
    crit = build_loss(cfg.to_namespace())
    lg, bx = synth.make_predictions(cfg, 2, 0)
    out = {"pred_logits": torch.from_numpy(lg[-1]), "pred_boxes": torch.from_numpy(bx[-1]), "aux_outputs": []}
    with pytest.raises(RuntimeError):
        crit(out, synth.targets_to_torch(synth.make_targets(cfg, 2, 0)))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "svol_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|from\s+\.*oracle|importlib.*oracle", src, flags=re.M), \
                    f"{f} imports the oracle"


@pytest.mark.parametrize("cfgname", ["C1a", "C2", "C2n4"])
def test_state_dict_keys_match_reference(golden_dir, cfgname):
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))[cfgname]
    cfg = synth.CONFIGS[cfgname]
    model = build_svanet(cfg.to_namespace())
    sd = model.state_dict()
    assert sorted(sd.keys()) == sorted(keys.keys())
    for k, shape in keys.items():
        assert list(sd[k].shape) == shape, k
    # synthetic weights load strictly, as reference checkpoints do (test.py:72-88)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()}, strict=True)


def test_build_loss_weight_dict_and_buffers():
    cfg = synth.CONFIGS["C2n4"]
    crit = build_loss(cfg.to_namespace())
    assert crit.weight_dict == {**{"loss_bbox": 5.0, "loss_giou": 1.0, "loss_label": 2.0},
                                **{f"{k}_{i}": v for i in range(3) for k, v in
                                   {"loss_bbox": 5.0, "loss_giou": 1.0, "loss_label": 2.0}.items()}}
    assert torch.allclose(crit.empty_weight, torch.tensor([1.0, 0.1]))
    assert type(build_matcher(cfg.to_namespace())).__name__ == "PerFrameMatcher"


def test_flatten_targets_offsets():
    cfg = synth.CONFIGS["tiny"]
    targets = synth.make_targets(cfg, 3, 5)
    flat = flatten_targets(synth.targets_to_torch(targets), torch.device("cpu"), True, cfg.num_frames,
                           cfg.num_queries, cfg.num_queries_per_frame)
    counts = np.array([n for t in targets for n in t["num_boxes_per_frame"]])
    assert flat.P == 3 * cfg.num_frames and flat.S == counts.sum()
    assert np.array_equal(flat.tgt_off.numpy(), np.concatenate([[0], np.cumsum(counts)]))
    assert flat.K == np.minimum(counts, cfg.num_queries_per_frame).sum()
    assert flat.cost_total == (counts * cfg.num_queries_per_frame).sum()
    assert flat.match_video.numel() == flat.K
    boxes = np.stack([o["bbox"] for t in targets for fr in t["bboxes"].values() for o in fr])
    assert np.array_equal(flat.tgt_boxes.numpy(), boxes)
    flat_v = flatten_targets(synth.targets_to_torch(targets), torch.device("cpu"), False, 0, cfg.num_queries, 0)
    assert flat_v.P == 3 and flat_v.rows_per_problem == cfg.num_queries


def test_gate_fused_shape_limit_is_a_host_query():
    """svol_gate_fused_supported is plain host code (no CUDA call): the one-launch gate takes clips whose token rows fit one
    cluster's shared memory (the headline L = 1568 does, the long clip L = 6272 falls back to gate_scores + gate_apply)."""
    from svol_b200 import _lib
    lib = _lib.get_lib()
    assert lib.svol_gate_fused_supported(1568) == 1 and lib.svol_gate_fused_supported(1) == 1
    assert lib.svol_gate_fused_supported(3384) == 1 and lib.svol_gate_fused_supported(3385) == 0
    assert lib.svol_gate_fused_supported(6272) == 0 and lib.svol_gate_fused_supported(0) == 0
