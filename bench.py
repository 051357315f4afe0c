#!/usr/bin/env python
"""Benchmark of the SVOL hot path: sketch-video pairs/s for head forward + Hungarian matching + set
criterion (BASELINE.json metric), on N GPUs of one node, one process per GPU.

  python bench.py --gpus 1 --steps 20 --warmup 5                       # this repo's CUDA path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                                 # the reference's CPU path (oracle/torch_port.py)

One step = one pass of the hot path over one batch of B synthetic pairs per GPU (weak scaling,
pairs are independent: no data-path collective).  Prints ONE JSON line on rank 0:

  value        pairs/s, device-timed (CUDA events over exactly K steps, max over ranks), inputs and
               flattened targets already resident in HBM
  e2e          pairs/s through the public API with HOST inputs: every step copies the step's frame
               features, sketch feature, masks and target boxes from pinned host memory, runs
               forward + criterion, and reads the losses back
  roofline     the dominant kernel of the step, timed live per launch with CUDA events
  cpu_baseline the reference's CPU path (oracle/torch_port.py: same ATen / scipy calls as the reference) on a bounded
               sample, rank 0, N=1
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sketch-video pairs/s fwd+match"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"],
                    help="reference: the reference's CPU path (what the driver's reference arm runs); reference-gpu: the same "
                         "eager PyTorch path on this GPU (fp32 and autocast bf16) -- the 'beat it on the same box' number")
    ap.add_argument("--batch", type=int, default=32, help="pairs per GPU per step")
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--config", default="C2", help="svol_b200.synth.CONFIGS key (C2 = BASELINE configs[1])")
    ap.add_argument("--graph", type=int, default=1, help="replay the forward as one CUDA graph")
    ap.add_argument("--ref-batch", type=int, default=2, help="pairs per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default="", help="write the per-kernel time table to this file")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2, 3, 4],
                    help="fwd_match mode: batches in flight (n: consecutive steps rotate over n streams / forward workspaces)")
    ap.add_argument("--mode", default="fwd_match", choices=["fwd_match", "train"],
                    help="fwd_match: the headline metric (BASELINE configs[1]); train: the training step of BASELINE "
                         "configs[2] (forward + matching + losses + backward + gradient all-reduce + AdamW)")
    return ap.parse_args()


def workload_config(args, cfg):
    return {"workload": f"{args.config}: SVOL head forward + PerFrameMatcher + SetCriterion (all decoder layers), "
                        f"B={args.batch} pairs/GPU, T={cfg.num_frames}, L={cfg.video_len}, D_in={cfg.input_vid_dim}, "
                        f"Q={cfg.num_queries}, layers={cfg.num_layers}",
            "pairs_per_gpu_per_step": args.batch, "layers": cfg.num_layers,
            "l2": "two alternating input sets (2 x %.0f MB frame features) and a >1 GB per-step activation stream: "
                  "no step finds its inputs in the 126 MB L2" % (args.batch * cfg.video_len * cfg.input_vid_dim * 4 / 1e6),
            "parallelism": f"dp{args.gpus} (independent pairs, no collective)",
            "batches_in_flight": getattr(args, "streams", 1),
            "inputs": "value: device-resident, written once into the forward plans' static input buffers "
                      "(HeadEngine.input_buffers; one CUDA-graph launch per forward); e2e: pinned host buffers copied every step"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples the SM clock and the throttle reasons while the timed region runs.  NVML is queried in-process
    (a thread, one sample every 10 ms): polling through a spawned `nvidia-smi -lms` loop was measured to slow the
    timed steps themselves by ~20 % (driver lock contention with kernel / graph launches).  nvidia-smi is only
    the fallback when the NVML binding is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.thread, self.stop_flag, self.max_mhz = None, None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: `index` is the CUDA ordinal of this process
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        self.rows = []
        if self.nvml is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = [r[0] for r in self.rows]
            reasons = sorted({k for r in self.rows for k in r[1]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 10 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_fn(cfg, batch, seed=0):
    """One step of the reference's CPU path: forward + matcher + criterion on `batch` pairs, through
    oracle/torch_port.py -- the reference's algorithm restated with the reference's own library calls (ATen
    multi_head_attention_forward / layer_norm / gelu / cdist / cross_entropy on all host threads, one scipy
    linear_sum_assignment per frame, the full cross-batch cost matrix).  The reference itself is Python and is not
    shipped to the GPU box, hence "kind": "port"; the port is pinned to the reference's golden outputs in
    tests/test_oracle_golden.py."""
    import torch
    from oracle import torch_port as tp
    from svol_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; use every host core
    sd = tp.state_dict_to_torch(synth.random_state_dict(cfg, seed))
    inp = synth.make_inputs(cfg, batch, seed, padded=True)
    targets = synth.make_targets(cfg, batch, seed, frame_mask=inp["frame_mask"])
    t = {k: torch.from_numpy(inp[k]) for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}

    def step():
        out = tp.svanet_forward(sd, t["src_sketch"], t["src_sketch_mask"], t["src_video"], t["src_video_mask"],
                                nheads=cfg.nheads)
        return tp.set_criterion(out, targets, cfg)
    return step


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from svol_b200 import synth
    from dataclasses import replace
    cfg = replace(synth.CONFIGS[args.config], num_layers=args.layers)
    b = args.ref_batch
    step = cpu_reference_step_fn(cfg, b)
    sec = time_cpu(step, args.steps, args.warmup)
    value = b / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(args, cfg), "pairs_per_step_cpu": b},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{b} pairs per step x {args.steps} steps of the same workload (the reference's CPU path "
                                       f"restated with its own ATen / scipy calls, oracle/torch_port.py, all host threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference_gpu(args):
    """The reference's eager PyTorch path ON THE B200: oracle/torch_port.py --
    the reference's own ATen / scipy call sequence (cross_modal_transformer.py:105-160 with nn.MultiheadAttention's
    need_weights=True score materialisation, matcher.py:38-119 with its full cross-batch cost matrix, the D2H copy and one
    scipy call per frame, loss.py:126-157) -- with every tensor on cuda:0, in fp32 (TF32 off, PyTorch's default: what the
    reference's scripts run) and under torch.autocast(bfloat16).  Same workload, batch and step definition as our arm;
    CUDA events around K steps; targets resident as the nested dict the reference consumes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from dataclasses import replace
    from oracle import torch_port as tp
    from svol_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg = replace(synth.CONFIGS[args.config], num_layers=args.layers)
    B = args.batch
    sd = {k: v.to(dev) for k, v in tp.state_dict_to_torch(synth.random_state_dict(cfg, 0)).items()}
    sets = []
    for s_ in range(2):
        inp = synth.make_inputs(cfg, B, seed=s_, padded=True)
        tg = synth.targets_to_torch(synth.make_targets(cfg, B, seed=s_, frame_mask=inp["frame_mask"]))
        sets.append(({k: torch.from_numpy(inp[k]).to(dev) for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}, tg))

    def step(i):
        t, tg = sets[i & 1]
        out = tp.svanet_forward(sd, t["src_sketch"], t["src_sketch_mask"], t["src_video"], t["src_video_mask"], nheads=cfg.nheads)
        out = {"pred_logits": out["pred_logits"].float(), "pred_boxes": out["pred_boxes"].float(),
               "aux_outputs": [{k: v.float() for k, v in a.items()} for a in out["aux_outputs"]]}
        return tp.set_criterion(out, tg, cfg)

    def timed(steps, warmup):
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def timed_parts(steps):
        """fp32: head forward alone and matcher + criterion alone (wall clock around a synchronised region: the
        reference's matcher is host code -- a D2H copy of the cost matrix and one scipy call per frame)."""
        t, tg = sets[0]
        fwd = lambda: tp.svanet_forward(sd, t["src_sketch"], t["src_sketch_mask"], t["src_video"], t["src_video_mask"], nheads=cfg.nheads)
        out = fwd()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fwd()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(steps):
            tp.set_criterion(out, tg, cfg)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return (t1 - t0) / steps * 1e3, (t2 - t1) / steps * 1e3

    sampler = ClockSampler(dev.index or 0)
    steps, warmup = args.steps, max(args.warmup, 3)
    with torch.no_grad():
        sampler.start()
        ms_f32 = timed(steps, warmup)
        clocks = sampler.stop()
        ms_fwd, ms_crit = timed_parts(steps)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_bf16 = timed(steps, warmup)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ms_tf32 = timed(steps, warmup)
    line = {"impl": "reference-gpu", "metric": METRIC, "value": B / (ms_f32 * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_f32, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, cfg), "clocks": clocks,
            "variants": {"fp32": {"value": B / (ms_f32 * 1e-3), "ms_per_step": ms_f32},
                         "fp32_tf32_matmul": {"value": B / (ms_tf32 * 1e-3), "ms_per_step": ms_tf32},
                         "autocast_bf16": {"value": B / (ms_bf16 * 1e-3), "ms_per_step": ms_bf16}},
            "parts_fp32_ms": {"head_forward": ms_fwd, "matcher_and_criterion": ms_crit},
            "what": "oracle/torch_port.py on cuda: the reference's eager ATen / scipy call sequence (pinned to the reference's "
                    "golden outputs on CPU), torch " + torch.__version__,
            "gpu": torch.cuda.get_device_name(dev)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    if args.mode == "train":
        return run_train(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args)
    # stdout carries exactly one JSON line: library banners (NCCL prints its version to stdout) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from svol_b200 import comm, synth
    from svol_b200.modeling import build_loss, build_svanet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = comm.bind_to_gpu_numa_node(local)          # pinned staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = replace(synth.CONFIGS[args.config], num_layers=args.layers)
    B = args.batch
    ns = cfg.to_namespace()
    ns.use_cuda_graph = bool(args.graph)
    model = build_svanet(ns)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()}, strict=True)
    model = model.to(dev).eval()
    criterion = build_loss(ns).to(dev)

    # two alternating synthetic input sets per rank (different pairs on every rank)
    sets = []
    for s in range(2):
        inp = synth.make_inputs(cfg, B, seed=100 * rank + s, padded=True)
        tg = synth.targets_to_torch(synth.make_targets(cfg, B, seed=100 * rank + s, frame_mask=inp["frame_mask"]))
        host = {k: torch.from_numpy(inp[k]).pin_memory() for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}
        devt = {k: v.to(dev) for k, v in host.items()}
        sets.append({"host": host, "dev": devt, "targets": tg})

    # Two batches in flight: step i runs on stream i & 1 (its own forward workspace and graph, HeadEngine.plan_for), so
    # the low-occupancy tail of one step (object-query chain, heads, matcher, criterion) overlaps the head of the next.
    # Pairs are independent; every step's work completes inside the timed region (both streams are joined before e1).
    step_streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else None

    # With two streams, input set s lives in the static input buffers of stream s's forward plan (HeadEngine.input_buffers:
    # what a producer that already runs on the device writes into): nothing is copied per step and the whole forward,
    # input LayerNorm included, is one CUDA-graph launch.  Otherwise the caller-owned tensors are read in place.
    resident = None
    if step_streams is not None and args.graph and not os.environ.get("SVOL_BENCH_NO_RESIDENT_BUFFERS"):
        resident = []
        L_tok, d_in = sets[0]["dev"]["src_video"].shape[1:]
        for si, st in enumerate(step_streams):
            with torch.cuda.stream(st):
                bufs = model.engine.input_buffers(B, L_tok, d_in)
                d = sets[si & 1]["dev"]
                bufs["src_video"].copy_(d["src_video"])
                bufs["src_sketch"].copy_(d["src_sketch"].reshape(B, -1))
                bufs["src_video_mask"].copy_(d["src_video_mask"])
                resident.append({"src_sketch": bufs["src_sketch"].view(B, 1, -1), "src_sketch_mask": d["src_sketch_mask"],
                                 "src_video": bufs["src_video"], "src_video_mask": bufs["src_video_mask"]})
        torch.cuda.synchronize()

    n_lanes = len(step_streams) if step_streams is not None else 1

    def step_resident(i):
        lane_i = i % n_lanes
        s = sets[lane_i & 1] if resident is not None else sets[i & 1]      # stream k always works on input set k & 1
        d = resident[lane_i] if resident is not None else s["dev"]
        if step_streams is None:
            out = model(d["src_sketch"], d["src_sketch_mask"], d["src_video"], d["src_video_mask"])
            return criterion(out, s["targets"])
        with torch.cuda.stream(step_streams[i % len(step_streams)]):
            out = model(d["src_sketch"], d["src_sketch_mask"], d["src_video"], d["src_video_mask"])
            return criterion(out, s["targets"])

    def fork_streams():
        if step_streams is not None:
            for st in step_streams:
                st.wait_stream(torch.cuda.current_stream())

    def join_streams():
        if step_streams is not None:
            for st in step_streams:
                torch.cuda.current_stream().wait_stream(st)

    # ---- end-to-end path: host buffers in, losses out, through the public API (model(...), criterion(...)).
    # Two device staging sets; the H2D copy of step i+1 runs on a copy stream while step i computes, and the
    # host reads step i's losses (pinned, async D2H) while step i+1 runs.  Every step's copies are inside the
    # timed region; nothing is reused across steps (targets are re-flattened and re-uploaded every step).
    stages = [{k: torch.empty_like(v, device=dev) for k, v in sets[0]["host"].items()} for _ in range(2)]
    # variant: the caller keeps the frame features in bf16 on the host (a precomputed feature cache): half the PCIe bytes
    for s_ in sets:
        s_["host_bf16"] = dict(s_["host"], src_video=s_["host"]["src_video"].to(torch.bfloat16).pin_memory())
    stages_bf16 = [dict(st, src_video=torch.empty_like(st["src_video"], dtype=torch.bfloat16)) for st in stages]
    loss_host = [torch.empty((cfg.num_layers, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    e2e_sink = []

    e2e_cfg = {"host": "host", "stages": stages, "two_lanes": False}

    def h2d(i):
        slot = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[slot])                      # the step that last used this slot is finished
            for k, v in sets[i & 1][e2e_cfg["host"]].items():
                e2e_cfg["stages"][slot][k].copy_(v, non_blocking=True)  # H2D of step i's inputs
            ev_copied[slot].record(copy_stream)

    def run_e2e(steps):
        # step i runs on step stream i & 1 (its own forward workspace / criterion workspace, like the resident loop): the
        # low-occupancy tail of one step overlaps the head of the next, and with bf16 features the loop is no longer bound by
        # one stream's device time
        main = torch.cuda.current_stream()
        # (only when the upload is shorter than a step -- bf16 features: with fp32 features the loop is bound by the PCIe copy,
        # and two overlapping steps only delay the loss read-back that paces the host's next copy: 2.05 vs 1.95 ms per step)
        lanes = step_streams[:2] if (e2e_cfg["two_lanes"] and step_streams is not None and len(step_streams) >= 2) else [main, main]
        for st_ in set(lanes):
            st_.wait_stream(main)
        h2d(0)
        for i in range(steps):
            slot = i & 1
            if i + 1 < steps:
                h2d(i + 1)
            lane = lanes[slot]
            with torch.cuda.stream(lane):
                lane.wait_event(ev_copied[slot])
                st = e2e_cfg["stages"][slot]
                criterion.matcher._cache._key = None                   # new targets every step: host walk + H2D
                out = model(st["src_sketch"], st["src_sketch_mask"], st["src_video"], st["src_video_mask"])
                losses = criterion(out, sets[i & 1]["targets"])
                vals = torch.stack([losses[k] for k in losses])
                loss_host[slot].view(-1)[: vals.numel()].copy_(vals, non_blocking=True)    # D2H of the step's result
                ev_done[slot].record(lane)
            if i > 0:
                ev_done[slot ^ 1].synchronize()                        # host consumes the previous step's losses
                e2e_sink.append(float(loss_host[slot ^ 1][0, 0]))
        ev_done[(steps - 1) & 1].synchronize()
        e2e_sink.append(float(loss_host[(steps - 1) & 1][0, 0]))
        for st_ in set(lanes):
            main.wait_stream(st_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []

    def timed(fn, steps, warmup, whole_loop=False, sampler=None):
        with torch.no_grad():
            if whole_loop:
                fn(warmup)
            else:
                for i in range(warmup):
                    fn(i)
            barrier()
            if sampler is not None:
                sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if sampler is not None:
                torch.cuda.nvtx.range_push("timed")      # `ncu --nvtx --nvtx-include "timed/"`: the steady-state steps only
            e0.record()
            t_host = time.perf_counter()
            if whole_loop:
                fn(steps)
            else:
                fork_streams()                      # the step streams start after e0 ...
                for i in range(steps):
                    fn(i)
                join_streams()                      # ... and e1 is recorded after both have drained
            host_ms.append((time.perf_counter() - t_host) / steps * 1e3)     # host time to ENQUEUE a step (not waiting for it)
            e1.record()
            if sampler is not None:
                torch.cuda.nvtx.range_pop()
            barrier()
        return comm.max_over_ranks(e0.elapsed_time(e1), device=dev) / steps      # slowest rank

    sampler = ClockSampler(local) if rank == 0 else None
    ms_step = timed(step_resident, args.steps, max(args.warmup, 3), sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(run_e2e, args.steps, 3, whole_loop=True)
    e2e_cfg.update(host="host_bf16", stages=stages_bf16, two_lanes=True)
    ms_e2e_bf16 = timed(run_e2e, args.steps, 3, whole_loop=True)
    with torch.no_grad():
        step_resident(0)
    torch.cuda.synchronize()
    criterion.check_status()

    flat = criterion.last_indices[2]
    h2d = sum(v.numel() * v.element_size() for v in sets[0]["host"].values()) + flat.tgt_boxes.numel() * 4 + \
        (flat.tgt_off.numel() + flat.match_off.numel() + flat.video_tgt_off.numel() + flat.video_match_off.numel()
         + flat.match_video.numel()) * 4 + flat.cost_off.numel() * 8
    d2h = cfg.num_layers * 4 * 4
    launches_per_step = model.engine.launches_per_forward + 3        # + match, match_finalize, criterion

    # end-to-end index agreement with the reference on the headline batch (north_star; svol_b200/parity.py): reference
    # forward -> reference matcher (committed fixture tests/golden/head_C2_b32.npz) vs this GPU's forward -> matcher
    agreement = None
    gpath = os.path.join(ROOT, "tests", "golden", "head_C2_b32.npz")
    if rank == 0 and args.config == "C2" and args.layers == 2 and os.path.exists(gpath):
        from svol_b200.parity import index_agreement
        golden = dict(np.load(gpath))
        with torch.no_grad():           # the fixture's weights (same architecture; the timed runs above used seed 0)
            model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, int(golden["seed"])).items()})
        agreement = index_agreement(model, criterion.matcher, cfg, golden, dev)
        with torch.no_grad():
            model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()})

    roofline, breakdown = None, None
    cpu_baseline = None
    if rank == 0:
        roofline, breakdown = kernel_breakdown(model, sets[0], cfg, B, dev)
        if args.breakdown:
            with open(args.breakdown, "w") as f:
                f.write(breakdown)
        if world == 1 and not args.no_cpu_baseline:
            b = args.ref_batch
            sec = time_cpu(cpu_reference_step_fn(cfg, b), 4, 1)
            cpu_baseline = {"value": b / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"{b} pairs per step x 4 steps (1 warm-up) of the same workload, the reference's CPU "
                                      f"path restated with its own ATen / scipy calls (oracle/torch_port.py, all host "
                                      f"threads), {sec * 1e3:.0f} ms per step"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pairs = B * world
    line = {"metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, cfg),
            "clocks": clocks,
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e},
            # not the headline: same loop with the frame features kept in bf16 on the host (feature cache), i.e. half the
            # PCIe traffic; the fp32 e2e above is bound by the 103 MB / step upload
            "e2e_bf16_features": {"value": pairs / (ms_e2e_bf16 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_bf16,
                                  "h2d_bytes_per_step": int(h2d - sets[0]["host"]["src_video"].numel() * 2)},
            "gpu_launches": int(launches_per_step * args.steps), "host_numa_bound": bool(numa_bound),
            "host_enqueue_ms_per_step": host_ms[0],
            "roofline": roofline, "cpu_baseline": cpu_baseline, "index_agreement_vs_reference": agreement,
            "tflops_algorithmic": synth.algorithmic_flops_per_pair(cfg) * pairs / (ms_step * 1e-3) / 1e12}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ training step (config 3)
TRAIN_METRIC = "sketch-video pairs/s training step"


def run_train(args):
    """BASELINE configs[2]: SVOL training step, data-parallel (per-GPU batch fixed, weak scaling): forward in train mode,
    PerFrameMatcher + SetCriterion on every decoder layer, backward, ONE NCCL all-reduce of the flat fp32 gradient
    buffer, fused AdamW (lr 1e-4, wd 1e-4: lib/configs.py:71-78), input dropout 0.4 as in lib/configs.py:127 (this library's
    counter-based mask, see svol_b200/train_engine.py).  `--impl reference` times the same step through the oracle's
    autograd (oracle/torch_port.py: the reference's ATen / scipy calls + torch.autograd + torch.optim.AdamW) on the
    host cores, rank 0 only."""
    from dataclasses import replace
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    from svol_b200 import synth
    cfg = replace(synth.CONFIGS[args.config], num_layers=args.layers)        # input_dropout = 0.4 (lib/configs.py:127)
    workload = {"workload": f"C3 on {args.config}: SVOL training step (train-mode forward + PerFrameMatcher + SetCriterion on all "
                            f"decoder layers + backward + gradient all-reduce + AdamW), B={args.batch} pairs/GPU, T={cfg.num_frames}, "
                            f"L={cfg.video_len}, D_in={cfg.input_vid_dim}, Q={cfg.num_queries}, layers={cfg.num_layers}, "
                            f"input_dropout={cfg.input_dropout}",
                "pairs_per_gpu_per_step": args.batch, "layers": cfg.num_layers,
                "l2": ">4 GB of saved activations per step: nothing survives in the 126 MB L2 between steps",
                "parallelism": f"dp{args.gpus}: one NCCL all-reduce of the flat fp32 gradient buffer per step"}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import torch_port as tp
        torch.set_num_threads(os.cpu_count())
        b = args.ref_batch
        sd = {k: v.clone().requires_grad_(True) for k, v in tp.state_dict_to_torch(synth.random_state_dict(cfg, 0)).items()}
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, weight_decay=1e-4)
        inp = synth.make_inputs(cfg, b, 0, padded=True)
        targets = synth.targets_to_torch(synth.make_targets(cfg, b, 0, frame_mask=inp["frame_mask"]))
        wd = {k: w for k, w in (("loss_bbox", cfg.set_cost_bbox), ("loss_giou", cfg.set_cost_giou), ("loss_label", cfg.set_cost_class))}
        wd.update({f"{k}_{i}": v for i in range(cfg.num_layers - 1) for k, v in list(wd.items())[:3]})

        def step():
            grads, _, _ = tp.training_step_gradients(sd, inp["src_sketch"], inp["src_sketch_mask"], inp["src_video"],
                                                     inp["src_video_mask"], targets, cfg, wd, dropout=(cfg.input_dropout, 1))
            for k, v in sd.items():
                v.grad = grads.get(k)
            opt.step()
        sec = time_cpu(step, max(1, min(args.steps, 3)), 1)
        value = b / sec
        print(json.dumps({"impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": dict(workload, pairs_per_step_cpu=b),
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"{b} pairs per step: oracle/torch_port.py forward + criterion under "
                                                     "torch.autograd + torch.optim.AdamW, all host threads"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    from svol_b200 import _lib, comm
    from svol_b200.modeling import build_loss, build_svanet
    from svol_b200.optim import FusedAdamW
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = comm.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    ns = cfg.to_namespace()
    model = build_svanet(ns)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.random_state_dict(cfg, 0).items()}, strict=True)
    model = model.to(dev).train()
    criterion = build_loss(ns).to(dev).train()
    opt = FusedAdamW(model, lr=1e-4, weight_decay=1e-4)
    model.train_engine.publish_grads = False        # the fused optimizer reads the flat gradient buffer directly
    wd = criterion.weight_dict
    sets = []
    for s_ in range(2):
        inp = synth.make_inputs(cfg, B, seed=100 * rank + s_, padded=True)
        tg = synth.targets_to_torch(synth.make_targets(cfg, B, seed=100 * rank + s_, frame_mask=inp["frame_mask"]))
        host = {k: torch.from_numpy(inp[k]).pin_memory() for k in ("src_sketch", "src_sketch_mask", "src_video", "src_video_mask")}
        sets.append({"host": host, "dev": {k: v.to(dev) for k, v in host.items()}, "targets": tg})
    # end-to-end: host (pinned) inputs in, loss out, every step.  Two device staging sets: the H2D copy of step i + 1 runs
    # on a copy stream while step i computes (the reference's DataLoader(pin_memory) + .to(non_blocking) pipeline,
    # train.py:213-221); the host reads step i's loss (pinned, async D2H) while step i + 1 runs.
    stages = [{k: torch.empty_like(v, device=dev) for k, v in sets[0]["host"].items()} for _ in range(2)]
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    e2e_sink = []

    def train_step(d, targets):
        out = model(d["src_sketch"], d["src_sketch_mask"], d["src_video"], d["src_video_mask"])
        loss_dict = criterion(out, targets)
        total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)           # train.py:227-228
        total.backward()
        scale, _ = comm.allreduce_gradients(model.train_engine.grad_flat)        # one collective (no-op on 1 GPU)
        opt.step(from_engine=True, grad_scale=scale)
        return total.detach()

    def step_resident(i):
        return train_step(sets[i & 1]["dev"], sets[i & 1]["targets"])

    def h2d(i):
        slot = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[slot])
            for k, v in sets[i & 1]["host"].items():
                stages[slot][k].copy_(v, non_blocking=True)                       # H2D of step i's inputs
            ev_copied[slot].record(copy_stream)

    def run_e2e(steps):
        main = torch.cuda.current_stream()
        h2d(0)
        last = None
        for i in range(steps):
            slot = i & 1
            if i + 1 < steps:
                h2d(i + 1)
            main.wait_event(ev_copied[slot])
            criterion.matcher._cache._key = None                                  # targets walked + uploaded every step
            total = train_step(stages[slot], sets[i & 1]["targets"])
            loss_host[slot].copy_(total.reshape(1), non_blocking=True)            # D2H of the step's loss
            ev_done[slot].record(main)
            if i > 0:
                ev_done[slot ^ 1].synchronize()
                e2e_sink.append(float(loss_host[slot ^ 1][0]))
        ev_done[(steps - 1) & 1].synchronize()
        e2e_sink.append(float(loss_host[(steps - 1) & 1][0]))
        return e2e_sink[-1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, whole_loop=False):
        if whole_loop:
            fn(warmup)
        else:
            for i in range(warmup):
                fn(i)
        barrier()
        if sampler is not None:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole_loop:
            last = fn(steps)
        else:
            for i in range(steps):
                last = fn(i)
        e1.record()
        barrier()
        return comm.max_over_ranks(e0.elapsed_time(e1), device=dev) / steps, last

    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, last = timed(step_resident, args.steps, max(args.warmup, 3), sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(run_e2e, args.steps, 3, whole_loop=True)
    criterion.check_status()
    plan = model.train_engine._last
    launches = len(plan["fwd"].calls) + len(plan["bwd"].calls) + 2 * 3 + 4       # + attention-backward pairs, matcher, criterion fwd/bwd, AdamW
    roofline = None
    if rank == 0:
        # dominant kernel family of the step: the video self-attention backward (delta + dQ + dK/dV kernels), timed live
        st = torch.cuda.current_stream().cuda_stream
        names = [c[0] for c in plan["bwd"].calls]
        idx = [i for i, n in enumerate(names) if n.endswith("sa_attn_bwd")]
        times = []
        for i in idx:
            name, fn, a = plan["bwd"].calls[i]
            for _ in range(2):
                _lib.check(fn(*a, st), name)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                _lib.check(fn(*a, st), name)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 5)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = float(peaks.get("bf16_tflops_sustained", 1374.6))
        flops = 10.0 * cfg.video_len * cfg.video_len * cfg.hidden_dim * B            # 5 GEMMs of 2 L^2 d per sample
        ms_k = float(np.mean(times))
        roofline = {"kernel": "attention backward (video self): delta + dQ + dK/dV", "bound": "tensor",
                    "achieved": flops / (ms_k * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops / (ms_k * 1e-3) / 1e12 / peak, "traffic": None, "ms_per_launch": ms_k,
                    "share_of_step": len(idx) * ms_k / ms_step,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1374.6",
                    "algorithmic_flops_per_launch": flops}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pairs = B * world
    h2d = sum(v.numel() * v.element_size() for v in sets[0]["host"].values())
    line = {"metric": TRAIN_METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload, "clocks": clocks,
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches * args.steps), "host_numa_bound": bool(numa_bound), "roofline": roofline,
            "cpu_baseline": None, "final_loss": float(last),
            "tflops_algorithmic": 3.0 * synth.algorithmic_flops_per_pair(cfg) * pairs / (ms_step * 1e-3) / 1e12}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


def kernel_breakdown(model, inset, cfg, B, dev):
    """Times every launch of the forward plan individually (CUDA events on the launching stream, 5 repeats
    after a warm-up, median) and derives the roofline entry of the dominant kernel family."""
    import torch
    from svol_b200 import synth
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks.update(json.load(open(pk)))
        peaks["_src"] = "measured"
    eng = model.engine
    d = inset["dev"]
    with torch.no_grad():           # fills the current stream's workspace (plans are per stream) with this input set
        model(d["src_sketch"], d["src_sketch_mask"], d["src_video"], d["src_video_mask"])
    plan = eng.plan_for(B, cfg.video_len, cfg.input_vid_dim)
    stream = torch.cuda.current_stream().cuda_stream
    reps = 5
    times = {}
    with torch.no_grad():
        plan.run(stream)
        torch.cuda.synchronize()
        for _ in range(reps):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan.calls) + 1)]
            evs[0].record()
            for i, (name, fn, a) in enumerate(plan.calls):
                rc = fn(*a, stream)
                assert rc == 0, name
                evs[i + 1].record()
            torch.cuda.synchronize()
            for i, (name, _, _) in enumerate(plan.calls):
                times.setdefault(name, []).append(evs[i].elapsed_time(evs[i + 1]))
    med = {k: float(np.median(v)) for k, v in times.items()}
    total = sum(med.values())
    L, Q, d, ff, D = cfg.video_len, cfg.num_queries, cfg.hidden_dim, cfg.dim_feedforward, cfg.input_vid_dim
    M, MQ, NL = B * L, B * Q, cfg.num_layers

    def fam(name):
        n = name.split(".")[-1]
        return {"sa_attn": "attention(video self)", "ta_attn": "attention(query self)", "ca_attn": "attention(cross)",
                "ffn1": "fused FFN (video tokens)", "ffn2": "fused FFN (queries)"}.get(n, n)
    flops = {"attention(video self)": 4.0 * L * L * d * B, "attention(query self)": 4.0 * Q * Q * d * B,
             "attention(cross)": 4.0 * Q * L * d * B, "fused FFN (video tokens)": 4.0 * M * d * ff,
             "fused FFN (queries)": 4.0 * MQ * d * ff,
             "sa_qk": 2.0 * M * d * 2 * d, "sa_v": 2.0 * M * d * d, "sa_out": 2.0 * M * d * d, "in_proj0": 2.0 * M * D * d,
             "in_proj1": 2.0 * M * d * d, "ca_k": 2.0 * M * d * d, "ca_v": 2.0 * M * d * d}
    agg = {}
    for name, t in med.items():
        f = fam(name)
        a = agg.setdefault(f, [0.0, 0])
        a[0] += t
        a[1] += 1
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    lines = [f"per-kernel-family device time, B={B} L={L} layers={NL} (median of {reps}, CUDA events between launches)",
             f"{'family':34s} {'launches':>8s} {'ms total':>10s} {'share':>7s} {'TFLOP/s':>9s}"]
    for f, (t, n) in rows:
        tf = flops.get(f, 0.0) * (n if f in flops else 0) / (t * 1e-3) / 1e12 if t > 0 else 0.0
        lines.append(f"{f:34s} {n:8d} {t:10.4f} {100 * t / total:6.1f}% {tf:9.1f}")
    lines.append(f"{'TOTAL':34s} {len(med):8d} {total:10.4f}")
    top, (t_top, n_top) = rows[0]
    per_launch_ms = t_top / n_top
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(top, {}).get("bytes")
    if top in flops:
        achieved = flops[top] / (per_launch_ms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops_sustained"])
        roof = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "ms_per_launch": per_launch_ms, "share_of_forward": t_top / total,
                "peak_source": f"{peaks['_src']} bf16_tflops_sustained (kernel timed inside the step)",
                "algorithmic_flops_per_launch": flops[top]}
        if top.startswith("attention"):
            # informational: at head_dim 32 the softmax's exponentials, not the tensor pipe, bound a flash-style attention tile
            # (one ex2 per 128 tensor FLOPs; MUFU.EX2 issues 16 results / clk / SM on sm_100a, tools/ubench/mufu.cu ->
            # profiles/r02_ubench_mufu_packed.txt).  exps = B * H * Lq * Lk of the shape; the clock is the one sampled under load.
            H = 8
            exps = {"attention(video self)": float(B) * H * L * L, "attention(query self)": float(B) * H * Q * Q,
                    "attention(cross)": float(B) * H * Q * L}[top]
            props = torch.cuda.get_device_properties(dev)
            mhz = float(peaks.get("sm_mhz_under_load", 0) or 1965.0)
            floor_ms = exps / (16.0 * props.multi_processor_count * mhz * 1e6) * 1e3
            roof["limiting_pipe"] = {"pipe": "MUFU.EX2 (16 results/clk/SM measured)", "exps_per_launch": exps, "sm_mhz_assumed": mhz,
                                     "floor_ms_per_launch": floor_ms, "frac_of_floor": floor_ms / per_launch_ms}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s", "frac": None,
                "traffic": traffic, "ms_per_launch": per_launch_ms, "share_of_forward": t_top / total}
    return roof, "\n".join(lines) + "\n"


if __name__ == "__main__":
    main()
