#!/usr/bin/env python
"""bench.py -- train iters/s of the multi-level AdaptSegNet step (BASELINE.json configs[1]:
720x1280 GTA5-shaped source + 512x1024 Cityscapes-shaped target, lambda-adv 0.0002/0.001, vanilla GAN).

  python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (restated, oracle/torch_ref)

Prints ONE JSON line (rank 0).  A step = one full training iteration: G forward/backward on source and
target, both discriminators, all three optimizers.  `value` is timed with the inputs resident in HBM,
`e2e` through the public API from pinned HOST buffers (H2D of both images + labels and D2H of the six
losses inside the timed region).  Timing is CUDA events bracketed by barrier + synchronize, max over
ranks.  The per-step working set (> 10 GB of trunk activations) exceeds the 126 MB L2, so consecutive
iterations do not hit each other's data.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train iters/s (720x1280 src+512x1024 tgt)"
UNIT = "iters/s"
SRC_HW = (720, 1280)
TGT_HW = (512, 1024)
SAMPLE_HW = (256, 512)  # bounded CPU sample: BASELINE.json configs[0] shape, src = tgt
SEED = 1338


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--level", default="multi-level", choices=["multi-level", "single-level"])
    ap.add_argument("--gan", default="Vanilla", choices=["Vanilla", "LS"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the iteration (up to the gradients) as a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (profiling runs)")
    ap.add_argument("--channels-last", type=int, default=1, help="run the (unchanged) trunk in channels_last")
    ap.add_argument("--no-kernel-events", action="store_true", help="skip the per-kernel event pass (ncu runs)")
    ap.add_argument("--cudnn-benchmark", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--trunk-dtype", default="tf32", choices=["tf32", "bf16"],
                    help="tf32: the trunk as the reference's GPU path computes it; bf16: the trunk modules under "
                         "torch.autocast(bfloat16) (execution mode of the untouched trunk, SURVEY.md 8f row 1)")
    ap.add_argument("--also-trunk-bf16", type=int, default=1,
                    help="after the headline measurement, time the same step once more with the trunk under bf16 autocast "
                         "and report it under 'trunk_bf16_autocast' (never the headline)")
    ap.add_argument("--tier", default="B", choices=["A", "B"],
                    help="B: low-res logits feed the fused upsample+softmax+CE kernel and the discriminators' input pack "
                         "(no full-res logits in HBM, SURVEY.md 8d); A: the reference's tensor-by-tensor chain")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the restated reference loop on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, level, gan):
    """iters/s of the full-size workload, extrapolated from a bounded sample on the host CPU.

    Sample: the same iteration at 256x512 source = target (configs[0]); the trunk, heads, upsample, CE and
    discriminators are all linear in the pixel count, so full-size rate = sample rate * sample_px / full_px."""
    import torch
    from oracle import torch_ref as TR

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    tr = TR.RefTrainer(level=level, gan=gan, device="cpu")
    src, lab, tgt = TR.synthetic_batch(SEED, SAMPLE_HW, SAMPLE_HW)
    for i in range(warmup):
        tr.step(src, lab, tgt, i_iter=i)
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(src, lab, tgt, i_iter=warmup + i)
    dt = (time.perf_counter() - t0) / steps
    full_px = SRC_HW[0] * SRC_HW[1] + TGT_HW[0] * TGT_HW[1]
    sample_px = 2 * SAMPLE_HW[0] * SAMPLE_HW[1]
    scale = sample_px / full_px
    return {"value": (1.0 / dt) * scale, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{steps} timed iterations (after {warmup} warm-up) of the restated reference loop "
                       f"(oracle/torch_ref.RefTrainer, torch {torch.__version__} CPU fp32, {cores} threads) at "
                       f"{SAMPLE_HW[0]}x{SAMPLE_HW[1]} src=tgt: {dt:.3f} s/iter; scaled by pixel count x{scale:.4f} "
                       f"to the 720x1280+512x1024 workload"),
            "sample_s_per_iter": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    base = cpu_reference_rate(steps, warmup, args.level, args.gan)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 / base["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    cfg = {"workload": f"{args.level} AdaptSegNet train step, DeeplabMulti(ResNet-101, 19 cls) + "
                       f"{'2x' if args.level == 'multi-level' else '1x'} FCDiscriminator, {args.gan} GAN, "
                       f"src 1x3x{SRC_HW[0]}x{SRC_HW[1]} + tgt 1x3x{TGT_HW[0]}x{TGT_HW[1]} per GPU, random init",
           "per_gpu_batch": "1 source + 1 target image", "global_pairs_per_step": world}
    if args.impl == "reference":
        cfg.update({"parallelism": "none (rank 0 only, host CPU)",
                    "hot_path": "the reference's own torch CPU ops (restated loop, oracle/torch_ref.RefTrainer)",
                    "execution": "eager, all host threads", "trunk": "ResNet-101 on torch CPU (oneDNN), fp32"})
        return cfg
    cfg.update({"parallelism": f"dp{world} (NCCL all-reduce of 3 flat gradient buffers per step)",
                "hot_path": "libasn_b200 sm_100a kernels (tcgen05 heads + discriminators, fused losses)",
                "execution": ("forward/backward of the iteration replayed as two CUDA graphs (generator part, discriminator "
                              "part; the generator's gradient all-reduce overlaps the second); fused optimizer steps eager"
                              if args.cuda_graph else "eager"),
                "trunk": ("ResNet-101 as PyTorch modules on cuDNN (" + ("bf16 autocast" if args.trunk_dtype == "bf16" else "TF32")
                          + (", channels_last" if args.channels_last else "") + "), timed, not rewritten"),
                "tier": ("B: upsample fused into its consumers (CE loss, discriminator input pack); no full-res logits in HBM"
                         if args.tier == "B" else "A: interp -> loss / softmax -> D tensor by tensor, as the reference"),
                "dedup": "the D-step's forward on the target prediction re-uses the activations of the G-step's identical "
                         "forward (same weights, same input: train...:617-618 vs :665-666) instead of recomputing them -- "
                         "SURVEY.md 8d's 544.6 GF variant; every other pass of the reference iteration is executed",
                "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no flush needed"})
    return cfg


# ------------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from adaptsegnet_b200 import ops, prof
    from adaptsegnet_b200.train_step import AdaptSegTrainer, TrainConfig
    from adaptsegnet_b200.utils.synthetic import synthetic_batch   # (this arm never imports oracle/)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)  # train_gta2cityscapes_multi.py:228

    torch.manual_seed(SEED)  # identical replicas
    trainer = AdaptSegTrainer(TrainConfig(level=args.level, gan=args.gan, lazy_upsample=args.tier == "B"), device=dev,
                              use_cuda_graph=bool(args.cuda_graph), channels_last=bool(args.channels_last),
                              trunk_bf16=args.trunk_dtype == "bf16")
    src_h, lab_h, tgt_h = synthetic_batch(SEED + rank, SRC_HW, TGT_HW)  # each rank its own pair
    src_h, lab_h, tgt_h = src_h.pin_memory(), lab_h.pin_memory(), tgt_h.pin_memory()
    src, lab, tgt = src_h.to(dev), lab_h.to(dev), tgt_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    it = [0]

    def step_resident(_):
        trainer.step(src, lab, tgt, i_iter=it[0])
        it[0] += 1

    n_losses = 6 if args.level == "multi-level" else 3
    losses_host = torch.empty(n_losses, dtype=torch.float32).pin_memory()

    def step_e2e(_):
        # host -> device copies of this step's inputs from pinned memory (straight into the graph's static
        # input buffers when replaying a CUDA graph), the step, and the losses back on the host
        if trainer.use_cuda_graph and trainer._graph is not None:
            s, l, t = trainer._static_in
            s.copy_(src_h, non_blocking=True)
            l.copy_(lab_h, non_blocking=True)
            t.copy_(tgt_h, non_blocking=True)
        else:
            s = src_h.to(dev, non_blocking=True)
            l = lab_h.to(dev, non_blocking=True)
            t = tgt_h.to(dev, non_blocking=True)
        out = trainer.step(s, l, t, i_iter=it[0])
        vals = torch.stack([v.float() for v in out.values()])
        losses_host.copy_(vals, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the losses every iteration
        it[0] += 1

    graph_note = None
    if trainer.use_cuda_graph:
        try:  # capture happens inside the first step; never let a capture problem take the benchmark down
            step_resident(0)
        except Exception as exc:  # noqa: BLE001
            graph_note = f"CUDA-graph capture failed ({type(exc).__name__}: {str(exc)[:120]}); ran eager"
            torch.cuda.synchronize()
            trainer.use_cuda_graph = False
            trainer._graph = None
    for _ in range(args.warmup):
        step_resident(0)
    # ---- device-resident timing with clock sampling ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(args.steps, step_resident)
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end timing from pinned host buffers ----
    ms_e2e = None
    if not args.no_e2e:
        step_e2e(0)
        ms_e2e = timed(args.steps, step_e2e)
    # ---- per-kernel CUDA-event timing, live inside K more steps.  Events cannot be recorded inside a captured
    #      graph, so this pass runs the same iteration eagerly (same kernels, same inputs, same order). ----
    graph_mode = trainer.use_cuda_graph
    if args.no_kernel_events:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * 1000.0 / (ms_total / args.steps), "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": ms_total / args.steps, "note": "profiling run: no kernel events"}),
                  flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    trainer.use_cuda_graph = False
    # on the stream the graphs were captured on: autograd keeps each parameter's gradient accumulation on the stream of
    # its first backward, and would warn about (and synchronise for) a different one
    ev_stream = getattr(trainer, "_capture_stream", None) or torch.cuda.current_stream()
    ev_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ev_stream):
        step_resident(0)
        prof.enable(True)
        launches0 = prof.launch_count()
        ms_eager = timed(args.steps, step_resident)
        launches = prof.launch_count() - launches0  # same kernels per step in the graph replays of the timed region
        kernels = prof.report()
        prof.enable(False)
    torch.cuda.current_stream().wait_stream(ev_stream)
    trainer.use_cuda_graph = graph_mode

    if rank == 0:
        peaks = load_peaks()
        ms_per_step = ms_total / args.steps
        value = world * 1000.0 / ms_per_step
        e2e_value = world * 1000.0 / (ms_e2e / args.steps) if ms_e2e else None
        # dominant kernel of the hot path (largest total device time among this library's kernels)
        def roofline_of(rec):
            """whichever of the tensor pipe and HBM takes longer for the kernel's algorithmic work bounds it"""
            sec = rec["ms"] * 1e-3
            t_tensor = rec["flops"] / (peaks["tflops_sustained"] * 1e12)
            t_hbm = rec["bytes"] / (peaks["hbm_gbs"] * 1e9)
            if t_tensor >= t_hbm:
                ach = rec["flops"] / sec / 1e12
                return {"bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tflops_sustained"]}
            ach = rec["bytes"] / sec / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"]}

        name, rec = max(kernels.items(), key=lambda kv: kv[1]["ms"])
        per_launch_ms = rec["ms"] / rec["launches"]
        roof = {"kernel": name, **roofline_of(rec), "traffic": None,
                "peak_source": peaks["source"] + " (sustained bf16 / copy bandwidth: kernel timed inside a long step)",
                "algorithmic_flops_per_launch": rec["flops"] / rec["launches"],
                "algorithmic_bytes_per_launch": rec["bytes"] / rec["launches"]}
        # measured DRAM traffic of that kernel from the committed ncu --set full capture (per launch, source-image shape)
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")) as f:
                tr = json.load(f)["kernels"].get(name)
            if tr:
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch at the " + tr["shape"] +
                                        " (ncu --set full, profiles/r01_hot_kernels_ncu_full.json); the step mixes source and "
                                        "target shapes, algorithmic_bytes_per_launch is their average; outputs that stay in the "
                                        "126 MB L2 do not show up as DRAM writes")
        except (OSError, ValueError, KeyError):
            pass
        roof["launches_in_timed_region"] = rec["launches"]
        roof["avg_launch_ms"] = per_launch_ms
        hot_ms = sum(r["ms"] for r in kernels.values()) / args.steps
        breakdown = {}
        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]):
            r = roofline_of(v)
            breakdown[k] = {"ms_per_step": round(v["ms"] / args.steps, 4),
                            "launches_per_step": v["launches"] / args.steps, "bound": r["bound"],
                            "achieved": round(r["achieved"], 1), "unit": r["unit"], "frac": round(r["frac"], 3)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": int(src_h.numel() * 4 + lab_h.numel() * 8 + tgt_h.numel() * 4),
                        "d2h_bytes_per_step": 4 * n_losses},
                "gpu_launches": int(launches), "roofline": roof,
                "cuda_graph": bool(graph_mode), "cuda_graph_note": graph_note,
                "eager_ms_per_step_with_kernel_events": ms_eager / args.steps,
                "hot_path_ms_per_step": hot_ms, "hot_path_kernels": breakdown}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(args.cpu_steps, 1, args.level, args.gan)
    # ---- extra (never the headline): the same step with the untouched trunk under bf16 autocast ----
    extra = None
    if args.also_trunk_bf16 and args.trunk_dtype != "bf16":
        try:
            del trainer
            torch.cuda.empty_cache()
            torch.manual_seed(SEED)
            trainer = AdaptSegTrainer(TrainConfig(level=args.level, gan=args.gan, lazy_upsample=args.tier == "B"),
                                      device=dev, use_cuda_graph=bool(args.cuda_graph),
                                      channels_last=bool(args.channels_last), trunk_bf16=True)
            for _ in range(max(args.warmup, 3) + 1):
                step_resident(0)
            ms_bf16 = timed(args.steps, step_resident)
            extra = {"value": world * 1000.0 / (ms_bf16 / args.steps), "unit": UNIT, "ms_per_step": ms_bf16 / args.steps,
                     "note": "ResNet-101 trunk modules under torch.autocast(bfloat16) (SURVEY.md 8f row 1); hot path "
                             "unchanged (fp32 features in, same kernels).  Reported for information: the headline runs the "
                             "trunk in TF32 like the reference's own GPU path"}
        except Exception as exc:  # noqa: BLE001
            extra = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
    if rank == 0:
        if extra is not None:
            line["trunk_bf16_autocast"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
