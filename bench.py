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
SEED = 1338


def shapes(args):
    """(source H x W, target H x W): config 2 / 3 shapes, or 512x1024 for both with the VGG variant"""
    return ((512, 1024), (512, 1024)) if args.model == "VGG" else (SRC_HW, TGT_HW)


def metric_name(args):
    if args.model == "VGG":
        return "train iters/s (DeeplabVGG single-level, 512x1024 src+512x1024 tgt)"
    if args.level == "single-level":
        return f"train iters/s (single-level {args.gan} GAN, 720x1280 src+512x1024 tgt)"
    return METRIC


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--level", default="multi-level", choices=["multi-level", "single-level"])
    ap.add_argument("--gan", default="Vanilla", choices=["Vanilla", "LS"])
    ap.add_argument("--model", default="DeepLab", choices=["DeepLab", "VGG"],
                    help="DeepLab: DeeplabMulti / ResNet-101 (BASELINE configs 2, 3); VGG: DeeplabVGG, single-level only, "
                         "512x1024 source and target (BASELINE configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the iteration (up to the gradients) as a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (profiling runs)")
    ap.add_argument("--overlap", type=int, default=1,
                    help="1: source and target pipelines of the iteration on two CUDA streams, discriminator step beside the "
                         "target backward (AdaptSegTrainer overlap=True)")
    ap.add_argument("--channels-last", type=int, default=1, help="run the (unchanged) trunk in channels_last")
    ap.add_argument("--no-kernel-events", action="store_true", help="skip the per-kernel event pass (ncu runs)")
    ap.add_argument("--cudnn-benchmark", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed full-size CPU iterations of the cpu_baseline leg")
    ap.add_argument("--mode", default="train", choices=["train", "eval"],
                    help="train: BASELINE configs[1] (the headline); eval: configs[4], evaluate_cityscapes + fast_hist "
                         "over --frames synthetic 1024x2048 frames")
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the reference-on-B200 (PyTorch eager) comparator block")
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
def cpu_reference_rate(steps, warmup, level, gan, budget_s=240.0, model="DeepLab"):
    """iters/s of the restated reference loop (oracle/torch_ref.RefTrainer, torch CPU fp32, all host threads) MEASURED at
    the full workload -- 720x1280 source + 512x1024 target, the same synthetic batch the GPU arm uses.  `steps` timed
    iterations after `warmup` untimed ones; if the first iterations show that the request would run past `budget_s` of
    wall clock, fewer are run and the counts actually used are returned (never extrapolated)."""
    import torch
    from oracle import torch_ref as TR

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    vgg = model == "VGG"
    tr = TR.RefTrainer(level="single-level" if vgg else level, gan=gan, device="cpu",
                       model=TR.RefDeeplabVGG(19) if vgg else None)
    src_hw, tgt_hw = ((512, 1024), (512, 1024)) if vgg else (SRC_HW, TGT_HW)
    src, lab, tgt = TR.synthetic_batch(SEED, src_hw, tgt_hw)
    t_start = time.perf_counter()
    t0 = time.perf_counter()
    tr.step(src, lab, tgt, i_iter=0)                      # first (cold) iteration: always untimed
    first = time.perf_counter() - t0
    warm_done = 1
    while warm_done < warmup and (time.perf_counter() - t_start) + (steps + 1) * first < budget_s:
        tr.step(src, lab, tgt, i_iter=warm_done)
        warm_done += 1
    left = budget_s - (time.perf_counter() - t_start)
    steps_run = max(1, min(steps, int(left / max(first, 1e-3))))
    t0 = time.perf_counter()
    for i in range(steps_run):
        tr.step(src, lab, tgt, i_iter=warm_done + i)
    dt = (time.perf_counter() - t0) / steps_run
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{steps_run} timed full-size iterations ({src_hw[0]}x{src_hw[1]} source + {tgt_hw[0]}x{tgt_hw[1]} target, after {warm_done} "
                       f"untimed) of the restated reference loop (oracle/torch_ref.RefTrainer, torch {torch.__version__} "
                       f"CPU fp32, {cores} threads): {dt:.3f} s/iter, measured, not extrapolated"),
            "s_per_iter": dt, "steps_timed": steps_run, "warmup_run": warm_done, "extrapolated": False}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    if args.mode == "eval":
        return run_reference_eval(args)
    base = cpu_reference_rate(max(1, args.steps), max(1, args.warmup), args.level, args.gan, model=args.model)
    line = {"impl": "reference", "metric": metric_name(args), "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": base["steps_timed"], "warmup": base["warmup_run"], "steps_requested": args.steps,
            "warmup_requested": args.warmup, "ms_per_step": 1000.0 / base["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_reference(dev, args, steps=5, warmup=3):
    """the reference's own modules and loop (restated, oracle/torch_ref.RefTrainer) on the SAME
    B200 in PyTorch eager -- NCHW fp32 tensors, cuDNN benchmark mode as train_gta2cityscapes_multi.py:228 sets it, .item()
    per loss, CPU-built GAN targets -- with TF32 on (torch's cuDNN default, what the reference gets on this GPU) and off
    (true fp32).  Whole step in ms, plus the hot-path groups through ATen next to libasn_b200's (bench_groups.py)."""
    import torch
    from oracle import torch_ref as TR
    import bench_groups

    out = {}
    src, lab, tgt = TR.synthetic_batch(SEED, *shapes(args))
    src, lab, tgt = src.to(dev), lab.to(dev), tgt.to(dev)
    vgg = args.model == "VGG"
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.manual_seed(SEED)
            tr = TR.RefTrainer(level="single-level" if vgg else args.level, gan=args.gan, device=dev,
                               model=TR.RefDeeplabVGG(19) if vgg else None)
            for i in range(warmup):
                tr.step(src, lab, tgt, i_iter=i)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                tr.step(src, lab, tgt, i_iter=warmup + i)
            e1.record()
            torch.cuda.synchronize(dev)
            out[f"step_ms_{tag}"] = e0.elapsed_time(e1) / steps
            del tr
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["steps"], out["warmup"] = steps, warmup
    out["what"] = ("restated reference loop (train_gta2cityscapes_multi.py:560-683) over the restated reference modules, "
                   "PyTorch eager on this GPU, NCHW fp32, cudnn.benchmark=True; CUDA events")
    if args.level == "multi-level" and not vgg:
        out["hot_path_groups"] = bench_groups.compare(dev, tier=args.tier)
        out["hot_path_groups_what"] = ("one iteration's hot path at config-2 shapes on synthetic features / logits, per group, "
                                       "median of 5 after 2 warm-ups, L2 flushed between repetitions: heads = 2 ASPP heads x "
                                       "(source, target) fwd+bwd; seg_loss = upsample+CE fwd+bwd (source, 2 heads); adversarial = "
                                       "all discriminator passes of both levels incl. target upsample+softmax and the GAN "
                                       "losses; optimizers = SGD + 2 x Adam.  ours_ms and aten_*_ms: both launched eagerly from "
                                       "Python, kernel by kernel (the speed-ups are computed from these); ours_graph_ms: the same "
                                       "group of libasn_b200 launches replayed as one CUDA graph, which is how the product runs them "
                                       "(AdaptSegTrainer captures the iteration) -- the device time without the host launch path")
    return out


# ------------------------------------------------------------------------------------------------------
# --mode eval: BASELINE configs[4] -- evaluate_cityscapes inference at 1024x2048 + fast_hist 19-class IoU
# ------------------------------------------------------------------------------------------------------
EVAL_METRIC = "eval frames/s (512x1024 image -> 1024x2048 argmax + fast_hist 19-class IoU)"
EVAL_IMG_HW, EVAL_LABEL_HW = (512, 1024), (1024, 2048)
# Cityscapes id -> train id (the absent dataset/cityscapes_list/info.json `label2train`, compute_iou.py:40): the public
# Cityscapes label definition; every id without a train id maps to 255
CITYSCAPES_LABEL2TRAIN = [[i, 255] for i in range(34)] + [[-1, 255]]
for _i, _t in zip((7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33), range(19)):
    CITYSCAPES_LABEL2TRAIN[_i] = [_i, _t]


def eval_frame(seed):
    """one synthetic validation frame: mean-subtracted BGR-like image (1,3,512,1024) fp32 and a blocky raw-id label map
    (1024,2048) uint8 with Cityscapes ids 0..33 (per-frame seed)"""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([104.00698793, 116.66876762, 122.67891434]).view(1, 3, 1, 1)
    img = torch.randint(0, 256, (1, 3) + EVAL_IMG_HW, generator=g).float() - mean
    coarse = torch.randint(0, 34, (1, 1, 32, 64), generator=g).float()
    lab = F.interpolate(coarse, size=EVAL_LABEL_HW, mode="nearest")[0].to(torch.uint8)
    return img, lab


def eval_workload_config(args, world, impl):
    cfg = {"workload": f"evaluate_cityscapes + compute_iou over {args.frames} synthetic frames: DeeplabMulti(ResNet-101, 19 cls) "
                       "forward on 1x3x512x1024, output2 -> 512x1024 -> 1024x2048 bilinear, argmax, label_mapping, "
                       "19x19 fast_hist, mIoU; random init; 8 distinct seeded frames cycled",
           "frames": args.frames, "per_gpu_batch": "1 frame"}
    if impl == "reference":
        cfg.update({"parallelism": "none (rank 0 only, host CPU)",
                    "hot_path": "the reference's torch CPU + numpy ops (evaluate_cityscapes.py:155-169, compute_iou.py:50-64 restated)"})
    else:
        cfg.update({"parallelism": f"dp{world}: frames sharded round-robin, one int64 all-reduce of the 19x19 matrix at the end",
                    "hot_path": "libasn_b200: tcgen05 layer6 head + ONE fused kernel for both bilinear stages, argmax, "
                                "label_mapping LUT and the confusion matrix; device mIoU",
                    "execution": "one CUDA graph per frame; trunk = unchanged PyTorch modules (cuDNN TF32, channels_last), eval mode",
                    "dead_work": "layer5 (output1) is not computed: the reference computes it and never uses it "
                                 "(evaluate_cityscapes.py:162-163)",
                    "l2": "8 distinct frames x (6.3 MB image + 2.1 MB label) cycled, trunk activations >> 126 MB L2"})
    return cfg


def cpu_reference_eval(frames, preds_by_seed=None, budget_s=60.0):
    """frames/s of the reference's CPU path per frame (evaluate_cityscapes.py:155-169 + compute_iou.py:53-57 restated with
    the reference's own torch / numpy calls), measured on a bounded number of frames.  If ``preds_by_seed`` (GPU
    predictions of the same frames) is given, also returns numpy's confusion matrix over THOSE predictions -- the checker
    for the GPU matrix (bit-exact by contract)."""
    import numpy as np
    import torch
    from oracle import np_oracle as O
    from oracle import torch_ref as TR

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    model = TR.RefDeeplabMulti(19).eval()
    mapping = np.array(CITYSCAPES_LABEL2TRAIN, dtype=np.int64)
    parts = {"forward": 0.0, "interp_d2h": 0.0, "argmax": 0.0, "label_mapping": 0.0, "fast_hist": 0.0}
    hist = np.zeros((19, 19))
    t_start = time.perf_counter()
    done = 0
    for f in range(frames + 1):                      # frame 0 is the untimed warm-up
        img, lab = eval_frame(SEED + (f % 8))
        t = [time.perf_counter()]
        with torch.no_grad():
            _, out2 = model(img, (img.shape[3], img.shape[2]))                                    # evaluate...:162
            t.append(time.perf_counter())
            interp = torch.nn.Upsample(size=EVAL_LABEL_HW, mode="bilinear", align_corners=True)    # :153
            output = interp(out2).cpu().data[0].numpy()                                            # :163
        t.append(time.perf_counter())
        pred = np.asarray(np.argmax(output.transpose(1, 2, 0), axis=2), dtype=np.uint8)            # :168-169
        t.append(time.perf_counter())
        label = O.label_mapping(lab.numpy(), mapping)                                              # compute_iou.py:55
        t.append(time.perf_counter())
        h = O.fast_hist(label.flatten(), pred.flatten(), 19)                                       # :57
        t.append(time.perf_counter())
        if f > 0:
            hist += h
            for k, a, b in zip(parts, t[:-1], t[1:]):
                parts[k] += b - a
            done += 1
        if f > 0 and time.perf_counter() - t_start > budget_s:
            break
    total = sum(parts.values())
    res = {"value": done / total, "unit": "frames/s", "cores": cores, "kind": "port", "frames_timed": done,
           "sample": (f"{done} frames (after 1 untimed) of the restated evaluate_cityscapes + compute_iou loop, torch "
                      f"{torch.__version__} CPU fp32 + numpy, {cores} threads: {total / done:.3f} s/frame, measured"),
           "s_per_frame_parts": {k: round(v / done, 4) for k, v in parts.items()}}
    if preds_by_seed is not None:
        chk = np.zeros((19, 19), dtype=np.int64)
        for seed, pred in preds_by_seed.items():
            _, lab = eval_frame(seed)
            chk += O.fast_hist(O.label_mapping(lab.numpy(), mapping).flatten(), pred.flatten(), 19)
        res["_check_hist"] = chk
    return res


def run_reference_eval(args):
    base = cpu_reference_eval(max(1, args.steps), budget_s=240.0)
    base.pop("_check_hist", None)
    line = {"impl": "reference", "metric": EVAL_METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": base["frames_timed"], "warmup": 1, "ms_per_step": 1000.0 / base["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": eval_workload_config(args, 1, "reference"), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_eval(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from adaptsegnet_b200 import ops, prof
    from adaptsegnet_b200.evaluate import Evaluator
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    torch.manual_seed(SEED)
    model = DeeplabMulti(19).to(dev).eval()
    pool_h = [eval_frame(SEED + i) for i in range(8)]
    pool_h = [(i.pin_memory(), l.pin_memory()) for i, l in pool_h]
    pool_d = [(i.to(dev), l.to(dev)) for i, l in pool_h]
    my_frames = [f for f in range(args.frames) if f % world == rank]   # round-robin shard

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ev = Evaluator(model, 19, EVAL_LABEL_HW, mapping=CITYSCAPES_LABEL2TRAIN, use_cuda_graph=bool(args.cuda_graph),
                   channels_last=bool(args.channels_last))
    for f in range(max(args.warmup, 3)):
        ev.step(*pool_d[f % 8])
    ev.hist.zero_()

    def run_resident():
        for f in my_frames:
            ev.step(*pool_d[f % 8])
        ev.all_reduce()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = prof.launch_count()
    ms_total = timed(run_resident)
    launches = prof.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    hist_all = ev.hist.clone()
    iu, miou = ev.result()
    # ---- e2e: image + label from pinned host memory every frame, the uint8 prediction back to the host ----
    ev2 = Evaluator(model, 19, EVAL_LABEL_HW, mapping=CITYSCAPES_LABEL2TRAIN, use_cuda_graph=bool(args.cuda_graph),
                    channels_last=bool(args.channels_last), keep_pred=True)
    img_d, lab_d = pool_d[0][0].clone(), pool_d[0][1].clone()
    pred_h = torch.empty((1,) + EVAL_LABEL_HW, dtype=torch.uint8).pin_memory()
    for f in range(3):
        ev2.step(img_d, lab_d)
    ev2.hist.zero_()

    def run_e2e():
        for f in my_frames:
            img_d.copy_(pool_h[f % 8][0], non_blocking=True)
            lab_d.copy_(pool_h[f % 8][1], non_blocking=True)
            pred = ev2.step(img_d, lab_d)
            pred_h.copy_(pred, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the caller consumes (saves) every prediction
        ev2.all_reduce()

    ms_e2e = None if args.no_e2e else timed(run_e2e)
    # ---- per-kernel events over a few eager frames (events cannot be recorded inside a captured graph) ----
    kernels = {}
    launches_per_frame = None
    if not args.no_kernel_events:
        ev3 = Evaluator(model, 19, EVAL_LABEL_HW, mapping=CITYSCAPES_LABEL2TRAIN, use_cuda_graph=False,
                        channels_last=bool(args.channels_last))
        for f in range(3):
            ev3.step(*pool_d[f % 8])
        prof.enable(True)
        n_ev = 16
        l0 = prof.launch_count()
        for f in range(n_ev):
            ev3.step(*pool_d[f % 8])
        torch.cuda.synchronize()
        launches_per_frame = (prof.launch_count() - l0) / n_ev   # the graph replays of the timed region launch the same kernels
        kernels = prof.report()
        prof.enable(False)
    if rank == 0:
        peaks = load_peaks()
        n = args.frames
        line = {"metric": EVAL_METRIC, "value": n * 1000.0 / ms_total, "unit": "frames/s", "n_gpus": world, "steps": n,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / len(my_frames), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": eval_workload_config(args, world, "b200"), "clocks": clocks,
                "e2e": {"value": (n * 1000.0 / ms_e2e) if ms_e2e else None, "unit": "frames/s",
                        "h2d_bytes_per_step": int(pool_h[0][0].numel() * 4 + pool_h[0][1].numel()),
                        "d2h_bytes_per_step": int(pred_h.numel())},
                # kernels of this library per frame (counted in the eager pass; the captured frames replay the same ones)
                "gpu_launches": int(round(launches_per_frame * n)) if launches_per_frame else int(launches) * world,
                "gpu_launches_per_frame": launches_per_frame, "miou": float(miou.item()),
                "e2e_hist_equals_resident": bool(torch.equal(ev2.hist, hist_all)) if ms_e2e else None}
        if kernels:
            def frac_of(rec):
                sec = rec["ms"] * 1e-3
                t_tensor = rec["flops"] / (peaks["tflops_sustained"] * 1e12)
                t_hbm = rec["bytes"] / (peaks["hbm_gbs"] * 1e9)
                if t_tensor >= t_hbm:
                    a = rec["flops"] / sec / 1e12
                    return {"bound": "tensor", "achieved": a, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                            "frac": a / peaks["tflops_sustained"]}
                a = rec["bytes"] / sec / 1e9
                return {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"]}
            name, rec = max(kernels.items(), key=lambda kv: kv[1]["ms"])
            line["roofline"] = {"kernel": name, **frac_of(rec), "traffic": None, "avg_launch_ms": rec["ms"] / rec["launches"],
                                "algorithmic_bytes_per_launch": rec["bytes"] / rec["launches"],
                                "algorithmic_flops_per_launch": rec["flops"] / rec["launches"],
                                "peak_source": peaks["source"]}
            line["hot_path_kernels"] = {k: {"ms_per_frame": round(v["ms"] / 16, 4), "launches_per_frame": v["launches"] / 16,
                                            **{kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in frac_of(v).items()
                                               if kk in ("bound", "achieved", "unit", "frac")}}
                                        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}
            line["hot_path_ms_per_frame"] = sum(v["ms"] for v in kernels.values()) / 16
        if world == 1 and not args.no_cpu_baseline:
            # GPU predictions of the 8 pool frames -> numpy fast_hist over them on the host = the checker
            ev4 = Evaluator(model, 19, EVAL_LABEL_HW, mapping=CITYSCAPES_LABEL2TRAIN, use_cuda_graph=False,
                            channels_last=bool(args.channels_last), keep_pred=True)
            preds = {}
            for i in range(8):
                preds[SEED + i] = ev4.step(*pool_d[i]).cpu().numpy()[0].copy()
            base = cpu_reference_eval(args.cpu_steps, preds_by_seed=preds, budget_s=45.0)
            chk = base.pop("_check_hist")
            line["cpu_baseline"] = base
            line["hist_bit_exact_vs_numpy"] = bool(np.array_equal(ev4.hist.cpu().numpy(), chk))
            # the 500-frame matrix is the pool's matrix taken frames/8 times (8 distinct frames cycled)
            reps = np.array([len([f for f in range(n) if f % 8 == i]) for i in range(8)])
            per_frame = []
            for i in range(8):
                from adaptsegnet_b200.compute_iou import fast_hist as gpu_hist
                per_frame.append(gpu_hist(torch.from_numpy(eval_frame(SEED + i)[1].numpy().reshape(-1)).to(dev),
                                          torch.from_numpy(preds[SEED + i].reshape(-1)).to(dev), 19,
                                          mapping=CITYSCAPES_LABEL2TRAIN).cpu().numpy())
            want_all = sum(int(r) * h for r, h in zip(reps, per_frame))
            line["hist_500_frames_consistent"] = bool(np.array_equal(hist_all.cpu().numpy(), want_all))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_config(args, world):
    (sh, sw), (th, tw) = shapes(args)
    level = "single-level" if args.model == "VGG" else args.level
    net = "DeeplabVGG(VGG-16, 19 cls, 2-branch head)" if args.model == "VGG" else "DeeplabMulti(ResNet-101, 19 cls)"
    cfg = {"workload": f"{level} AdaptSegNet train step, {net} + "
                       f"{'2x' if level == 'multi-level' else '1x'} FCDiscriminator, {args.gan} GAN, "
                       f"src 1x3x{sh}x{sw} + tgt 1x3x{th}x{tw} per GPU, random init",
           "per_gpu_batch": "1 source + 1 target image", "global_pairs_per_step": world}
    if args.impl == "reference":
        cfg.update({"parallelism": "none (rank 0 only, host CPU)",
                    "hot_path": "the reference's own torch CPU ops (restated loop, oracle/torch_ref.RefTrainer)",
                    "execution": "eager, all host threads", "trunk": "ResNet-101 on torch CPU (oneDNN), fp32"})
        return cfg
    cfg.update({"parallelism": f"dp{world} (NCCL all-reduce of 3 flat gradient buffers per step)",
                "hot_path": "libasn_b200 sm_100a kernels (tcgen05 heads + discriminators, fused losses)",
                "execution": (("forward/backward of the iteration replayed as ONE CUDA graph with two streams inside: the source "
                               "pipeline and the discriminator step on one, the target pipeline on the other; fused optimizer "
                               "steps eager" if args.overlap else
                               "forward/backward of the iteration replayed as two CUDA graphs (generator part, discriminator "
                               "part; the generator's gradient all-reduce overlaps the second); fused optimizer steps eager")
                              if args.cuda_graph else "eager"),
                "trunk": (("VGG-16 features" if args.model == "VGG" else "ResNet-101") + " as PyTorch modules on cuDNN ("
                          + ("bf16 autocast" if args.trunk_dtype == "bf16" else "TF32")
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
        opts = dist.ProcessGroupNCCL.Options()     # (a high-priority NCCL stream measured neutral: 34.31 vs 34.37 ms at 2 GPUs)
        opts.is_high_priority_stream = os.environ.get("ASN_NCCL_HIGH_PRIORITY", "0") == "1"
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)  # train_gta2cityscapes_multi.py:228

    torch.manual_seed(SEED)  # identical replicas
    if args.model == "VGG":
        args.level = "single-level"

    def make_model():
        if args.model == "VGG":
            from adaptsegnet_b200.model.deeplab_vgg import DeeplabVGG
            return DeeplabVGG(19)
        return None

    trainer = AdaptSegTrainer(TrainConfig(level=args.level, gan=args.gan, lazy_upsample=args.tier == "B"), device=dev,
                              model=make_model(), use_cuda_graph=bool(args.cuda_graph),
                              channels_last=bool(args.channels_last), trunk_bf16=args.trunk_dtype == "bf16",
                              overlap=bool(args.overlap))
    src_h, lab_h, tgt_h = synthetic_batch(SEED + rank, *shapes(args))  # each rank its own pair
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

    # e2e input pipeline: what a DataLoader with pinned memory and one batch of prefetch does -- the NEXT step's images and
    # labels travel host -> device on a copy stream into a staging set while the current step computes; the step itself
    # starts with a device-to-device copy of its (already resident) staging set into the graph's static input buffers.
    # Every step's 24.7 MB still cross PCIe inside the timed region, the losses still come back and are waited for.
    copy_stream = torch.cuda.Stream()
    staging = [tuple(torch.empty_like(t) for t in (src, lab, tgt)) for _ in range(2)]
    staged = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"primed": False}

    def prefetch(k):
        copy_stream.wait_stream(torch.cuda.current_stream())   # the staging set's previous reader is done
        with torch.cuda.stream(copy_stream):
            for d, h in zip(staging[k], (src_h, lab_h, tgt_h)):
                d.copy_(h, non_blocking=True)
            staged[k].record(copy_stream)

    def step_e2e(i):
        cur = it[0] & 1
        if not e2e_state["primed"]:          # the very first step has nothing prefetched: its copy is exposed
            prefetch(cur)
            e2e_state["primed"] = True
        torch.cuda.current_stream().wait_event(staged[cur])
        if trainer.use_cuda_graph and trainer._graph is not None:
            ins = trainer._static_in
            for d, s_ in zip(ins, staging[cur]):
                d.copy_(s_, non_blocking=True)
        else:
            ins = staging[cur]
        prefetch(cur ^ 1)                    # next step's inputs: host -> device beside this step's compute
        out = trainer.step(*ins, i_iter=it[0])
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
    overlap_mode, trainer.overlap = trainer.overlap, False   # per-kernel times: one kernel at a time (sequential schedule)
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
    trainer.overlap = overlap_mode

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
        # measured DRAM traffic of that kernel from the committed ncu --set full capture (per launch; newest round first)
        for tag in ("r02", "r01"):
            try:
                with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")) as f:
                    tr = json.load(f)["kernels"].get(name)
            except (OSError, ValueError, KeyError):
                tr = None
            if tr:
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch, " + tr["shape"] +
                                        f" (ncu --set full, profiles/{tag}_hot_kernels_ncu_full.json); outputs that stay in the "
                                        "126 MB L2 do not show up as DRAM writes")
                break
        roof["launches_in_timed_region"] = rec["launches"]
        roof["avg_launch_ms"] = per_launch_ms
        hot_ms = sum(r["ms"] for r in kernels.values()) / args.steps
        breakdown = {}
        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]):
            r = roofline_of(v)
            breakdown[k] = {"ms_per_step": round(v["ms"] / args.steps, 4),
                            "launches_per_step": v["launches"] / args.steps, "bound": r["bound"],
                            "achieved": round(r["achieved"], 1), "unit": r["unit"], "frac": round(r["frac"], 3)}
        line = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": int(src_h.numel() * 4 + lab_h.numel() * 8 + tgt_h.numel() * 4),
                        "d2h_bytes_per_step": 4 * n_losses},
                "gpu_launches": int(launches) * world, "gpu_launches_per_rank": int(launches), "roofline": roof,
                "cuda_graph": bool(graph_mode), "cuda_graph_note": graph_note,
                "eager_ms_per_step_with_kernel_events": ms_eager / args.steps,
                "kernel_events_note": "per-kernel CUDA events are taken in the sequential single-stream eager schedule (one "
                                      "kernel at a time); the headline runs the same kernels as a CUDA graph"
                                      + (" with the two-stream schedule" if overlap_mode else ""),
                "hot_path_ms_per_step": hot_ms, "hot_path_kernels": breakdown}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(args.cpu_steps, 1, args.level, args.gan, budget_s=60.0, model=args.model)
    # ---- the kernel-for-kernel bar: the reference's own modules on this GPU in PyTorch eager (rank 0, N = 1) ----
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        try:
            del trainer
            trainer = None
            torch.cuda.empty_cache()
            line["reference_gpu_eager"] = gpu_eager_reference(dev, args)
            ref_ms = line["reference_gpu_eager"]["step_ms_tf32"]
            line["reference_gpu_eager"]["speedup_step_vs_tf32"] = ref_ms / ms_per_step
        except Exception as exc:  # noqa: BLE001
            line["reference_gpu_eager"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    # ---- extra (never the headline): the same step with the untouched trunk under bf16 autocast ----
    extra = None
    if args.also_trunk_bf16 and args.trunk_dtype != "bf16":
        try:
            trainer = None
            torch.cuda.empty_cache()
            torch.manual_seed(SEED)
            trainer = AdaptSegTrainer(TrainConfig(level=args.level, gan=args.gan, lazy_upsample=args.tier == "B"),
                                      device=dev, model=make_model(), use_cuda_graph=bool(args.cuda_graph),
                                      channels_last=bool(args.channels_last), trunk_bf16=True, overlap=bool(args.overlap))
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
    elif args.mode == "eval":
        run_eval(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
