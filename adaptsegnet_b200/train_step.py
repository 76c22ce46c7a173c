"""One AdaptSegNet training iteration on the new modules.

The reference's loops live inside ``main()`` of train_gta2cityscapes_multi.py and cannot be called
(SURVEY.md Q2); this module restates them operation for operation -- multi-level :560-683,
single-level :373-464 -- on top of the drop-in modules, so the order of forwards, backwards,
``requires_grad`` toggles, loss scalings and optimizer steps is the reference's.  Differences, all
outside the arithmetic: the GAN target is a scalar instead of a CPU-built tensor (Q15), the losses
stay on the device (no ``.item()`` sync per loss, Q16) and, under data parallelism, gradients are
averaged once per optimizer right before ``step()``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import ops
from .model.deeplab_multi import DeeplabMulti
from .model.discriminator import FCDiscriminator
from .utils.loss import GANLoss, SegCrossEntropy

SOURCE_LABEL = 0  # train_gta2cityscapes_multi.py:364-365
TARGET_LABEL = 1


@dataclass
class TrainConfig:
    """defaults of train_gta2cityscapes_multi.py:24-69 (GAN / level as the BASELINE configs name them)"""
    num_classes: int = 19
    learning_rate: float = 2.5e-4
    momentum: float = 0.9
    weight_decay: float = 0.0005
    learning_rate_D: float = 1e-4
    power: float = 0.9
    num_steps: int = 250000
    iter_size: int = 1
    lambda_seg: float = 0.1
    lambda_adv_target1: float = 0.0002
    lambda_adv_target2: float = 0.001
    gan: str = "Vanilla"
    level: str = "multi-level"  # or "single-level"
    fuse_softmax: bool = True   # D(softmax(pred)) with the softmax fused into D's input pack
    reuse_target_forward: bool = True  # D-step on the target reuses the G-step's D(softmax(pred_target)) activations
    # Tier-B: the heads' low-res logits go straight into the fused upsample+softmax+CE kernel and into
    # the discriminators' input pack (upsample + softmax inside); no full-resolution logits / probabilities exist.
    # Same mathematics as the reference's interp -> loss / interp -> softmax -> D chains.  Needs fuse_softmax.
    lazy_upsample: bool = False


def lr_poly(base_lr, it, max_iter, power):
    """train_gta2cityscapes_multi.py:162-163"""
    return base_lr * ((1 - float(it) / max_iter) ** power)


class FlatGrads:
    """All gradients of one optimizer as views into one flat fp32 buffer, so that data-parallel
    averaging is a single NCCL all-reduce per optimizer per iteration."""

    def __init__(self, params):
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        total = sum(p.numel() for p in uniq)
        self.flat = torch.zeros(total, dtype=torch.float32, device=uniq[0].device)
        off = 0
        for p in uniq:
            seg = self.flat[off:off + p.numel()]
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                n, c, h, w = p.shape   # same strides as the channels_last parameter (autograd's layout contract)
                p.grad = seg.view(n, h, w, c).permute(0, 3, 1, 2)
            else:
                p.grad = seg.view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        self.all_reduce_finish(self.all_reduce_start(group), group)

    def all_reduce_start(self, group=None):
        """asynchronous SUM all-reduce of the flat gradient buffer (None when there is nothing to reduce)"""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return None

    def all_reduce_finish(self, work, group=None):
        import torch.distributed as dist

        if work is not None:
            work.wait()
            self.flat.mul_(1.0 / dist.get_world_size(group))


class _FlatView:
    """the slice of a shared FlatParams that belongs to one module (gradient views for tests and tools)"""

    def __init__(self, parent, begin, end):
        self.parent, self.begin, self.end = parent, begin, end

    @property
    def flat(self):
        return self.parent.flat[self.begin:self.end]

    @property
    def values(self):
        return self.parent.values[self.begin:self.end]

    def zero(self):
        self.flat.zero_()


class AdaptSegTrainer:
    """Holds G (DeeplabMulti), D1/D2 (FCDiscriminator), their optimizers and runs iterations."""

    def __init__(self, cfg: TrainConfig | None = None, device="cuda", model=None, model_D1=None, model_D2=None,
                 use_cuda_graph=False, channels_last=False, fused_optimizers=None, trunk_bf16=False, overlap=False):
        """``use_cuda_graph``: capture everything of an iteration up to the gradients (both G forwards/backwards,
        all discriminator passes, ~2 300 kernel launches) into one CUDA graph and replay it; the gradient
        all-reduce and the three optimizer steps stay eager.  Same kernels, same order, no per-launch CPU cost."""
        self.cfg = cfg = cfg or TrainConfig()
        if cfg.iter_size != 1:
            # the reference accumulates iter_size sub-batches per optimizer step (train...:578-679); step() takes exactly
            # one (source, target) pair, so anything else would silently train on 1/iter_size of the gradient
            raise NotImplementedError("AdaptSegTrainer.step() runs one sub-iteration per optimizer step: iter_size must be 1")
        self.device = torch.device(device)
        # fused optimizer steps on flat parameter buffers (optim.py) wherever the CUDA library runs; torch.optim on CPU
        self.fused_optimizers = (self.device.type == "cuda") if fused_optimizers is None else bool(fused_optimizers)
        self.use_cuda_graph = bool(use_cuda_graph)
        # overlap: the source pipeline (forward, seg loss, backward) and the target pipeline (forward, discriminator
        # forward, adversarial backward) of one iteration run on two CUDA streams, and the discriminator step runs beside
        # the target backward (_grads_step_overlap).  Same kernels, same accumulation order, more of the GPU busy.
        self.overlap = bool(overlap)
        self._stream_t = None
        # data parallelism + two-stream schedule: the generator's gradient is all-reduced in three buckets (layer4 + heads,
        # layer3, the rest) that start as soon as the target backward has passed them -- see _bucket_plan / step()
        self._buckets = None
        # opt-in (ASN_BUCKETED_ALLREDUCE=1): bit-identical to the single all-reduce (tools/dp_bucket_check.py, 2 GPUs) but
        # measured neutral -- 34.38 vs 34.34 ms/step at 2 GPUs, 33.76 without any exchange; see _bucket_plan
        self.bucketed_allreduce = os.environ.get("ASN_BUCKETED_ALLREDUCE", "0") == "1"
        if self.overlap:   # gradients produced on `_stream_t` are accumulated on the current stream on purpose (see below)
            try:
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
        self.channels_last = bool(channels_last)
        self._graph = None
        self._capture_stream = None   # the stream warm-up and capture ran on (autograd remembers it per parameter)
        self.multi = cfg.level == "multi-level"
        self.model = (model or DeeplabMulti(cfg.num_classes)).to(self.device).train()
        # single-output trunk + head (DeeplabVGG, BASELINE config 4): forward() returns low-res logits, one level only
        self.single_head = not hasattr(self.model, "layer5")
        if self.single_head and self.multi:
            raise ValueError("a single-head model (DeeplabVGG) trains with level='single-level'")
        if trunk_bf16:   # execution mode of the untouched trunk; the hot path is unaffected
            self.model.trunk_autocast = True
        if self.channels_last:
            # execution detail of the unchanged trunk: cuDNN's sm_100 kernels are NHWC, so an NCHW trunk spends a
            # quarter of its time in layout conversions; the head kernels take the channels_last features as they are
            for name in (("features",) if self.single_head else ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4")):
                getattr(self.model, name).to(memory_format=torch.channels_last)
        self.model_D2 = (model_D2 or FCDiscriminator(cfg.num_classes)).to(self.device).train()
        self.model_D1 = (model_D1 or FCDiscriminator(cfg.num_classes)).to(self.device).train() if self.multi else None
        self.bce_loss = GANLoss(cfg.gan)
        self.seg_loss = SegCrossEntropy(ignore_index=255)
        # optimizers exactly as train...:532-540 (the duplicated trunk parameters included, Q11)
        groups = self.model.optim_parameters(cfg)
        if not (isinstance(groups, list) and groups and isinstance(groups[0], dict)):
            groups = [{"params": list(groups), "lr": cfg.learning_rate}]     # DeeplabVGG: one group (deeplab_vgg.py:53-54)
        if self.fused_optimizers:
            from .optim import FlatParams, FusedAdam, FusedSGD
            self.flat_G = FlatParams(self.model.parameters())
            # both discriminators in ONE flat buffer: the reference's two Adam optimizers (train...:538-540) have the same
            # hyper-parameters, learning-rate schedule and step count, and Adam is element-wise -- one kernel and, under
            # data parallelism, one all-reduce give bit-identical results to two
            d_params = (list(self.model_D1.parameters()) if self.multi else []) + list(self.model_D2.parameters())
            self.flat_D = FlatParams(d_params)
            n1 = sum((p.numel() + 3) // 4 * 4 for p in self.model_D1.parameters()) if self.multi else 0
            self.flat_D1 = _FlatView(self.flat_D, 0, n1) if self.multi else None
            self.flat_D2 = _FlatView(self.flat_D, n1, self.flat_D.numel)
            self.optimizer = FusedSGD(self.flat_G, groups, lr=cfg.learning_rate,
                                      momentum=cfg.momentum, weight_decay=cfg.weight_decay)
            self.optimizer_D = FusedAdam(self.flat_D, lr=cfg.learning_rate_D, betas=(0.9, 0.99))
            self.optimizer_D2 = self.optimizer_D
            self.optimizer_D1 = self.optimizer_D if self.multi else None
        else:
            self.optimizer = torch.optim.SGD(groups, lr=cfg.learning_rate,
                                             momentum=cfg.momentum, weight_decay=cfg.weight_decay)
            self.optimizer_D2 = torch.optim.Adam(self.model_D2.parameters(), lr=cfg.learning_rate_D, betas=(0.9, 0.99))
            self.optimizer_D1 = (torch.optim.Adam(self.model_D1.parameters(), lr=cfg.learning_rate_D, betas=(0.9, 0.99))
                                 if self.multi else None)
            self.flat_G = FlatGrads(self.model.parameters())
            self.flat_D2 = FlatGrads(self.model_D2.parameters())
            self.flat_D1 = FlatGrads(self.model_D1.parameters()) if self.multi else None
            self.flat_D = self.optimizer_D = None
        self.sync_replicas()

    def sync_replicas(self, group=None):
        """Data-parallel replicas must start from identical weights and optimizer state: broadcast rank 0's (flat
        parameter buffers, momentum / Adam moments, BatchNorm statistics) once when a process group is active.  Seeding
        every rank alike already gives that; this makes it hold for a checkpoint loaded on rank 0 only."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        bufs = []
        for fp in (self.flat_G, self.flat_D):
            if fp is not None and hasattr(fp, "values"):
                bufs.append(fp.values)
        if not bufs:   # torch.optim path: parameters are ordinary tensors
            bufs = [p.data for m in (self.model, self.model_D2, self.model_D1) if m is not None for p in m.parameters()]
        for opt in (self.optimizer, self.optimizer_D2, None if self.optimizer_D1 is self.optimizer_D2 else self.optimizer_D1):
            for name in ("momentum_buffer", "exp_avg", "exp_avg_sq"):
                if opt is not None and isinstance(getattr(opt, name, None), torch.Tensor):
                    bufs.append(getattr(opt, name))
        bufs += [b for b in self.model.buffers() if b.is_floating_point()]
        for b in bufs:
            dist.broadcast(b, src=0, group=group)

    # ---- pieces of the loop ------------------------------------------------------------------
    def _adjust_lr(self, i_iter):
        cfg = self.cfg
        lr = lr_poly(cfg.learning_rate, i_iter, cfg.num_steps, cfg.power)
        self.optimizer.param_groups[0]["lr"] = lr
        if len(self.optimizer.param_groups) > 1:                 # train...:166-170
            self.optimizer.param_groups[1]["lr"] = lr * 10
        lr_d = lr_poly(cfg.learning_rate_D, i_iter, cfg.num_steps, cfg.power)
        for opt in (self.optimizer_D1, self.optimizer_D2):
            if opt is not None:
                opt.param_groups[0]["lr"] = lr_d

    def _d_out(self, D, pred, up_size=None):
        if self.cfg.fuse_softmax:
            return D(pred, from_logits=True, up_size=up_size)
        return D(ops.softmax_channels(pred))

    @staticmethod
    def _set_requires_grad(module, flag):
        if module is not None:
            for p in module.parameters():
                p.requires_grad = flag

    def _packs(self):
        heads = [self.model.classifier] if self.single_head else [self.model.layer5, self.model.layer6]
        packs = [h._pack for h in heads] + [self.model_D2._pack]
        if self.multi:
            packs.append(self.model_D1._pack)
        return packs

    def _grads_step(self, src_images, src_labels, tgt_images):
        """Everything of train...:578-679: forwards, losses, backwards; gradients end up in the flat buffers."""
        if self.overlap:
            return self._grads_step_overlap(src_images, src_labels, tgt_images)
        out, carry = self._g_part(src_images, src_labels, tgt_images)
        out.update(self._d_part(carry))
        return out

    def _grads_step_overlap(self, src_images, src_labels, tgt_images):
        """The same iteration as _g_part + _d_part, scheduled on two streams.  What depends on what (train...:578-679):

          source pipeline  S: forward(src) -> seg loss -> backward(src)                      needs the weights only
          target pipeline  T: forward(tgt) -> D(softmax(pred_tgt)) -> adversarial loss -> backward(tgt)
                              forward(tgt) must follow forward(src) (both update the BatchNorm running statistics, in this
                              order); its backward accumulates into the generator's gradients AFTER backward(src)'s
          discriminator    D: D(pred_src.detach()) + backward, replay of T's discriminator forward + backward
                              needs S's and T's predictions and the (unchanged) discriminator weights

        The current stream runs S, then D; stream `_stream_t` forks after forward(src) and runs T.  autograd executes every
        backward node on the stream of its forward op and every parameter's gradient accumulation on the stream the
        parameter was first used on (the current one), inserting the event waits itself -- so backward(tgt)'s convolutions
        run on `_stream_t` beside backward(src) and the discriminator step, while the `+=` into the flat gradient buffers
        stays on one stream, in the reference's order (source first, then target).  Host order matters only in so far as
        every stream executes what it is handed in order: the discriminator step is issued BEFORE backward(tgt), because
        backward(tgt)'s gradient accumulations land on the current stream and would otherwise sit in front of it."""
        cfg = self.cfg
        it = cfg.iter_size
        main = torch.cuda.current_stream()
        if self._stream_t is None:
            self._stream_t = torch.cuda.Stream()
        st = self._stream_t
        self.flat_G.zero()
        if self.flat_D is not None:
            self.flat_D.zero()
        else:
            self.flat_D2.zero()
            if self.multi:
                self.flat_D1.zero()
        out = {}
        self._set_requires_grad(self.model_D1, False)
        self._set_requires_grad(self.model_D2, False)
        if self.channels_last:
            src_images = src_images.contiguous(memory_format=torch.channels_last)
            tgt_images = tgt_images.contiguous(memory_format=torch.channels_last)
        lazy = cfg.lazy_upsample and cfg.fuse_softmax
        up_s = tuple(src_images.shape[-2:]) if lazy else None
        up_t = tuple(tgt_images.shape[-2:]) if lazy else None
        Ds = [(self.model_D2, "D2", cfg.lambda_adv_target2)]
        if self.multi:
            Ds.insert(0, (self.model_D1, "D1", cfg.lambda_adv_target1))
        if ops.precision_mode() == "bf16":
            for D, _, _ in Ds:      # weight packs are shared by both streams: pack them here, before the fork
                D._pack.get(D._params(), cfg.num_classes, D.conv1.weight.shape[0])

        def preds(images):
            if lazy:
                return self.model.low_res_logits(images)
            if self.single_head:
                return None, ops.upsample_bilinear(self.model(images), tuple(images.shape[-2:]))
            return self.model(images)

        # ---- S: forward(src), seg loss ----
        pred1, pred2 = preds(src_images)
        seg = (lambda z: ops.upsample_softmax_cross_entropy(z, up_s, src_labels, ignore_label=255)) if lazy else \
            (lambda z: self.seg_loss(z, src_labels))
        loss_seg2 = seg(pred2)
        loss = loss_seg2
        if self.multi:
            loss_seg1 = seg(pred1)
            loss = loss_seg2 + cfg.lambda_seg * loss_seg1
            out["loss_seg1"] = loss_seg1.detach() / it
        out["loss_seg2"] = loss_seg2.detach() / it
        # ---- T (forked): forward(tgt), discriminator forward, adversarial loss ----
        st.wait_stream(main)
        reuse = cfg.reuse_target_forward and cfg.fuse_softmax and ops.precision_mode() == "bf16"
        saved, adv = {}, {}
        probe_installed = self._install_bucket_probe(main)
        with torch.cuda.stream(st):
            pred_t = dict(zip(("D1", "D2"), preds(tgt_images)))
            if probe_installed:
                self.model.grad_probe = None
            loss_t = 0
            for D, key, lam in Ds:
                if reuse:
                    d, saved[key] = D(pred_t[key], from_logits=True, return_saved=True, up_size=up_t)
                else:
                    d = self._d_out(D, pred_t[key], up_t)
                adv[key] = self.bce_loss(d, SOURCE_LABEL)
                loss_t = lam * adv[key] + loss_t
            loss_t = loss_t / it
            fwd_t_done = torch.cuda.Event()
            fwd_t_done.record(st)
        # ---- S: backward(src) (beside T's forward) ----
        (loss / it).backward()
        # ---- D: the discriminator step on the current stream ----
        self._set_requires_grad(self.model_D1, True)
        self._set_requires_grad(self.model_D2, True)
        pred_s = {"D1": pred1, "D2": pred2}
        l_src = {}
        for D, key, _ in Ds:         # source passes first: they need nothing from T
            l_src[key] = self.bce_loss(self._d_out(D, pred_s[key].detach(), up_s), SOURCE_LABEL) / it / 2
            l_src[key].backward()
        main.wait_event(fwd_t_done)  # T's predictions and discriminator activations exist from here on
        for D, key, _ in Ds:
            d_tgt = D.replay(saved[key]) if saved.get(key) is not None else self._d_out(D, pred_t[key].detach(), up_t)
            l_tgt = self.bce_loss(d_tgt, TARGET_LABEL) / it / 2
            l_tgt.backward()
            out["loss_" + key] = l_src[key].detach() + l_tgt.detach()
        if probe_installed:
            self._buckets["events"]["d_step"].record(main)   # the discriminators' gradients are final from here on
        # ---- T: backward(tgt); its convolutions run on `_stream_t`, the accumulation into the flat buffers on `main` ----
        with torch.cuda.stream(st):
            loss_t.backward()
        main.wait_stream(st)
        for key in adv:
            out["loss_adv_target" + key[1]] = adv[key].detach() / it
        self._keep = (pred1, pred2, pred_t, saved, adv)   # cross-stream tensors stay alive until the next iteration
        return out

    def _bucket_plan(self):
        """[(begin, end, event-or-None)] over flat_G, last layers first: a bucket's gradient is final once the TARGET backward
        (the second and last one into the generator) has passed its layers; the event is recorded at that point."""
        if self._buckets is None:
            fp = self.flat_G
            # Two buckets.  Measured at 2 GPUs (ms/step; no exchange at all: 33.49): layer4 + heads (68 MB) started when the
            # target backward has passed layer4 is hidden completely (33.43); every later bucket (layer3 in one piece or
            # in groups of six blocks, with or without a high-priority NCCL stream) runs beside the last 1-2 ms of the
            # backward pass, where the NCCL kernels take SMs from the critical path and cost as much as they hide
            # (33.81-34.37 against 34.03-34.07 for one all-reduce after the iteration).  So: the early bucket, the
            # discriminators' buffer as soon as the discriminator step is done, and the rest when the iteration ends.
            names = ["layer4.0"]
            mods = dict(self.model.named_modules())
            starts = [fp.begin[fp.index[id(mods[n].conv1.weight)]] for n in names]
            ev = {n: torch.cuda.Event(external=True) for n in names}
            ev["d_step"] = torch.cuda.Event(external=True)
            slices, end = [], fp.numel
            for n, b in zip(names, starts):
                slices.append((b, end, n))
                end = b
            slices.append((0, end, None))     # conv1 .. layer3: final only when the iteration ends
            self._buckets = {"events": ev, "slices": slices, "comm": torch.cuda.Stream()}
        return self._buckets

    def _install_bucket_probe(self, main):
        """for the next trunk forward (the target's): gradient hooks on layer4's and layer3's inputs that record an event on
        the accumulation stream `main` -- every parameter gradient of the layers behind them has been accumulated by then
        (AccumulateGrad nodes run before anything else that is ready)."""
        import torch.distributed as dist

        if not self.bucketed_allreduce or self.single_head or not self.fused_optimizers or \
                not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return False
        plan = self._bucket_plan()

        def probe(name, t):
            if t.requires_grad and name in plan["events"]:
                t.register_hook(lambda g, e=plan["events"][name]: e.record(main))
        self.model.grad_probe = probe
        return True

    def _all_reduce_generator_bucketed(self, group):
        """issued right after the (asynchronous) launch of the iteration: three all-reduces over slices of the flat gradient
        buffer, each behind the event that says its layers' gradients are final, on a side stream -- they overlap the rest of
        the target backward.  -> list of NCCL work handles (the last bucket waits for the end of the iteration)."""
        import torch.distributed as dist

        plan, flat, works = self._bucket_plan(), self.flat_G.flat, []
        comm = plan["comm"]
        diag = os.environ.get("ASN_DIAG_BUCKETS")   # diagnosis only: reduce just the first k buckets
        work_d = None
        for k, (b, e, name) in enumerate(plan["slices"]):
            if diag is not None and k >= int(diag):
                continue
            if name is None:
                works.append(dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM, group=group, async_op=True))
            else:
                with torch.cuda.stream(comm):
                    comm.wait_event(plan["events"][name])
                    works.append(dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM, group=group, async_op=True))
                    if work_d is None and self.flat_D is not None:   # (NCCL runs its collectives in issue order)
                        comm.wait_event(plan["events"]["d_step"])
                        work_d = dist.all_reduce(self.flat_D.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        self._pending_d = work_d
        return works

    def _g_part(self, src_images, src_labels, tgt_images):
        """train...:578-633: the generator's two forward/backward passes.  After it the generator's gradient is final.
        Returns (losses, what the discriminator part needs)."""
        cfg = self.cfg
        it = cfg.iter_size
        self.flat_G.zero()
        if self.flat_D is not None:
            self.flat_D.zero()
        else:
            self.flat_D2.zero()
            if self.multi:
                self.flat_D1.zero()
        out = {}
        # ---------------- train G: discriminators frozen (train...:583-587) ----------------
        self._set_requires_grad(self.model_D1, False)
        self._set_requires_grad(self.model_D2, False)
        if self.channels_last:
            src_images = src_images.contiguous(memory_format=torch.channels_last)
            tgt_images = tgt_images.contiguous(memory_format=torch.channels_last)
        lazy = cfg.lazy_upsample and cfg.fuse_softmax
        up_s = tuple(src_images.shape[-2:]) if lazy else None   # where the consumers upsample to (Tier-B)
        up_t = tuple(tgt_images.shape[-2:]) if lazy else None
        def preds(images):
            if lazy:
                return self.model.low_res_logits(images)
            if self.single_head:   # DeeplabVGG.forward returns low-res logits; the caller interpolates (evaluate...:164-166)
                return None, ops.upsample_bilinear(self.model(images), tuple(images.shape[-2:]))
            return self.model(images)

        pred1, pred2 = preds(src_images)
        if lazy:
            def seg(z):
                return ops.upsample_softmax_cross_entropy(z, up_s, src_labels, ignore_label=255)
        else:
            def seg(z):
                return self.seg_loss(z, src_labels)
        loss_seg2 = seg(pred2)
        if self.multi:
            loss_seg1 = seg(pred1)
            loss = loss_seg2 + cfg.lambda_seg * loss_seg1
            out["loss_seg1"] = loss_seg1.detach() / it
        else:
            loss = loss_seg2
        (loss / it).backward()
        out["loss_seg2"] = loss_seg2.detach() / it

        pred_target1, pred_target2 = preds(tgt_images)
        reuse = cfg.reuse_target_forward and cfg.fuse_softmax and ops.precision_mode() == "bf16"
        saved = {}

        def d_target(D, pred, key):
            if reuse:
                out, saved[key] = D(pred, from_logits=True, return_saved=True, up_size=up_t)
                return out
            return self._d_out(D, pred, up_t)

        loss_adv2 = self.bce_loss(d_target(self.model_D2, pred_target2, "D2"), SOURCE_LABEL)
        loss = cfg.lambda_adv_target2 * loss_adv2
        if self.multi:
            loss_adv1 = self.bce_loss(d_target(self.model_D1, pred_target1, "D1"), SOURCE_LABEL)
            loss = cfg.lambda_adv_target1 * loss_adv1 + loss
            out["loss_adv_target1"] = loss_adv1.detach() / it
        (loss / it).backward()
        out["loss_adv_target2"] = loss_adv2.detach() / it
        det = lambda t: None if t is None else t.detach()  # noqa: E731
        return out, (det(pred1), det(pred2), det(pred_target1), det(pred_target2), saved, up_s, up_t)

    def _d_part(self, carry):
        """train...:635-679: both discriminators on the (detached) source and target predictions."""
        pred1, pred2, pred_target1, pred_target2, saved, up_s, up_t = carry
        it = self.cfg.iter_size
        out = {}
        self._set_requires_grad(self.model_D1, True)
        self._set_requires_grad(self.model_D2, True)
        levels = [(self.model_D2, pred2, pred_target2, "loss_D2", "D2")]
        if self.multi:
            levels.insert(0, (self.model_D1, pred1, pred_target1, "loss_D1", "D1"))
        for D, p_src, p_tgt, name, key in levels:
            l_src = self.bce_loss(self._d_out(D, p_src.detach(), up_s), SOURCE_LABEL) / it / 2
            l_src.backward()
            # train...:665-666 recomputes D(softmax(pred_target)) with unchanged weights and input; the replay
            # re-attaches the G-step's output (and activations) to D's parameters instead
            d_tgt = D.replay(saved[key]) if saved.get(key) is not None else self._d_out(D, p_tgt.detach(), up_t)
            l_tgt = self.bce_loss(d_tgt, TARGET_LABEL) / it / 2
            l_tgt.backward()
            out[name] = l_src.detach() + l_tgt.detach()
        return out

    def _capture(self, src_images, src_labels, tgt_images):
        """warm up on a side stream (cuDNN autotuning, lazy kernel attributes), then capture one iteration as TWO graphs
        sharing a memory pool: the generator part and the discriminator part.  Between their replays step() starts the
        all-reduce of the (by then final) generator gradient, which overlaps the discriminator part."""
        self._static_in = (src_images.clone(), src_labels.clone(), tgt_images.clone())
        bn_state = {k: v.clone() for k, v in self.model.state_dict().items()
                    if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        side = self._capture_stream = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._grads_step(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.model.load_state_dict(bn_state, strict=False)  # the warm-up must not count as training steps
        for pk in self._packs():
            pk.invalidate()                                  # weight packing becomes part of the graph
        # both captures (and the warm-up above) on ONE stream: autograd ties every parameter's gradient accumulation to
        # the stream it first ran on
        self._graph = torch.cuda.CUDAGraph()
        if self.overlap:   # one graph: the two pipelines and the discriminator step are interleaved inside it
            self._graph_d = None
            with torch.cuda.graph(self._graph, stream=side):
                self._static_out = self._grads_step_overlap(*self._static_in)
            self.model.load_state_dict(bn_state, strict=False)
            return
        with torch.cuda.graph(self._graph, stream=side):
            out_g, self._carry = self._g_part(*self._static_in)   # (kept alive: the D graph reads these buffers)
        self._graph_d = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_d, pool=self._graph.pool(), stream=side):
            out_d = self._d_part(self._carry)
        self._static_out = {**out_g, **out_d}
        self.model.load_state_dict(bn_state, strict=False)  # capture does not execute, but keep it explicit

    def step(self, src_images, src_labels, tgt_images, i_iter=0, group=None, do_optimizer_step=True):
        """One iteration.  Returns the dict of (device) loss scalars the reference prints.
        ``do_optimizer_step=False`` leaves the accumulated gradients in place (parity tests)."""
        self._adjust_lr(i_iter)
        if self.use_cuda_graph:
            if self._graph is None:
                self._capture(src_images, src_labels, tgt_images)
            for dst, src in zip(self._static_in, (src_images, src_labels, tgt_images)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._graph.replay()
            if self.overlap and self.model.grad_probe is None and self._buckets is not None and self.bucketed_allreduce:
                pending = self._all_reduce_generator_bucketed(group)
            else:
                pending = self.flat_G.all_reduce_start(group)   # 178 MB over NVLink while the discriminators train
            if self._graph_d is not None:
                self._graph_d.replay()
            out = self._static_out
        elif self.overlap:
            out = self._grads_step_overlap(src_images, src_labels, tgt_images)
            pending = self.flat_G.all_reduce_start(group)
        else:
            out, carry = self._g_part(src_images, src_labels, tgt_images)
            pending = self.flat_G.all_reduce_start(group)
            out.update(self._d_part(carry))
        # ---------------- data-parallel averaging, then the optimizer steps (train...:681-683) ----------------
        if self.flat_D is not None:
            # one all-reduce for both discriminators, started right behind the discriminator part; the generator's (started
            # before it) is waited for first, its fused SGD step then runs while the discriminators' reduce is in flight
            pending_d = getattr(self, "_pending_d", None) if isinstance(pending, list) else None
            self._pending_d = None
            if pending_d is None:
                pending_d = self.flat_D.all_reduce_start(group)
            if isinstance(pending, list):
                import torch.distributed as dist
                for w in pending:
                    w.wait()
                self.flat_G.grad_scale = 1.0 / dist.get_world_size(group)
            else:
                self.flat_G.all_reduce_finish(pending, group)
            if do_optimizer_step:
                self.optimizer.step()
            self.flat_D.all_reduce_finish(pending_d, group)
            if do_optimizer_step:
                self.optimizer_D.step()
            else:   # parity tests read averaged gradients
                for fp in (self.flat_G, self.flat_D):
                    if fp.grad_scale != 1.0:
                        fp.flat.mul_(fp.grad_scale)
                        fp.grad_scale = 1.0
            return out
        self.flat_G.all_reduce_finish(pending, group)
        self.flat_D2.all_reduce_mean(group)
        if self.multi:
            self.flat_D1.all_reduce_mean(group)
        if do_optimizer_step:
            self.optimizer.step()
            if self.multi:
                self.optimizer_D1.step()
            self.optimizer_D2.step()
        return out
