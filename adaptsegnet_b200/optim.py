"""Flat parameter / gradient storage and the fused optimizer steps on it.

``FlatParams`` moves the parameters of one optimizer into ONE fp32 buffer (each ``nn.Parameter`` becomes a view) with
a gradient buffer of the same layout, so that data-parallel averaging is one NCCL all-reduce and the optimizer step is
one kernel (asn_sgd_step / asn_adam_step) instead of torch.optim's ~10 multi-tensor launches per step.

``FusedSGD`` / ``FusedAdam`` keep the small part of the torch.optim surface the training loop touches
(``param_groups[i]["lr"]``, ``step()``, ``zero_grad()``) and the reference's arithmetic
(train_gta2cityscapes_multi.py:244,347,351-355,532-540,681-683), including the fact that its SGD parameter groups
name most trunk parameters several times (model/deeplab_multi.py:196-218): a sequential optimizer
updates such a parameter once per mention, and so does the kernel.  (torch's CUDA ``foreach`` implementation handles
the duplicated list entries in concurrent chunks of one launch, i.e. with a data race; the fused step is
well-defined and reproduces torch.optim.SGD's sequential path -- ``foreach=False``, what the reference runs on the
CPU here -- first-step buffer initialisation included; tests/test_gpu_optim.py.)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check

_ALIGN = 4  # elements: segments start on 16-byte boundaries so that float4 accesses never straddle two parameters


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams:
    """Parameters (unique, requires_grad) of ``params`` as views into ``values``; their ``.grad`` as views into ``flat``
    (same attribute names as train_step.FlatGrads, which this class extends to the parameters themselves)."""

    def __init__(self, params):
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                uniq.append(p)
        if not uniq:
            raise ValueError("no trainable parameters")
        self.params = uniq
        dev = uniq[0].device
        self.begin, off = [], 0
        for p in uniq:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("flat storage holds fp32 parameters of one device")
            self.begin.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.values = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, b in zip(uniq, self.begin):
                pv, gv = self._view(self.values, p, b), self._view(self.flat, p, b)
                pv.copy_(p.data)
                p.data = pv
                p.grad = gv
        self.index = {id(p): i for i, p in enumerate(uniq)}

    @staticmethod
    def _view(buf, p, begin):
        seg = buf[begin:begin + p.numel()]
        if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
            n, c, h, w = p.shape   # same strides as the channels_last parameter (autograd's gradient layout contract)
            return seg.view(n, h, w, c).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def zero(self):
        self.flat.zero_()

    def mark_updated(self):
        """The kernels write through raw pointers: bump the parameters' autograd version counters by hand, so that
        anything keyed on them (ops.AsppWeightPack / FcdWeightPack re-pack when a version changes) sees the step."""
        try:
            torch._C._increment_version(self.params)
        except TypeError:   # older torch: one tensor per call
            for p in self.params:
                torch._C._increment_version(p)

    def all_reduce_mean(self, group=None):
        """synchronous SUM all-reduce and division by the world size, in place"""
        self.all_reduce_finish(self.all_reduce_start(group), group)
        if self.grad_scale != 1.0:
            self.flat.mul_(self.grad_scale)
            self.grad_scale = 1.0

    def all_reduce_start(self, group=None):
        """asynchronous SUM all-reduce of the flat gradient buffer (None when there is nothing to reduce)"""
        import os
        import torch.distributed as dist

        if os.environ.get("ASN_DIAG_SKIP_ALLREDUCE") == "1":   # diagnosis only: how much of an N > 1 step is the exchange
            return None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return None

    def all_reduce_finish(self, work, group=None):
        """waits for the SUM all-reduce; the division by the world size happens inside the fused optimizer kernel
        (`grad_scale`), not in a separate pass over the buffer"""
        import torch.distributed as dist

        if work is not None:
            work.wait()
            self.grad_scale = 1.0 / dist.get_world_size(group)

    grad_scale = 1.0   # what the next fused step multiplies the gradient by (reset by the step)


class FusedSGD:
    """torch.optim.SGD(param_groups, lr, momentum, weight_decay) on a FlatParams, one kernel per step.

    ``param_groups``: list of {"params": iterable, "lr": float} exactly as handed to torch.optim.SGD; a parameter
    mentioned k times in a group is stepped k times per ``step()`` (sequentially, like a for-loop optimizer)."""

    def __init__(self, flat: FlatParams, param_groups, lr, momentum=0.0, weight_decay=0.0):
        self.flatp = flat
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.param_groups = []
        n = len(flat.params)
        group_of, repeat = [-1] * n, [0] * n
        for gi, g in enumerate(param_groups):
            plist = [p for p in g["params"]]
            self.param_groups.append({"lr": float(g.get("lr", lr)), "params": plist})
            for p in plist:
                i = flat.index.get(id(p))
                if i is None:
                    continue   # not trainable / not in this flat buffer
                if group_of[i] not in (-1, gi):
                    raise ValueError("a parameter appears in two parameter groups")
                group_of[i] = gi
                repeat[i] += 1
        for i in range(n):
            if group_of[i] < 0:      # trainable but not handed to the optimizer: never updated
                group_of[i], repeat[i] = 0, 0
        dev = flat.values.device
        self.seg_begin = torch.tensor(flat.begin + [flat.numel], dtype=torch.int64, device=dev)
        self.seg_group = torch.tensor(group_of, dtype=torch.int32, device=dev)
        self.seg_repeat = torch.tensor(repeat, dtype=torch.int32, device=dev)
        self.repeat = repeat
        self.momentum_buffer = torch.zeros_like(flat.values)
        self.steps = 0

    def step(self):
        f = self.flatp
        lrs = (C.c_float * len(self.param_groups))(*[g["lr"] for g in self.param_groups])
        check(_lib.load().asn_sgd_step(f.values.data_ptr(), f.flat.data_ptr(), self.momentum_buffer.data_ptr(), f.numel,
                                       self.seg_begin.data_ptr(), self.seg_group.data_ptr(), self.seg_repeat.data_ptr(),
                                       len(f.params), lrs, len(self.param_groups), self.momentum, self.weight_decay,
                                       int(self.steps == 0), float(f.grad_scale), _stream()), "asn_sgd_step")
        self.steps += 1
        f.grad_scale = 1.0
        f.mark_updated()

    def zero_grad(self, set_to_none=False):
        self.flatp.zero()


class FusedAdam:
    """torch.optim.Adam(params, lr, betas) on a FlatParams (no weight decay, no amsgrad), one kernel per step."""

    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.flatp = flat
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.param_groups = [{"lr": float(lr), "params": list(flat.params)}]
        self.exp_avg = torch.zeros_like(flat.values)
        self.exp_avg_sq = torch.zeros_like(flat.values)
        self.steps = 0

    def step(self):
        f = self.flatp
        self.steps += 1
        check(_lib.load().asn_adam_step(f.values.data_ptr(), f.flat.data_ptr(), self.exp_avg.data_ptr(),
                                        self.exp_avg_sq.data_ptr(), f.numel, self.param_groups[0]["lr"], self.betas[0],
                                        self.betas[1], self.eps, self.steps, float(f.grad_scale), _stream()),
              "asn_adam_step")
        f.grad_scale = 1.0
        f.mark_updated()

    def zero_grad(self, set_to_none=False):
        self.flatp.zero()
