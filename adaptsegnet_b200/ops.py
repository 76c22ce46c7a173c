"""torch-facing operators over the C ABI (include/asn_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every
arithmetic op below is a call into libasn_b200.so on the current CUDA stream.  CPU tensors
are rejected -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import _lib
from ._lib import check

GAN_BCE, GAN_MSE = 0, 1
_LABEL_CODE = {torch.uint8: 0, torch.int32: 1, torch.int64: 2}

# kernels launched through this module (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype=None, name: str = "tensor") -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.AsnError(f"{name} must be a CUDA tensor: adaptsegnet_b200 has no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _is_channels_last(t: torch.Tensor) -> bool:
    """a CUDA fp32 (N,C,H,W) tensor stored NHWC (and not also NCHW-contiguous, i.e. C > 1 and H*W > 1)"""
    return (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 4 and t.dtype == torch.float32
            and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)
            and t.shape[1] % 4 == 0)


def precision_mode() -> str:
    """'bf16' (tcgen05 tensor cores, default) or 'fp32' (CUDA-core FFMA, 1e-4 mode)."""
    m = os.environ.get("ASN_PRECISION", "bf16").lower()
    if m not in ("bf16", "fp32"):
        raise ValueError("ASN_PRECISION must be bf16 or fp32")
    return m


# --------------------------------------------------------------------------------------
# K7 fast_hist
# --------------------------------------------------------------------------------------
def mapping_lut(mapping, device) -> torch.Tensor:
    """256-entry uint8 LUT of compute_iou.py:24-28 (label_mapping): lut[src] = dst for every (src, dst) row of
    ``mapping`` (later rows win, as the reference's successive passes over the ORIGINAL labels do), identity elsewhere."""
    lut = list(range(256))
    for src, dst in mapping:
        src, dst = int(src), int(dst)
        if 0 <= src < 256:
            if not 0 <= dst < 256:
                raise ValueError(f"label_mapping target {dst} does not fit the 8-bit LUT")
            lut[src] = dst
        # sources outside [0,256) cannot occur in 8-bit label PNGs; wider labels with such ids stay unmapped
    return torch.tensor(lut, dtype=torch.uint8, device=device)


def fast_hist(label: torch.Tensor, pred: torch.Tensor, n_cls: int, hist: torch.Tensor | None = None,
              lut: torch.Tensor | None = None):
    """compute_iou.py:15-17 on the device.  Returns (hist int64 [n,n], overflow int64 [1]).
    ``lut`` (uint8[256], see mapping_lut): label_mapping (compute_iou.py:24-28) fused into the counting."""
    label = _req(label, None, "label")
    pred = _req(pred, torch.uint8, "pred")
    if label.dtype not in _LABEL_CODE:
        raise TypeError(f"label dtype {label.dtype} not supported (uint8, int32, int64)")
    if label.numel() != pred.numel():
        raise ValueError("label and pred must have the same number of pixels")
    if hist is None:
        hist = torch.zeros((n_cls, n_cls), dtype=torch.int64, device=label.device)
    overflow = torch.zeros(1, dtype=torch.int64, device=label.device)
    lib = _lib.load()
    if lut is not None:
        lut = _req(lut, torch.uint8, "lut")
        if lut.numel() != 256:
            raise ValueError("lut must have 256 entries")
        check(lib.asn_fast_hist_lut(label.data_ptr(), _LABEL_CODE[label.dtype], lut.data_ptr(), pred.data_ptr(),
                                    label.numel(), n_cls, hist.data_ptr(), overflow.data_ptr(), _stream()),
              "asn_fast_hist_lut")
    else:
        check(lib.asn_fast_hist(label.data_ptr(), _LABEL_CODE[label.dtype], pred.data_ptr(), label.numel(), n_cls,
                                hist.data_ptr(), overflow.data_ptr(), _stream()), "asn_fast_hist")
    _count()
    return hist, overflow


# --------------------------------------------------------------------------------------
# input pipeline: dataset/gta5_dataset.py:58-71 on the raw 8-bit buffers
# --------------------------------------------------------------------------------------
GTA5_ID_TO_TRAINID = {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11, 25: 12, 26: 13,
                      27: 14, 28: 15, 31: 16, 32: 17, 33: 18}   # dataset/gta5_dataset.py:28-30


def image_to_tensor(rgb_u8: torch.Tensor, mean_bgr) -> torch.Tensor:
    """(N, H, W, 3) uint8 RGB on the device -> (N, 3, H, W) fp32, BGR, mean-subtracted: `image[:, :, ::-1] - mean` then
    `transpose((2, 0, 1))` of dataset/gta5_dataset.py:66-69, bit-exact (8-bit values and fp32 means subtract exactly)."""
    rgb_u8 = _req(rgb_u8, torch.uint8, "image")
    N, H, W, c3 = rgb_u8.shape
    if c3 != 3:
        raise ValueError("image must be (N, H, W, 3) uint8")
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=rgb_u8.device)
    check(_lib.load().asn_image_u8_to_bgr_f32(rgb_u8.data_ptr(), out.data_ptr(), N, H, W, float(mean_bgr[0]),
                                              float(mean_bgr[1]), float(mean_bgr[2]), _stream()), "asn_image_u8_to_bgr_f32")
    _count()
    return out


def label_to_trainid(ids_u8: torch.Tensor, id_to_trainid=None, ignore_label=255) -> torch.Tensor:
    """uint8 label ids -> int64 train ids (ignore_label where the table has no entry): dataset/gta5_dataset.py:61-64 + the
    `.long()` of train...:595 in one pass."""
    ids_u8 = _req(ids_u8, torch.uint8, "label")
    table = GTA5_ID_TO_TRAINID if id_to_trainid is None else id_to_trainid
    lut = torch.full((256,), int(ignore_label), dtype=torch.uint8)
    for k, v in table.items():
        lut[int(k)] = int(v)
    lut = lut.to(ids_u8.device)
    out = torch.empty(ids_u8.shape, dtype=torch.int64, device=ids_u8.device)
    check(_lib.load().asn_label_u8_to_trainid_i64(ids_u8.data_ptr(), lut.data_ptr(), out.data_ptr(), ids_u8.numel(), _stream()),
          "asn_label_u8_to_trainid_i64")
    _count()
    return out


# --------------------------------------------------------------------------------------
# K2 upsample
# --------------------------------------------------------------------------------------
def upsample_fwd_raw(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    x = _req(x, torch.float32, "x")
    N, Cc, h, w = x.shape
    y = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    check(_lib.load().asn_upsample_bilinear_fwd(x.data_ptr(), N, Cc, h, w, y.data_ptr(), H, W, _stream()),
          "asn_upsample_bilinear_fwd")
    _count()
    return y


def upsample_bwd_raw(dy: torch.Tensor, h: int, w: int) -> torch.Tensor:
    dy = _req(dy, torch.float32, "dy")
    N, Cc, H, W = dy.shape
    lib = _lib.load()
    dx = torch.empty((N, Cc, h, w), dtype=torch.float32, device=dy.device)
    nbytes = lib.asn_upsample_bwd_workspace_bytes(N, Cc, H, W, h, w)
    ws = _ws(nbytes, dy.device)
    check(lib.asn_upsample_bilinear_bwd(dy.data_ptr(), N, Cc, H, W, dx.data_ptr(), h, w, ws.data_ptr(), nbytes,
                                        _stream()), "asn_upsample_bilinear_bwd")
    _count(2)
    return dx


class _Upsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        ctx.hw = (x.shape[2], x.shape[3])
        return upsample_fwd_raw(x, H, W)

    @staticmethod
    def backward(ctx, dy):
        return upsample_bwd_raw(dy, *ctx.hw), None, None


def upsample_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """nn.Upsample(size=size, mode='bilinear', align_corners=True)(x) -- model/deeplab_multi.py:188-189."""
    return _Upsample.apply(x, int(size[0]), int(size[1]))


def upsample_argmax(x: torch.Tensor, size) -> torch.Tensor:
    """evaluate_cityscapes.py:153,163,168-169 fused: uint8 [N,H,W] class ids, first max wins."""
    x = _req(x, torch.float32, "x")
    N, Cc, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    pred = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    check(_lib.load().asn_upsample_argmax_u8(x.data_ptr(), N, Cc, h, w, pred.data_ptr(), H, W, _stream()),
          "asn_upsample_argmax_u8")
    _count()
    return pred


def upsample2_argmax(x: torch.Tensor, mid_size, size) -> torch.Tensor:
    """argmax_c interp_size(interp_mid(x)) as uint8 [N,H,W]: the fork's evaluation chain with BOTH of its bilinear stages
    (model/deeplab_multi.py:188-189 to the input size ``mid_size``, then evaluate_cityscapes.py:153,163 to ``size``),
    :168-169 fused; the intermediate tensor only ever exists in shared memory."""
    x = _req(x, torch.float32, "x")
    N, Cc, h, w = x.shape
    Hm, Wm = int(mid_size[0]), int(mid_size[1])
    H, W = int(size[0]), int(size[1])
    pred = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    check(_lib.load().asn_upsample2_argmax_u8(x.data_ptr(), N, Cc, h, w, Hm, Wm, pred.data_ptr(), H, W, _stream()),
          "asn_upsample2_argmax_u8")
    _count()
    return pred


def upsample2_argmax_hist(x, mid_size, size, label, n_cls, hist, overflow, lut=None, want_pred=False):
    """upsample2_argmax with compute_iou.py:55-57 fused in: ``hist`` (int64 [n,n]) += fast_hist(label, pred, n) for the
    frame(s), in the same kernel that forms the prediction.  Returns the uint8 prediction if ``want_pred`` else None."""
    x = _req(x, torch.float32, "x")
    label = _req(label, None, "label")
    if label.dtype not in _LABEL_CODE:
        raise TypeError(f"label dtype {label.dtype} not supported (uint8, int32, int64)")
    N, Cc, h, w = x.shape
    Hm, Wm = int(mid_size[0]), int(mid_size[1])
    H, W = int(size[0]), int(size[1])
    if label.numel() != N * H * W:
        raise ValueError("label and prediction must have the same number of pixels")
    if lut is not None:
        lut = _req(lut, torch.uint8, "lut")
    pred = torch.empty((N, H, W), dtype=torch.uint8, device=x.device) if want_pred else None
    check(_lib.load().asn_upsample2_argmax_hist(x.data_ptr(), N, Cc, h, w, Hm, Wm, pred.data_ptr() if want_pred else None,
                                                H, W, label.data_ptr(), _LABEL_CODE[label.dtype],
                                                lut.data_ptr() if lut is not None else None, int(n_cls), hist.data_ptr(),
                                                overflow.data_ptr(), _stream()), "asn_upsample2_argmax_hist")
    _count()
    return pred


def per_class_iu_device(hist: torch.Tensor):
    """(iu float64 [n], miou float64 []) on the device -- compute_iou.py:20-21,61-64."""
    hist = _req(hist, torch.int64, "hist")
    n = hist.shape[0]
    iu = torch.empty(n, dtype=torch.float64, device=hist.device)
    miou = torch.empty((), dtype=torch.float64, device=hist.device)
    check(_lib.load().asn_per_class_iu(hist.data_ptr(), n, iu.data_ptr(), miou.data_ptr(), _stream()), "asn_per_class_iu")
    _count()
    return iu, miou


# --------------------------------------------------------------------------------------
# K3 softmax cross entropy
# --------------------------------------------------------------------------------------
class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, y, ignore_label, mask_negative, weight, size_average):
        z = _req(z, torch.float32, "predict")
        y = _req(y, torch.int64, "target")
        N, Cc, H, W = z.shape
        stats = torch.empty(4, dtype=torch.int64, device=z.device)
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        wptr = _req(weight, torch.float32, "weight").data_ptr() if weight is not None else None
        check(_lib.load().asn_softmax_ce_fwd(z.data_ptr(), y.data_ptr(), N, Cc, H, W, ignore_label,
                                             int(mask_negative), wptr, int(size_average), stats.data_ptr(),
                                             loss.data_ptr(), _stream()), "asn_softmax_ce_fwd")
        _count(2)
        ctx.save_for_backward(z, y, stats, weight if weight is not None else torch.empty(0, device=z.device))
        ctx.cfg = (ignore_label, int(mask_negative), weight is not None, int(size_average))
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, gloss, _gstats):
        z, y, stats, weight = ctx.saved_tensors
        ignore_label, mask_negative, has_w, size_average = ctx.cfg
        N, Cc, H, W = z.shape
        dz = torch.empty_like(z)
        g = _req(gloss.to(torch.float32), torch.float32, "grad")
        check(_lib.load().asn_softmax_ce_bwd(z.data_ptr(), y.data_ptr(), N, Cc, H, W, ignore_label, mask_negative,
                                             weight.data_ptr() if has_w else None, size_average, stats.data_ptr(),
                                             g.data_ptr(), dz.data_ptr(), _stream()), "asn_softmax_ce_bwd")
        _count()
        return dz, None, None, None, None, None


def softmax_cross_entropy(z, y, ignore_label=255, mask_negative=False, weight=None, size_average=True,
                          return_stats=False):
    """Fused log-softmax + NLL over (N,C,H,W) logits with an ignore label.

    torch.nn.CrossEntropyLoss(ignore_index=255) as used at train_gta2cityscapes_multi.py:599-600
    (mask_negative=False) and utils/loss.py:7-36 (mask_negative=True).  stats[2] is the exact
    valid-pixel count, stats[3] the number of out-of-range targets (loss is nan if non-zero).
    """
    loss, stats = _SoftmaxCE.apply(z, y, int(ignore_label), bool(mask_negative), weight, bool(size_average))
    if os.environ.get("ASN_STRICT_LABELS") == "1" and int(stats[3].item()) != 0:
        raise IndexError("Target out of bounds")  # what torch raises for the reference
    return (loss, stats) if return_stats else loss


class _UpsampleCE(torch.autograd.Function):
    """asn_upsample_ce_fwd_bwd: loss and d loss / d z_low in one pass; backward only scales the saved gradient."""

    @staticmethod
    def forward(ctx, z_low, size, y, ignore_label, mask_negative, weight, size_average):
        z_low = _req(z_low, torch.float32, "predict")
        y = _req(y, torch.int64, "target")
        N, Cc, h, w = z_low.shape
        H, W = size
        lib = _lib.load()
        stats = torch.empty(4, dtype=torch.int64, device=z_low.device)
        loss = torch.empty((), dtype=torch.float32, device=z_low.device)
        dz = torch.empty_like(z_low)
        nbytes = lib.asn_upsample_ce_workspace_bytes(N, Cc, h, w, H, W)
        ws = _ws(nbytes, z_low.device)
        wptr = _req(weight, torch.float32, "weight").data_ptr() if weight is not None else None
        check(lib.asn_upsample_ce_fwd_bwd(z_low.data_ptr(), y.data_ptr(), N, Cc, h, w, H, W, ignore_label,
                                          int(mask_negative), wptr, int(size_average), stats.data_ptr(), loss.data_ptr(),
                                          dz.data_ptr(), ws.data_ptr(), nbytes, _stream()), "asn_upsample_ce_fwd_bwd")
        _count(2)
        ctx.save_for_backward(dz)
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, gloss, _gstats):
        (dz,) = ctx.saved_tensors
        return dz * gloss.to(torch.float32), None, None, None, None, None, None


def upsample_ce_supported(z_low, size) -> bool:
    N, Cc, h, w = z_low.shape
    return precision_ok_for_lazy() and bool(_lib.load().asn_upsample_ce_supported(Cc, h, w, int(size[0]), int(size[1])))


def precision_ok_for_lazy() -> bool:
    return True  # the lazy kernels are fp32 CUDA-core code: valid in both precision modes


def upsample_softmax_cross_entropy(z_low, size, y, ignore_label=255, mask_negative=False, weight=None,
                                   size_average=True, return_stats=False):
    """CrossEntropyLoss(ignore_index)(interp(z_low), y) without materialising interp(z_low) (Tier-B):
    model/deeplab_multi.py:188-189 + train_gta2cityscapes_multi.py:599-600 in one kernel, forward and backward.
    Falls back to the two unfused kernels for shapes the fused kernel does not cover."""
    size = (int(size[0]), int(size[1]))
    if tuple(y.shape[-2:]) != size:
        raise ValueError(f"target {tuple(y.shape)} does not match the upsampled size {size}")
    if not upsample_ce_supported(z_low, size):
        return softmax_cross_entropy(upsample_bilinear(z_low, size), y, ignore_label, mask_negative, weight,
                                     size_average, return_stats)
    loss, stats = _UpsampleCE.apply(z_low, size, y, int(ignore_label), bool(mask_negative), weight, bool(size_average))
    if os.environ.get("ASN_STRICT_LABELS") == "1" and int(stats[3].item()) != 0:
        raise IndexError("Target out of bounds")
    return (loss, stats) if return_stats else loss


# --------------------------------------------------------------------------------------
# K4 softmax over channels
# --------------------------------------------------------------------------------------
class _Softmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = _req(z, torch.float32, "input")
        N, Cc, H, W = z.shape
        p = torch.empty_like(z)
        check(_lib.load().asn_softmax_fwd(z.data_ptr(), N, Cc, H, W, p.data_ptr(), _stream()), "asn_softmax_fwd")
        _count()
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        dp = _req(dp, torch.float32, "grad")
        N, Cc, H, W = p.shape
        dz = torch.empty_like(p)
        check(_lib.load().asn_softmax_bwd(p.data_ptr(), dp.data_ptr(), N, Cc, H, W, dz.data_ptr(), _stream()),
              "asn_softmax_bwd")
        _count()
        return dz


def softmax_channels(z: torch.Tensor) -> torch.Tensor:
    """F.softmax(z) with the legacy implicit dim (= 1 for 4-D), train_gta2cityscapes_multi.py:617-618."""
    if z.dim() != 4:
        raise ValueError("softmax_channels expects (N,C,H,W)")
    return _Softmax.apply(z)


# --------------------------------------------------------------------------------------
# K6 adversarial losses against a constant label
# --------------------------------------------------------------------------------------
class _GanLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, kind):
        x = _req(x, torch.float32, "input")
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        need = x.requires_grad
        dx = torch.empty_like(x) if need else None
        check(_lib.load().asn_gan_loss_fwd_bwd(x.data_ptr(), x.numel(), float(target), int(kind), 1.0,
                                               loss.data_ptr(), dx.data_ptr() if need else None, _stream()),
              "asn_gan_loss_fwd_bwd")
        _count()
        if need:
            ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None


def gan_loss(x: torch.Tensor, target: float, kind: int = GAN_BCE) -> torch.Tensor:
    """BCEWithLogitsLoss()(x, full_like(x, target)) (kind=GAN_BCE) or MSELoss() (GAN_MSE) without
    materialising the target tensor (train_gta2cityscapes_multi.py:620-624)."""
    return _GanLoss.apply(x, float(target), int(kind))


# --------------------------------------------------------------------------------------
# fp32 convolutions on CUDA cores
# --------------------------------------------------------------------------------------
def conv2d_fwd_f32(x, w, bias, stride, pad, dil, lrelu_slope=1.0, out=None, accumulate=False):
    x = _req(x, torch.float32, "x")
    w = _req(w, torch.float32, "w")
    N, Cc, H, W = x.shape
    O, _, KH, KW = w.shape
    OH = (H + 2 * pad - dil * (KH - 1) - 1) // stride + 1
    OW = (W + 2 * pad - dil * (KW - 1) - 1) // stride + 1
    if out is None:
        out = torch.empty((N, O, OH, OW), dtype=torch.float32, device=x.device)
    b = _req(bias, torch.float32, "bias") if bias is not None else None
    check(_lib.load().asn_conv2d_fwd_f32(x.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                         out.data_ptr(), N, Cc, H, W, O, KH, KW, stride, pad, dil,
                                         float(lrelu_slope), int(accumulate), _stream()), "asn_conv2d_fwd_f32")
    _count()
    return out


def conv2d_dgrad_f32(dy, w, x_shape, stride, pad, dil, out=None, accumulate=False):
    dy = _req(dy, torch.float32, "dy")
    w = _req(w, torch.float32, "w")
    N, Cc, H, W = x_shape
    O, _, KH, KW = w.shape
    if out is None:
        out = torch.empty(tuple(x_shape), dtype=torch.float32, device=dy.device)
    check(_lib.load().asn_conv2d_dgrad_f32(dy.data_ptr(), w.data_ptr(), out.data_ptr(), N, Cc, H, W, O, KH, KW,
                                           stride, pad, dil, int(accumulate), _stream()), "asn_conv2d_dgrad_f32")
    _count()
    return out


def conv2d_wgrad_f32(x, dy, w_shape, stride, pad, dil, want_bias=True):
    x = _req(x, torch.float32, "x")
    dy = _req(dy, torch.float32, "dy")
    N, Cc, H, W = x.shape
    O, _, KH, KW = w_shape
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    db = torch.empty((O,), dtype=torch.float32, device=x.device) if want_bias else None
    check(_lib.load().asn_conv2d_wgrad_f32(x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                                           db.data_ptr() if want_bias else None, N, Cc, H, W, O, KH, KW, stride,
                                           pad, dil, _stream()), "asn_conv2d_wgrad_f32")
    _count(3)
    return dw, db


def lrelu_bwd_f32(dy, post, slope):
    dy = _req(dy, torch.float32, "dy")
    post = _req(post, torch.float32, "post")
    dx = torch.empty_like(dy)
    check(_lib.load().asn_lrelu_bwd_f32(dy.data_ptr(), post.data_ptr(), dx.data_ptr(), dy.numel(), float(slope),
                                        _stream()), "asn_lrelu_bwd_f32")
    _count()
    return dx


# --------------------------------------------------------------------------------------
# raw tcgen05 GEMM (tests / microbenchmarks)
# --------------------------------------------------------------------------------------
def gemm_bf16_tn(a: torch.Tensor, b: torch.Tensor, split_k: int = 1) -> torch.Tensor:
    """C[M,N] fp32 = A[M,K] @ B[N,K]^T for bf16 row-major A, B (K % 8 == 0)."""
    a = _req(a, torch.bfloat16, "A")
    b = _req(b, torch.bfloat16, "B")
    M, K = a.shape
    Nn, K2 = b.shape
    assert K == K2
    lib = _lib.load()
    c = torch.empty((max(split_k, 1), M, Nn), dtype=torch.float32, device=a.device)
    check(lib.asn_gemm_bf16_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, Nn, K, K, K, Nn, split_k, _stream()),
          "asn_gemm_bf16_tn")
    _count()
    return c


def gemm_bf16_nt_mn(a: torch.Tensor, b: torch.Tensor, split_k: int = 1) -> torch.Tensor:
    """C[M,N] fp32 = A[K,M]^T @ B[K,N] for bf16 row-major A, B (reduction index = row; M, N % 8 == 0)."""
    a = _req(a, torch.bfloat16, "A")
    b = _req(b, torch.bfloat16, "B")
    K, M = a.shape
    K2, Nn = b.shape
    assert K == K2
    lib = _lib.load()
    c = torch.zeros((max(split_k, 1), M, Nn), dtype=torch.float32, device=a.device)
    check(lib.asn_gemm_bf16_nt_mn(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, Nn, K, M, Nn, Nn, split_k, _stream()),
          "asn_gemm_bf16_nt_mn")
    _count()
    return c


# --------------------------------------------------------------------------------------
# K1 ASPP classifier head
# --------------------------------------------------------------------------------------
class AsppWeightPack:
    """bf16 GEMM-layout shadow of the four fp32 OIHW branch weights (asn_aspp_pack_weights).
    Re-packed when the parameters' version counters change (i.e. after optimizer.step()).  One entry per device and a
    lock: nn.DataParallel replicas share the module's pack object across devices and threads."""

    def __init__(self):
        self._by_dev = {}
        self._lock = threading.Lock()

    def key_on(self, device):
        """key of the pack currently held for `device` (None if there is none)"""
        e = self._by_dev.get(torch.device(device).index)
        return e[0] if e else None

    def invalidate(self):
        """force a re-pack on the next use (CUDA-graph capture: the pack kernels must be part of the graph)"""
        with self._lock:
            self._by_dev.clear()

    def get(self, weights, biases, n_active):
        key = tuple((w.data_ptr(), w._version) for w in list(weights) + list(biases)) + (n_active,)
        dev = weights[0].device.index
        with self._lock:
            entry = self._by_dev.get(dev)
            if entry is None or entry[0] != key:
                lib = _lib.load()
                w0 = weights[0]
                n_cls, cin = w0.shape[0], w0.shape[1]
                NP = lib.asn_aspp_np(n_cls, n_active)
                wp = torch.empty((NP, cin), dtype=torch.bfloat16, device=w0.device)
                wpt = torch.empty((cin, NP), dtype=torch.bfloat16, device=w0.device)
                ws = [_req(w.detach(), torch.float32, "weight") for w in weights[:n_active]]
                check(lib.asn_aspp_pack_weights(_lib.ptr_array([w.data_ptr() for w in ws]), n_active, n_cls, cin,
                                                wp.data_ptr(), wpt.data_ptr(), _stream()),
                      "asn_aspp_pack_weights")
                _count()
                with torch.no_grad():
                    bias_sum = torch.stack([b.detach() for b in biases[:n_active]]).sum(0).contiguous()
                entry = (key, wp, wpt, bias_sum)
                self._by_dev[dev] = entry
        return entry[1], entry[2], entry[3]


class _AsppHeadTC(torch.autograd.Function):
    """bf16 tcgen05 path: asn_aspp_fwd / asn_aspp_bwd."""

    @staticmethod
    def forward(ctx, x, pack, dils, n_active, *params):
        nb = len(params) // 2
        weights, biases = params[:nb], params[nb:]
        cl = _is_channels_last(x)
        x = x if cl else _req(x, torch.float32, "x")
        wp, wpt, bias_sum = pack.get(weights, biases, n_active)
        N, cin, H, W = x.shape
        n_cls = weights[0].shape[0]
        lib = _lib.load()
        nbytes = lib.asn_aspp_workspace_bytes(N, cin, H, W, n_cls, n_active)
        ws = _ws(nbytes, x.device)
        y = torch.empty((N, n_cls, H, W), dtype=torch.float32, device=x.device)
        need_w = any(ctx.needs_input_grad[4:4 + nb])
        # the NHWC bf16 copy the forward makes anyway is all the backward needs of x (weight gradient)
        x_bf16 = torch.empty((N * H * W, cin), dtype=torch.bfloat16, device=x.device) if need_w else None
        check(lib.asn_aspp_fwd(x.data_ptr(), int(cl), x_bf16.data_ptr() if need_w else None, wp.data_ptr(),
                               bias_sum.data_ptr(), y.data_ptr(), N, cin, H, W, n_cls,
                               _lib.int_array(dils), n_active, ws.data_ptr(), nbytes, _stream()), "asn_aspp_fwd")
        _count(3)
        ctx.save_for_backward(x_bf16 if need_w else torch.empty(0, device=x.device), wpt)
        ctx.cfg = (tuple(dils), n_active, nb, n_cls, tuple(w.shape for w in weights), cl, (N, cin, H, W))
        return y

    @staticmethod
    def backward(ctx, dy):
        x_bf16, wpt = ctx.saved_tensors
        dils, n_active, nb, n_cls, wshapes, cl, (N, cin, H, W) = ctx.cfg
        dy = _req(dy, torch.float32, "dy")
        dev = dy.device
        lib = _lib.load()
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[4:4 + nb])
        need_b = any(ctx.needs_input_grad[4 + nb:])
        nbytes = lib.asn_aspp_workspace_bytes(N, cin, H, W, n_cls, n_active)
        ws = _ws(nbytes, dev)
        dx = None
        if need_x:  # same memory format as the x that came in (NCHW or channels_last)
            dx = torch.empty((N, cin, H, W), dtype=torch.float32, device=dev,
                             memory_format=torch.channels_last if cl else torch.contiguous_format)
        dws = [torch.empty(wshapes[i], dtype=torch.float32, device=dev) for i in range(n_active)] if need_w else None
        db = torch.empty((n_cls,), dtype=torch.float32, device=dev) if need_b else None
        check(lib.asn_aspp_bwd(x_bf16.data_ptr() if need_w else None, int(cl), wpt.data_ptr(), dy.data_ptr(),
                               dx.data_ptr() if need_x else None,
                               _lib.ptr_array([t.data_ptr() for t in dws]) if need_w else None,
                               db.data_ptr() if need_b else None, N, cin, H, W, n_cls, _lib.int_array(dils),
                               n_active, ws.data_ptr(), nbytes, _stream()), "asn_aspp_bwd")
        _count(1 + ((1 if cl else N) if need_x else 0) + (2 if need_w else 0) + (1 if need_b else 0))
        gw = [(dws[i] if (need_w and i < n_active) else None) for i in range(nb)]
        # inactive branches (early-return variants) receive no gradient
        gb = [(db if (need_b and i < n_active) else None) for i in range(nb)]
        return (dx, None, None, None, *gw, *gb)


class _AsppHeadF32(torch.autograd.Function):
    """fp32 CUDA-core path (ASN_PRECISION=fp32): branch-by-branch direct convolution."""

    @staticmethod
    def forward(ctx, x, dils, n_active, *params):
        nb = len(params) // 2
        weights, biases = params[:nb], params[nb:]
        x = _req(x, torch.float32, "x")
        y = None
        for i in range(n_active):
            y = conv2d_fwd_f32(x, weights[i], biases[i], 1, dils[i], dils[i], 1.0, out=y, accumulate=i > 0)
        ctx.save_for_backward(x, *weights)
        ctx.cfg = (tuple(dils), n_active, nb)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *weights = ctx.saved_tensors
        dils, n_active, nb = ctx.cfg
        dy = _req(dy, torch.float32, "dy")
        dx = None
        gw, gb = [None] * nb, [None] * nb
        for i in range(n_active):
            if ctx.needs_input_grad[0]:
                dx = conv2d_dgrad_f32(dy, weights[i], x.shape, 1, dils[i], dils[i], out=dx, accumulate=i > 0)
            if ctx.needs_input_grad[3 + i] or ctx.needs_input_grad[3 + nb + i]:
                gw[i], gb[i] = conv2d_wgrad_f32(x, dy, weights[i].shape, 1, dils[i], dils[i], True)
        return (dx, None, None, *gw, *gb)


def aspp_head(x, weights, biases, dilations, n_active, pack: AsppWeightPack | None = None):
    """sum_{b < n_active} conv_b(x) -- Classifier_Module.forward, model/deeplab_multi.py:117-121."""
    if precision_mode() == "fp32":
        return _AsppHeadF32.apply(x, tuple(dilations), n_active, *weights, *biases)
    pack = pack if pack is not None else AsppWeightPack()
    return _AsppHeadTC.apply(x, pack, tuple(dilations), n_active, *weights, *biases)


# --------------------------------------------------------------------------------------
# K5 FCDiscriminator
# --------------------------------------------------------------------------------------
FCD_LAYERS = ("conv1", "conv2", "conv3", "conv4", "classifier")
LRELU_SLOPE = 0.2  # model/discriminator.py:16


class FcdWeightPack:
    """bf16 implicit-GEMM shadows of the discriminator's fp32 OIHW parameters
    (asn_fcd_pack_weights); rebuilt when a parameter's version counter changes.  Per device, locked (see AsppWeightPack)."""

    def __init__(self):
        self._by_dev = {}
        self._lock = threading.Lock()

    def key_on(self, device):
        """key of the pack currently held for `device` (None if there is none)"""
        e = self._by_dev.get(torch.device(device).index)
        return e[0] if e else None

    def invalidate(self):
        with self._lock:
            self._by_dev.clear()

    def get(self, params, n_cls, ndf):
        key = tuple((p.data_ptr(), p._version) for p in params)
        dev = params[0].device.index
        with self._lock:
            entry = self._by_dev.get(dev)
            if entry is None or entry[0] != key:
                lib = _lib.load()
                nbytes = lib.asn_fcd_wpack_bytes(n_cls, ndf)
                if nbytes == 0:
                    raise _lib.AsnError(f"FCDiscriminator(num_classes={n_cls}, ndf={ndf}) is outside the tensor-core "
                                        "path (needs num_classes <= 32, ndf a multiple of 64); use ASN_PRECISION=fp32")
                buf = _ws(nbytes, params[0].device)
                ps = [_req(p.detach(), torch.float32, "parameter") for p in params]
                check(lib.asn_fcd_pack_weights(_lib.ptr_array([p.data_ptr() for p in ps]), n_cls, ndf,
                                               buf.data_ptr(), _stream()), "asn_fcd_pack_weights")
                _count(1)
                entry = (key, buf)
                self._by_dev[dev] = entry
        return entry[1]


class _FcdTC(torch.autograd.Function):
    """tcgen05 path: asn_fcd_fwd / asn_fcd_bwd.  params = (conv1.w, conv1.b, ..., classifier.w, classifier.b)"""

    @staticmethod
    def forward(ctx, x, pack, x_is_logits, up_size, *params):
        x = _req(x, torch.float32, "x")
        N, n_cls, xh, xw = x.shape
        H, W = up_size if up_size is not None else (xh, xw)   # what the discriminator sees
        ndf = params[0].shape[0]
        lib = _lib.load()
        wpack = pack.get(params, n_cls, ndf)
        acts = _ws(lib.asn_fcd_acts_bytes(N, n_cls, ndf, H, W), x.device)
        if acts.numel() <= 16:
            raise _lib.AsnError(f"FCDiscriminator input {H}x{W} is outside the tensor-core path (needs >= 32x32)")
        oh, ow = H, W
        for _ in range(5):
            oh, ow = (oh + 2 - 4) // 2 + 1, (ow + 2 - 4) // 2 + 1
        out = torch.empty((N, 1, oh, ow), dtype=torch.float32, device=x.device)
        if (xh, xw) != (H, W):  # Tier-B: upsample + softmax inside the input pack
            if not x_is_logits:
                raise _lib.AsnError("a low-res discriminator input must be logits (from_logits=True)")
            check(lib.asn_fcd_fwd_lowres(x.data_ptr(), xh, xw, wpack.data_ptr(), acts.data_ptr(), out.data_ptr(), N,
                                         n_cls, ndf, H, W, None, 0, _stream()), "asn_fcd_fwd_lowres")
        else:
            check(lib.asn_fcd_fwd(x.data_ptr(), int(x_is_logits), wpack.data_ptr(), acts.data_ptr(), out.data_ptr(),
                                  N, n_cls, ndf, H, W, None, 0, _stream()), "asn_fcd_fwd")
        _count(6)
        ctx.save_for_backward(x if x_is_logits else torch.empty(0, device=x.device), wpack, acts)
        ctx.cfg = (N, n_cls, ndf, H, W, bool(x_is_logits), tuple(p.shape for p in params), (xh, xw))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wpack, acts = ctx.saved_tensors
        N, n_cls, ndf, H, W, x_is_logits, pshapes, (xh, xw) = ctx.cfg
        dout = _req(dout, torch.float32, "dout")
        lib = _lib.load()
        need_x = ctx.needs_input_grad[0]
        need_p = any(ctx.needs_input_grad[4:])
        lowres = (xh, xw) != (H, W)
        nbytes = (lib.asn_fcd_workspace_bytes_lowres(N, n_cls, ndf, H, W, xh, xw) if lowres
                  else lib.asn_fcd_workspace_bytes(N, n_cls, ndf, H, W))
        ws = _ws(nbytes, dout.device)
        dx = torch.empty((N, n_cls, xh, xw), dtype=torch.float32, device=dout.device) if need_x else None
        dps = [torch.empty(s, dtype=torch.float32, device=dout.device) for s in pshapes] if need_p else None
        pp = _lib.ptr_array([t.data_ptr() for t in dps]) if need_p else None
        if lowres:
            check(lib.asn_fcd_bwd_lowres(dout.data_ptr(), x.data_ptr(), xh, xw, wpack.data_ptr(), acts.data_ptr(),
                                         dx.data_ptr() if need_x else None, pp, N, n_cls, ndf, H, W, ws.data_ptr(),
                                         nbytes, _stream()), "asn_fcd_bwd_lowres")
        else:
            check(lib.asn_fcd_bwd(dout.data_ptr(), x.data_ptr() if x_is_logits else None, wpack.data_ptr(),
                                  acts.data_ptr(), dx.data_ptr() if need_x else None, pp, N, n_cls, ndf, H, W,
                                  ws.data_ptr(), nbytes, _stream()), "asn_fcd_bwd")
        _count(1 + (3 if need_p else 0) + 4 * (4 if need_p else 0) + 4 + ((2 if lowres else 1) if need_x else 0))
        return (dx, None, None, None, *(dps if need_p else [None] * len(pshapes)))


class FcdSaved:
    """What a tensor-core discriminator forward left behind (bf16 activations, packed weights, output):
    enough to run another backward through the SAME forward without recomputing it (fcd_replay)."""

    def __init__(self, x_logits, wpack, acts, cfg, out, key):
        self.x_logits, self.wpack, self.acts, self.cfg, self.out, self.key = x_logits, wpack, acts, cfg, out, key


class _FcdReplay(torch.autograd.Function):
    """Output of an earlier forward, re-attached to the parameters: backward = asn_fcd_bwd (parameter gradients
    only) on the saved activations.  The reference runs D(softmax(pred_target)) twice per iteration with identical
    weights and input (train...:617-618 and :665-666); the second pass is this replay."""

    @staticmethod
    def forward(ctx, saved, *params):
        ctx.saved = saved
        ctx.pshapes = tuple(p.shape for p in params)
        return saved.out.detach().clone()

    @staticmethod
    def backward(ctx, dout):
        sv = ctx.saved
        N, n_cls, ndf, H, W = sv.cfg[:5]
        dout = _req(dout, torch.float32, "dout")
        lib = _lib.load()
        nbytes = lib.asn_fcd_workspace_bytes(N, n_cls, ndf, H, W)
        ws = _ws(nbytes, dout.device)
        dps = [torch.empty(s, dtype=torch.float32, device=dout.device) for s in ctx.pshapes]
        check(lib.asn_fcd_bwd(dout.data_ptr(), None, sv.wpack.data_ptr(), sv.acts.data_ptr(), None,
                              _lib.ptr_array([t.data_ptr() for t in dps]), N, n_cls, ndf, H, W, ws.data_ptr(), nbytes,
                              _stream()), "asn_fcd_bwd")
        _count(12)
        return (None, *dps)


def fcd_replay(saved: FcdSaved, params, pack: "FcdWeightPack"):
    """D(x) for the x and weights of an earlier fcd_forward(..., return_saved=True), without recomputing it."""
    params = tuple(params)
    key = pack.key_on(params[0].device)
    if key != saved.key or key != tuple((p.data_ptr(), p._version) for p in params):
        raise _lib.AsnError("fcd_replay: the discriminator's weights changed since the saved forward")
    return _FcdReplay.apply(saved, *params)


class _FcdF32(torch.autograd.Function):
    """fp32 CUDA-core path (ASN_PRECISION=fp32)."""

    @staticmethod
    def forward(ctx, x, *params):
        h = _req(x, torch.float32, "x")
        acts = []
        for i in range(4):
            h = conv2d_fwd_f32(h, params[2 * i], params[2 * i + 1], 2, 1, 1, LRELU_SLOPE)
            acts.append(h)
        out = conv2d_fwd_f32(h, params[8], params[9], 2, 1, 1, 1.0)
        ctx.save_for_backward(x, *acts, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        x, acts, params = saved[0], list(saved[1:5]), saved[5:]
        need_x = ctx.needs_input_grad[0]
        need_p = any(ctx.needs_input_grad[1:])
        inputs = [x] + acts
        g = _req(dout, torch.float32, "dout")
        grads = [None] * 10
        for li in range(4, -1, -1):
            w = params[2 * li]
            if li < 4:
                g = lrelu_bwd_f32(g, acts[li], LRELU_SLOPE)
            if need_p:
                grads[2 * li], grads[2 * li + 1] = conv2d_wgrad_f32(inputs[li], g, w.shape, 2, 1, 1, True)
            if li > 0 or need_x:
                g = conv2d_dgrad_f32(g, w, inputs[li].shape, 2, 1, 1)
        return (g if need_x else None, *grads)


def fcd_saved_activations(out: torch.Tensor):
    """The post-LeakyReLU activations conv1..conv4 the forward saved for its backward, as fp32 NCHW
    tensors (inspection / tests: backward parity is defined on the saved activations, exactly as
    autograd defines it for the reference)."""
    fn = out.grad_fn
    if fn is None:
        raise ValueError("output has no autograd history")
    saved = fn.saved_tensors
    if type(fn).__name__.startswith("_FcdF32"):
        return [t.detach().clone() for t in saved[1:5]]
    _x, _wpack, acts = saved
    return fcd_decode_activations(acts, fn.cfg)


def fcd_decode_activations(acts: torch.Tensor, cfg):
    """conv1..conv4 post-LeakyReLU activations as fp32 NCHW tensors out of the packed bf16 NHWC buffer of a tensor-core
    forward (`cfg` = the (N, n_cls, ndf, H, W, ...) tuple of that forward, e.g. FcdSaved.cfg)."""
    N, n_cls, ndf, H, W = cfg[:5]
    lay = (C.c_int64 * 20)()
    check(_lib.load().asn_fcd_act_layout(N, n_cls, ndf, H, W, lay), "asn_fcd_act_layout")
    res = []
    for l in range(1, 5):
        off, h, w, c = lay[4 * l], lay[4 * l + 1], lay[4 * l + 2], lay[4 * l + 3]
        a = acts[off:off + N * h * w * c * 2].view(torch.bfloat16).view(N, h, w, c)
        res.append(a.permute(0, 3, 1, 2).float().contiguous())
    return res


def fcd_forward(x, params, pack: FcdWeightPack | None = None, x_is_logits: bool = False, return_saved: bool = False,
                up_size=None):
    """FCDiscriminator.forward (model/discriminator.py:21-34).  With x_is_logits the channel softmax
    F.softmax(x) of train_gta2cityscapes_multi.py:617-618 is fused into the input pack (bf16 path).
    up_size=(H, W): x holds LOW-RES logits and the discriminator sees softmax(interp(x, (H, W))) -- the bilinear
    upsample of model/deeplab_multi.py:188-189 happens inside the input pack too (Tier-B; needs x_is_logits).
    return_saved: also return an FcdSaved handle for fcd_replay (tensor-core path; None in fp32 mode)."""
    params = tuple(params)
    if up_size is not None:
        up_size = (int(up_size[0]), int(up_size[1]))
        if up_size == tuple(x.shape[-2:]):
            up_size = None
    lazy_ok = up_size is not None and x_is_logits and precision_mode() != "fp32" and \
        bool(_lib.load().asn_upsample_ce_supported(x.shape[1], x.shape[2], x.shape[3], up_size[0], up_size[1]))
    if up_size is not None and not lazy_ok:
        x = upsample_bilinear(x, up_size)
        up_size = None
    if precision_mode() == "fp32":
        if x_is_logits:
            x = softmax_channels(x)
        out = _FcdF32.apply(x, *params)
        return (out, None) if return_saved else out
    pack = pack if pack is not None else FcdWeightPack()
    out = _FcdTC.apply(x, pack, bool(x_is_logits), up_size, *params)
    if not return_saved:
        return out
    fn = out.grad_fn
    if fn is None:  # nothing requires grad: no autograd node holds the activations
        return out, None
    xs, wpack, acts = fn.saved_tensors
    return out, FcdSaved(xs if fn.cfg[5] else None, wpack, acts, fn.cfg, out, pack.key_on(params[0].device))
