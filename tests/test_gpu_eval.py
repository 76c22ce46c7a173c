"""-m gpu parity of the evaluation hot path (BASELINE config 5): fused upsample+argmax and the running
confusion matrix against the restated reference eval step (oracle/torch_ref.eval_step, host CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as TR
from gpu_util import gpu

pytestmark = gpu


def test_predict_labels_matches_reference_eval_step():
    from adaptsegnet_b200.evaluate import predict_labels
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    os.environ["ASN_PRECISION"] = "fp32"
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = TR.seeded_init_(TR.RefDeeplabMulti(19), 11).eval()
        mine = DeeplabMulti(19).cuda().eval()
        mine.load_state_dict(ref.state_dict())
        # 72x136 -> features 9x17: (72-1) % (9-1) != 0, the two bilinear stages of the chain do not compose into one
        img, _, _ = TR.synthetic_batch(5, (72, 136), (72, 136))
        want = TR.eval_step(ref, img, size=(144, 272))                       # evaluate_cityscapes.py:153-169
        got = predict_labels(mine, img.cuda(), size=(144, 272))[0].cpu().numpy()
        assert got.dtype == np.uint8 and got.shape == want.shape
        # where they differ the reference's own top-2 margin must be at rounding level (trunk on cuDNN vs CPU)
        with torch.no_grad():
            _, lo = ref(img, (136, 72))
            up = torch.nn.Upsample(size=(144, 272), mode="bilinear", align_corners=True)(lo)[0]
        top2 = up.topk(2, dim=0).values
        margin = (top2[0] - top2[1]).numpy()
        diff = got != want
        assert diff.mean() < 0.01
        assert (margin[diff] < 1e-3 * float(up.abs().max())).all()
    finally:
        os.environ.pop("ASN_PRECISION", None)
        torch.backends.cudnn.allow_tf32 = tf32


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_two_stage_argmax_bit_exact_vs_reference_golden(golden):
    """fixed logits -> input size -> label size -> argmax (model/deeplab_multi.py:188-189 + evaluate_cityscapes.py:
    153,163,168-169): identical to the unmodified reference's uint8 prediction, pixel for pixel, at shapes where a
    single-stage resize is wrong -- incl. BASELINE config 5's 64x128 -> 512x1024 -> 1024x2048 (pinned by digest)."""
    from adaptsegnet_b200 import ops
    g = golden("eval2")
    mh, mw = (int(v) for v in g["model_mid_hw"])
    got = ops.upsample2_argmax(torch.from_numpy(g["model_low"]).cuda(), (mh, mw), (144, 272))[0].cpu().numpy()
    assert np.array_equal(got, g["model_pred"])
    got = ops.upsample2_argmax(torch.from_numpy(g["mid_x"]).cuda(), (259, 515), (518, 1030))[0].cpu().numpy()
    assert np.array_equal(got, g["mid_pred"])
    one = ops.upsample_argmax(torch.from_numpy(g["mid_x"]).cuda(), (518, 1030))[0].cpu().numpy()
    assert (one != g["mid_pred"]).any()            # the round-1 single-stage composition is NOT the reference
    # config 5 shapes: the input is re-created from the seed (checked by digest), the prediction by digest + histogram
    x5 = torch.randn((1, 19, 64, 128), generator=torch.Generator().manual_seed(int(g["cfg5_seed"]))) * 3
    assert _sha(x5.numpy()) == str(g["cfg5_x_sha"])
    pred5 = ops.upsample2_argmax(x5.cuda(), (512, 1024), (1024, 2048))[0].cpu().numpy()
    assert np.array_equal(np.bincount(pred5.ravel(), minlength=19), g["cfg5_pred_bincount"])
    assert _sha(pred5) == str(g["cfg5_pred_sha"])
    one5 = ops.upsample_argmax(x5.cuda(), (1024, 2048))[0].cpu().numpy()
    assert int((one5 != pred5).sum()) == int(g["cfg5_one_stage_differs"])   # both kernels reproduce ATen's rounding


@pytest.mark.parametrize("low,mid,size", [((1, 19, 5, 7), (23, 40), (61, 77)), ((2, 19, 16, 32), (128, 256), (130, 515)),
                                           ((1, 3, 1, 9), (1, 33), (4, 100)), ((1, 19, 9, 17), (65, 129), (65, 129))])
def test_two_stage_argmax_oracle(low, mid, size):
    """ragged tiles, batch > 1, degenerate rows, identity second stage: bit exact against the float32 oracle"""
    from adaptsegnet_b200 import ops
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(low) * 3).astype(np.float32)
    if low[1] > 11:
        x[:, 4] = x[:, 11]
    got = ops.upsample2_argmax(torch.from_numpy(x).cuda(), mid, size).cpu().numpy()
    for n in range(low[0]):
        assert np.array_equal(got[n], O.upsample_argmax(x[n:n + 1], size[0], size[1], mid=mid))


def test_running_confusion_matrix_full_frames():
    """hist += fast_hist(label, pred, 19) over frames (compute_iou.py:57), 1024x2048 each, bit exact; IoU equal"""
    from adaptsegnet_b200.evaluate import ConfusionMatrix
    rng = np.random.default_rng(1338)
    cm = ConfusionMatrix(19)
    ref = np.zeros((19, 19), dtype=np.int64)
    for f in range(3):
        lab = np.repeat(np.repeat(rng.integers(0, 19, (64, 128)), 16, 0), 16, 1).astype(np.int64)
        lab[rng.random(lab.shape) < 0.1] = 255
        pred = np.where(rng.random(lab.shape) < 0.7, lab % 19, rng.integers(0, 19, lab.shape)).astype(np.uint8)
        ref += O.fast_hist(lab.ravel(), pred.ravel(), 19)
        cm.update(torch.from_numpy(lab).cuda(), torch.from_numpy(pred).cuda())
    assert np.array_equal(cm.hist.cpu().numpy(), ref)
    iu, iu_ref = cm.per_class_iu(), O.per_class_iu(ref)
    assert np.array_equal(np.isnan(iu), np.isnan(iu_ref)) and np.array_equal(iu[~np.isnan(iu)], iu_ref[~np.isnan(iu_ref)])


def test_numpy_fast_hist_api_like_reference():
    from adaptsegnet_b200.compute_iou import fast_hist, per_class_iu
    rng = np.random.default_rng(3)
    a = rng.integers(-1, 21, 10000).astype(np.int64)
    b = rng.integers(0, 19, 10000).astype(np.uint8)
    h = fast_hist(a, b, 19)
    assert isinstance(h, np.ndarray) and h.dtype == np.int64 and np.array_equal(h, O.fast_hist(a, b, 19))
    assert np.allclose(per_class_iu(h), O.per_class_iu(h), equal_nan=True)
    with pytest.raises(ValueError):
        fast_hist(np.array([18]), np.array([19], dtype=np.uint8), 19)


@pytest.mark.parametrize("ldtype", [np.uint8, np.int32, np.int64])
def test_fused_eval_tail_hist_bit_exact(ldtype):
    """two-stage upsample + argmax + label_mapping + confusion matrix in ONE kernel == the unfused chain
    (oracle argmax -> compute_iou.label_mapping -> fast_hist), bit exact, accumulated over frames; device mIoU == numpy"""
    from adaptsegnet_b200 import ops
    rng = np.random.default_rng(11)
    mapping = [[7, 0], [8, 1], [11, 2], [12, 3], [13, 4], [17, 5], [19, 6], [20, 7], [21, 8], [22, 9], [23, 10],
               [24, 11], [25, 12], [26, 13], [27, 14], [28, 15], [31, 16], [32, 17], [33, 18]]
    lut = ops.mapping_lut(mapping, "cuda")
    hist = torch.zeros((19, 19), dtype=torch.int64, device="cuda")
    ovf = torch.zeros(1, dtype=torch.int64, device="cuda")
    want = np.zeros((19, 19), dtype=np.int64)
    for f, (low, mid, size) in enumerate([((1, 19, 9, 17), (72, 136), (144, 272)), ((2, 19, 5, 11), (33, 70), (61, 131))]):
        x = (rng.standard_normal(low) * 3).astype(np.float32)
        raw = rng.integers(0, 34, (low[0],) + size).astype(ldtype)       # raw dataset ids; unmapped ids stay >= 19 or map
        if ldtype != np.uint8:
            raw[:, :3] = -1
        pred = ops.upsample2_argmax_hist(torch.from_numpy(x).cuda(), mid, size, torch.from_numpy(raw).cuda(), 19, hist,
                                         ovf, lut=lut, want_pred=True).cpu().numpy()
        for n in range(low[0]):
            p_ref = O.upsample_argmax(x[n:n + 1], size[0], size[1], mid=mid)
            assert np.array_equal(pred[n], p_ref)
            want += O.fast_hist(O.label_mapping(raw[n], np.array(mapping)).ravel(), p_ref.ravel(), 19)
    assert int(ovf.item()) == 0 and np.array_equal(hist.cpu().numpy(), want)
    iu, miou = ops.per_class_iu_device(hist)
    iu_ref = O.per_class_iu(want)
    assert np.array_equal(np.isnan(iu.cpu().numpy()), np.isnan(iu_ref))
    assert np.array_equal(iu.cpu().numpy()[~np.isnan(iu_ref)], iu_ref[~np.isnan(iu_ref)])
    assert abs(miou.item() - np.nanmean(iu_ref)) <= 4e-16 * abs(np.nanmean(iu_ref))   # summation order: 1-2 ulp
    # pred=None flavour accumulates the same counts
    h2 = torch.zeros_like(hist)
    ops.upsample2_argmax_hist(torch.from_numpy(x).cuda(), mid, size, torch.from_numpy(raw).cuda(), 19, h2, ovf, lut=lut)
    h3 = torch.zeros_like(hist)
    ops.fast_hist(torch.from_numpy(raw).cuda().reshape(-1), torch.from_numpy(pred).cuda().reshape(-1), 19, hist=h3, lut=lut)
    assert torch.equal(h2, h3)


def test_evaluator_graph_matches_eager_and_unfused():
    """Evaluator (CUDA graph, channels_last, fused tail) over 3 frames == predict_labels + ConfusionMatrix frame by frame"""
    from adaptsegnet_b200.evaluate import ConfusionMatrix, Evaluator, predict_labels
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    torch.manual_seed(3)
    model = DeeplabMulti(19).cuda().eval()
    ev = Evaluator(model, 19, size=(144, 272), use_cuda_graph=True, keep_pred=True)
    cm = ConfusionMatrix(19)
    for f in range(3):
        img, _, _ = TR.synthetic_batch(20 + f, (72, 136), (72, 136))
        lab = torch.randint(0, 20, (1, 144, 272), generator=torch.Generator().manual_seed(f)).to(torch.uint8)
        lab[lab == 19] = 255
        pred = ev.step(img.cuda(), lab.cuda()).clone()
        ref_pred = predict_labels(model, img.cuda().contiguous(memory_format=torch.channels_last), size=(144, 272))
        assert torch.equal(pred, ref_pred)
        cm.update(lab.cuda(), ref_pred)
    assert torch.equal(ev.hist, cm.hist) and ev.frames == 3
    iu, miou = ev.result()
    assert np.allclose(iu.cpu().numpy(), cm.per_class_iu(), equal_nan=True, rtol=0, atol=0)
