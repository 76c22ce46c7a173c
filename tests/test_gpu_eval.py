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
        img, _, _ = TR.synthetic_batch(5, (65, 129), (65, 129))
        want = TR.eval_step(ref, img, size=(130, 258))                       # evaluate_cityscapes.py:153-169
        got = predict_labels(mine, img.cuda(), size=(130, 258))[0].cpu().numpy()
        assert got.dtype == np.uint8 and got.shape == want.shape
        # where they differ the reference's own top-2 margin must be at rounding level (trunk on cuDNN vs CPU)
        with torch.no_grad():
            _, lo = ref(img, (129, 65))
            up = torch.nn.Upsample(size=(130, 258), mode="bilinear", align_corners=True)(lo)[0]
        top2 = up.topk(2, dim=0).values
        margin = (top2[0] - top2[1]).numpy()
        diff = got != want
        assert diff.mean() < 0.01
        assert (margin[diff] < 1e-3 * float(up.abs().max())).all()
    finally:
        os.environ.pop("ASN_PRECISION", None)
        torch.backends.cudnn.allow_tf32 = tf32


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
