"""-m gpu parity: HBM-bound kernels (K2, K3, K4, K6, K7, K9) and the fp32 CUDA-core convolution
against the oracle and the committed golden fixtures.  Integer results bit-exact; fp32 results
within 1e-5 (norm-wise, relative to max|ref|) -- these kernels compute in fp32 like the reference."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from conftest import rel_err
from gpu_util import cuda, host, gpu

pytestmark = gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from adaptsegnet_b200 import ops as _ops
    return _ops


# ---------------------------------------------------------------- K7 fast_hist
def test_fast_hist_golden(ops, golden):
    g = golden("hist")
    for a_key, b_key, h_key in (("a", "b", "hist"), ("a_u8", "b_u8", "hist_u8"), ("a_spill", "b_spill", "hist_spill")):
        hist, ovf = ops.fast_hist(cuda(g[a_key].ravel()), cuda(g[b_key].ravel()), 19)
        assert hist.dtype == torch.int64
        assert np.array_equal(host(hist), g[h_key])
        assert int(ovf.item()) == 0


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.int64])
@pytest.mark.parametrize("n_px", [0, 1, 15, 16, 17, 4099, 1024 * 2048 + 5])
def test_fast_hist_random(ops, dtype, n_px):
    rng = np.random.default_rng(n_px + 7)
    a = rng.integers(0, 19, n_px).astype(np.int64)
    if n_px:
        a[rng.random(n_px) < 0.1] = 255
        a[rng.random(n_px) < 0.02] = 19
        if dtype != np.uint8:
            a[rng.random(n_px) < 0.02] = -1
            a[rng.random(n_px) < 0.01] = 2 ** 31 - 1 if dtype == np.int32 else 2 ** 40 + 3
    a = a.astype(dtype)
    b = rng.integers(0, 19, n_px).astype(np.uint8)
    hist, ovf = ops.fast_hist(cuda(a), cuda(b), 19)
    assert np.array_equal(host(hist), O.fast_hist(a, b, 19))
    assert int(ovf.item()) == 0


def test_fast_hist_blocky_accumulate_and_overflow(ops):
    rng = np.random.default_rng(3)
    a = np.repeat(np.repeat(rng.integers(0, 19, (64, 128)), 16, 0), 16, 1).astype(np.int64)  # 1024 x 2048 blocky
    a[:40] = 255
    b = np.where(rng.random(a.shape) < 0.85, a % 19, rng.integers(0, 19, a.shape)).astype(np.uint8)
    ref = O.fast_hist(a.ravel(), b.ravel(), 19)
    hist, _ = ops.fast_hist(cuda(a.ravel()), cuda(b.ravel()), 19)
    assert np.array_equal(host(hist), ref)
    hist, _ = ops.fast_hist(cuda(a.ravel()), cuda(b.ravel()), 19, hist=hist)  # accumulates like `hist +=`
    assert np.array_equal(host(hist), 2 * ref)
    # checksum-of-checksums property at full frame size: total count == number of valid labels
    assert int(hist.sum().item()) == 2 * int(((a >= 0) & (a < 19)).sum())
    # flat index >= n*n: the reference's reshape raises; the ABI reports the count instead
    _, ovf = ops.fast_hist(cuda(np.array([18, 18, 3], dtype=np.int64)), cuda(np.array([19, 255, 1], dtype=np.uint8)), 19)
    assert int(ovf.item()) == 2


# ---------------------------------------------------------------- K2 upsample
@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_upsample_golden(ops, golden, tag):
    g = golden("upsample")
    x = cuda(g[tag + "_x"]).requires_grad_(True)
    H, W = g[tag + "_y"].shape[2:]
    y = ops.upsample_bilinear(x, (H, W))
    assert rel_err(host(y), g[tag + "_y"]) < TOL
    y.backward(cuda(g[tag + "_dy"]))
    assert rel_err(host(x.grad), g[tag + "_dx"]) < TOL


@pytest.mark.parametrize("shape,size", [((1, 19, 33, 65), (257, 513)), ((1, 19, 45, 80), (360, 640)),
                                        ((2, 5, 7, 9), (30, 31)), ((1, 3, 16, 16), (16, 16))])
def test_upsample_oracle(ops, shape, size):
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape).astype(np.float32)
    dy = rng.standard_normal(shape[:2] + size).astype(np.float32)
    xt = cuda(x).requires_grad_(True)
    y = ops.upsample_bilinear(xt, size)
    assert rel_err(host(y), O.upsample_bilinear(x, *size)) < TOL
    y.backward(cuda(dy))
    assert rel_err(host(xt.grad), O.upsample_bilinear_bwd(dy, *shape[2:])) < TOL


def test_upsample_adjoint_full_size(ops):
    """size-independent property at config-2 size: <U x, dy> == <x, U^T dy>"""
    torch.manual_seed(0)
    x = torch.randn(1, 19, 90, 160, device="cuda")
    dy = torch.randn(1, 19, 720, 1280, device="cuda")
    y = ops.upsample_fwd_raw(x, 720, 1280)
    dx = ops.upsample_bwd_raw(dy, 90, 160)
    lhs = (y.double() * dy.double()).sum().item()
    rhs = (x.double() * dx.double()).sum().item()
    assert abs(lhs - rhs) < 1e-6 * max(abs(lhs), 1.0)
    # constant field stays constant (weights sum to one)
    ones = ops.upsample_fwd_raw(torch.full((1, 19, 90, 160), 3.25, device="cuda"), 720, 1280)
    assert torch.allclose(ones, torch.full_like(ones, 3.25), atol=1e-5)


# ---------------------------------------------------------------- K9 upsample + argmax
def test_upsample_argmax_golden(ops, golden):
    g = golden("argmax")
    pred = ops.upsample_argmax(cuda(g["x"]), (64, 128))
    assert pred.dtype == torch.uint8 and np.array_equal(host(pred)[0], g["pred"])


@pytest.mark.parametrize("shape,size", [((1, 19, 33, 65), (130, 259)), ((1, 19, 64, 128), (128, 256))])
def test_upsample_argmax_oracle(ops, shape, size):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    x[0, 4] = x[0, 11]  # exact ties: first maximum must win
    pred = ops.upsample_argmax(cuda(x), size)
    assert np.array_equal(host(pred)[0], O.upsample_argmax(x, *size))  # bit exact


# ---------------------------------------------------------------- K3 softmax + CE
def test_ce_golden(ops, golden):
    g = golden("ce")
    z = cuda(g["z"]).requires_grad_(True)
    loss, stats = ops.softmax_cross_entropy(z, cuda(g["y"]), return_stats=True)
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    assert int(stats[2].item()) == int(g["n_valid"])  # exact
    loss.backward()
    assert rel_err(host(z.grad), g["dz"]) < TOL
    # CrossEntropy2d semantics: negative labels masked; class weights; sum reduction
    z2 = cuda(g["z"]).requires_grad_(True)
    l2, st2 = ops.softmax_cross_entropy(z2, cuda(g["y_neg"]), mask_negative=True, return_stats=True)
    assert abs(l2.item() - float(g["loss_2d"])) < TOL * abs(float(g["loss_2d"]))
    assert int(st2[2].item()) == int(g["n_valid_2d"])
    l2.backward()
    assert rel_err(host(z2.grad), g["dz_2d"]) < TOL
    z3 = cuda(g["z"]).requires_grad_(True)
    l3 = ops.softmax_cross_entropy(z3, cuda(g["y_neg"]), mask_negative=True, weight=cuda(g["weight"]))
    assert abs(l3.item() - float(g["loss_2d_w"])) < TOL * abs(float(g["loss_2d_w"]))
    l3.backward()
    assert rel_err(host(z3.grad), g["dz_2d_w"]) < TOL
    z4 = cuda(g["z"]).requires_grad_(True)
    l4 = ops.softmax_cross_entropy(z4, cuda(g["y_neg"]), mask_negative=True, size_average=False)
    assert abs(l4.item() - float(g["loss_2d_sum"])) < TOL * abs(float(g["loss_2d_sum"]))
    l4.backward()
    assert rel_err(host(z4.grad), g["dz_2d_sum"]) < TOL
    # all ignored -> nan like the reference (Q18)
    la = ops.softmax_cross_entropy(cuda(g["z"]), cuda(np.full_like(g["y"], 255)))
    assert np.isnan(la.item())
    # an out-of-range target without the mask: torch raises; we flag it
    lb, stb = ops.softmax_cross_entropy(cuda(g["z"]), cuda(g["y_neg"]), return_stats=True)
    assert int(stb[3].item()) == 5 and np.isnan(lb.item())


@pytest.mark.parametrize("shape", [(1, 19, 64, 128), (2, 19, 33, 47), (1, 7, 20, 24), (1, 21, 9, 11)])
def test_ce_oracle(ops, shape):
    rng = np.random.default_rng(17)
    z = (rng.standard_normal(shape) * 4).astype(np.float32)
    y = rng.integers(0, shape[1], (shape[0],) + shape[2:]).astype(np.int64)
    y[rng.random(y.shape) < 0.1] = 255
    zt = cuda(z).requires_grad_(True)
    loss, stats = ops.softmax_cross_entropy(zt, cuda(y), return_stats=True)
    ref, nv = O.cross_entropy2d(z, y)
    assert abs(loss.item() - ref) < TOL * abs(ref) and int(stats[2].item()) == nv
    (loss * 0.37).backward()
    assert rel_err(host(zt.grad), 0.37 * O.cross_entropy2d_bwd(z, y)) < TOL


def test_ce_full_size_properties(ops):
    """config-2 size: exact valid count, shift invariance, gradient rows sum to zero"""
    torch.manual_seed(1)
    z = (torch.randn(1, 19, 720, 1280, device="cuda") * 3).requires_grad_(True)
    y = torch.randint(0, 19, (1, 720, 1280), device="cuda")
    y[torch.rand_like(y, dtype=torch.float32) < 0.1] = 255
    loss, stats = ops.softmax_cross_entropy(z, y, return_stats=True)
    assert int(stats[2].item()) == int((y != 255).sum().item())
    loss2 = ops.softmax_cross_entropy(z.detach() + 2.5, y)
    assert abs(loss.item() - loss2.item()) < 2e-5 * abs(loss.item())
    loss.backward()
    assert z.grad.sum(dim=1).abs().max().item() < 1e-9 + 1e-6 / 1000
    assert (z.grad[:, :, y[0] == 255] == 0).all()


# ---------------------------------------------------------------- K4 softmax
def test_softmax_golden_and_oracle(ops, golden):
    g = golden("softmax")
    z = cuda(g["z"]).requires_grad_(True)
    p = ops.softmax_channels(z)
    assert rel_err(host(p), g["p"]) < TOL
    p.backward(cuda(g["dp"]))
    assert rel_err(host(z.grad), g["dz"]) < TOL
    rng = np.random.default_rng(2)
    for shape in [(1, 19, 64, 128), (1, 19, 5, 7), (2, 21, 9, 10)]:
        zz = (rng.standard_normal(shape) * 5).astype(np.float32)
        dp = rng.standard_normal(shape).astype(np.float32)
        zt = cuda(zz).requires_grad_(True)
        pt = ops.softmax_channels(zt)
        pr = O.softmax_c(zz)
        assert rel_err(host(pt), pr) < TOL
        assert abs(host(pt).sum(1) - 1).max() < 1e-5
        pt.backward(cuda(dp))
        assert rel_err(host(zt.grad), O.softmax_c_bwd(pr, dp)) < TOL


# ---------------------------------------------------------------- K6 GAN losses
@pytest.mark.parametrize("tag", ["src", "tgt", "odd"])
def test_gan_loss_golden(ops, golden, tag):
    g = golden("ganloss")
    for kind, lname in ((ops.GAN_BCE, "bce"), (ops.GAN_MSE, "mse")):
        for t in (0, 1):
            x = cuda(g[tag + "_x"]).requires_grad_(True)
            loss = ops.gan_loss(x, t, kind)
            ref = float(g[f"{tag}_{lname}{t}_loss"])
            assert abs(loss.item() - ref) < TOL * max(1.0, abs(ref))
            (loss * 0.001).backward()
            assert rel_err(host(x.grad), 0.001 * g[f"{tag}_{lname}{t}_dx"]) < TOL


# ---------------------------------------------------------------- fp32 CUDA-core convolution
@pytest.mark.parametrize("cfg", [
    dict(N=1, C=19, H=34, W=50, O=16, K=4, s=2, p=1, d=1),
    dict(N=2, C=8, H=20, W=28, O=19, K=3, s=1, p=12, d=12),
    dict(N=1, C=70, H=9, W=11, O=5, K=3, s=1, p=24, d=24),
    dict(N=1, C=16, H=5, W=7, O=1, K=4, s=2, p=1, d=1),
])
def test_conv_f32_oracle(ops, cfg):
    rng = np.random.default_rng(23)
    x = rng.standard_normal((cfg["N"], cfg["C"], cfg["H"], cfg["W"])).astype(np.float32)
    w = (rng.standard_normal((cfg["O"], cfg["C"], cfg["K"], cfg["K"])) * 0.1).astype(np.float32)
    b = rng.standard_normal(cfg["O"]).astype(np.float32)
    s, p, d = cfg["s"], cfg["p"], cfg["d"]
    yref = O.conv2d(x, w, b, s, p, d)
    y = ops.conv2d_fwd_f32(cuda(x), cuda(w), cuda(b), s, p, d)
    assert rel_err(host(y), yref) < 1e-5
    ylr = ops.conv2d_fwd_f32(cuda(x), cuda(w), cuda(b), s, p, d, lrelu_slope=0.2)
    assert rel_err(host(ylr), O.leaky_relu(yref)) < 1e-5
    dy = rng.standard_normal(yref.shape).astype(np.float32)
    dxr, dwr, dbr = O.conv2d_bwd(x, w, dy, s, p, d)
    dx = ops.conv2d_dgrad_f32(cuda(dy), cuda(w), x.shape, s, p, d)
    assert rel_err(host(dx), dxr) < 1e-5
    dw, db = ops.conv2d_wgrad_f32(cuda(x), cuda(dy), w.shape, s, p, d)
    assert rel_err(host(dw), dwr) < 1e-5 and rel_err(host(db), dbr) < 1e-5


# ---------------------------------------------------------------- K7 + label_mapping
CITYSCAPES_LABEL2TRAIN = [[0, 255], [1, 255], [2, 255], [3, 255], [4, 255], [5, 255], [6, 255], [7, 0], [8, 1], [9, 255],
                          [10, 255], [11, 2], [12, 3], [13, 4], [14, 255], [15, 255], [16, 255], [17, 5], [18, 255], [19, 6],
                          [20, 7], [21, 8], [22, 9], [23, 10], [24, 11], [25, 12], [26, 13], [27, 14], [28, 15], [29, 255],
                          [30, 255], [31, 16], [32, 17], [33, 18], [-1, 255]]


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.int64])
def test_fast_hist_with_label_mapping(ops, dtype):
    """label_mapping (compute_iou.py:24-28) fused into fast_hist == the reference's two steps, bit for bit"""
    rng = np.random.default_rng(21)
    n_px = 1024 * 2048 + 3
    raw = rng.integers(0, 34, n_px).astype(np.int64)
    if dtype != np.uint8:
        raw[rng.random(n_px) < 0.01] = -1
        raw[rng.random(n_px) < 0.01] = 300
    pred = rng.integers(0, 19, n_px).astype(np.uint8)
    ref = O.fast_hist(O.label_mapping(raw, CITYSCAPES_LABEL2TRAIN), pred, 19)
    lut = ops.mapping_lut(CITYSCAPES_LABEL2TRAIN, "cuda")
    hist, ovf = ops.fast_hist(cuda(raw.astype(dtype)), cuda(pred), 19, lut=lut)
    assert np.array_equal(host(hist), ref) and int(ovf.item()) == 0
    from adaptsegnet_b200 import compute_iou as CI
    assert np.array_equal(CI.fast_hist(raw.astype(dtype), pred, 19, mapping=CITYSCAPES_LABEL2TRAIN), ref)
    assert np.array_equal(CI.label_mapping(raw, CITYSCAPES_LABEL2TRAIN), O.label_mapping(raw, CITYSCAPES_LABEL2TRAIN))


# ---------------------------------------------------------------- input pipeline tail
def test_input_pipeline_kernels(ops, golden):
    g = golden("preprocess")
    img = ops.image_to_tensor(cuda(g["rgb"][None]), g["mean"])
    assert np.array_equal(host(img)[0], g["image"])                      # bit exact vs the reference's dataset class
    lab = ops.label_to_trainid(cuda(g["ids"][None]))
    assert lab.dtype == torch.int64 and np.array_equal(host(lab)[0], g["label"].astype(np.int64))
    rng = np.random.default_rng(9)                                       # a batch at the real size
    rgb = rng.integers(0, 256, (2, 720, 1280, 3)).astype(np.uint8)
    ids = rng.integers(0, 256, (2, 720, 1280)).astype(np.uint8)
    mean = np.array((104.00698793, 116.66876762, 122.67891434), dtype=np.float32)
    out = host(ops.image_to_tensor(cuda(rgb), mean))
    lab = host(ops.label_to_trainid(cuda(ids)))
    for n in range(2):
        assert np.array_equal(out[n], O.gta5_image_to_tensor(rgb[n], mean))
        assert np.array_equal(lab[n], O.gta5_label_to_trainid(ids[n]).astype(np.int64))
