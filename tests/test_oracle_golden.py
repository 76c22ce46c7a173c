"""Pins the oracle (oracle/np_oracle.py) to the reference's own outputs
(tests/golden/*.npz, generated from /root/reference by make_golden.py)."""
import numpy as np
import pytest

from oracle import np_oracle as O
from conftest import rel_err

TOL = 2e-5  # reference fixtures are torch CPU fp32; oracle is float64


@pytest.mark.parametrize("tag,n_active", [("multi", 4), ("multi_odd", 4), ("twobranch", 2)])
def test_aspp(golden, tag, n_active):
    g = golden("aspp")
    ws = [g[f"{tag}_w{i}"] for i in range(4)]
    bs = [g[f"{tag}_b{i}"] for i in range(4)]
    y = O.aspp_head_fwd(g[tag + "_x"], ws, bs, n_active=n_active)
    assert rel_err(y, g[tag + "_y"]) < TOL
    dx, dws, dbs = O.aspp_head_bwd(g[tag + "_x"], ws, g[tag + "_dy"], n_active=n_active)
    assert rel_err(dx, g[tag + "_dx"]) < TOL
    for i in range(4):
        if i < n_active:
            assert rel_err(dws[i], g[f"{tag}_dw{i}"]) < TOL
            assert rel_err(dbs[i], g[f"{tag}_db{i}"]) < TOL
        else:  # early-return variant: unused branches receive no gradient (Q9)
            assert not g[f"{tag}_dw{i}"].any() and not dws[i].any()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_upsample(golden, tag):
    g = golden("upsample")
    x, y, dy, dx = (g[f"{tag}_{k}"] for k in ("x", "y", "dy", "dx"))
    assert rel_err(O.upsample_bilinear(x, y.shape[2], y.shape[3]), y) < TOL
    assert rel_err(O.upsample_bilinear_bwd(dy, x.shape[2], x.shape[3]), dx) < TOL


def test_cross_entropy(golden):
    g = golden("ce")
    loss, nv = O.cross_entropy2d(g["z"], g["y"])
    assert nv == int(g["n_valid"])
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_err(O.cross_entropy2d_bwd(g["z"], g["y"]), g["dz"]) < TOL
    # CrossEntropy2d: negative labels masked, same value as the builtin otherwise
    loss2, nv2 = O.cross_entropy2d(g["z"], g["y_neg"], mask_negative=True)
    assert nv2 == int(g["n_valid_2d"])
    assert abs(loss2 - float(g["loss_2d"])) < 1e-5 * abs(float(g["loss_2d"]))
    assert rel_err(O.cross_entropy2d_bwd(g["z"], g["y_neg"], mask_negative=True), g["dz_2d"]) < TOL
    loss3, _ = O.cross_entropy2d(g["z"], g["y_neg"], mask_negative=True, weight=g["weight"])
    assert abs(loss3 - float(g["loss_2d_w"])) < 1e-5 * abs(float(g["loss_2d_w"]))
    assert rel_err(O.cross_entropy2d_bwd(g["z"], g["y_neg"], mask_negative=True, weight=g["weight"]),
                   g["dz_2d_w"]) < TOL
    loss4, _ = O.cross_entropy2d(g["z"], g["y_neg"], mask_negative=True, size_average=False)
    assert abs(loss4 - float(g["loss_2d_sum"])) < 1e-5 * abs(float(g["loss_2d_sum"]))
    assert rel_err(O.cross_entropy2d_bwd(g["z"], g["y_neg"], mask_negative=True, size_average=False),
                   g["dz_2d_sum"]) < TOL
    assert abs(float(g["loss_2d_same_labels"]) - float(g["loss"])) < 1e-5
    # all ignored -> nan in the reference and in the oracle (Q18)
    assert np.isnan(g["loss_all_ignored"]) and np.isnan(g["loss_2d_all_ignored"])
    la, nva = O.cross_entropy2d(g["z"], np.full_like(g["y"], 255))
    assert np.isnan(la) and nva == 0


def test_softmax(golden):
    g = golden("softmax")
    p = O.softmax_c(g["z"])
    assert rel_err(p, g["p"]) < TOL
    assert rel_err(O.softmax_c_bwd(p, g["dp"]), g["dz"]) < TOL


def test_fcd_small(golden):
    g = golden("fcd")
    params = {n: (g[f"small_{n}.weight"], g[f"small_{n}.bias"]) for n in O.FCD_LAYERS}
    out, acts = O.fcd_fwd(g["small_x"], params)
    assert out.shape == g["small_out"].shape
    assert rel_err(out, g["small_out"]) < TOL
    dx, grads = O.fcd_bwd(g["small_x"], params, acts, g["small_dout"])
    assert rel_err(dx, g["small_dx"]) < TOL
    for n in O.FCD_LAYERS:
        assert rel_err(grads[n][0], g[f"small_d_{n}.weight"]) < TOL
        assert rel_err(grads[n][1], g[f"small_d_{n}.bias"]) < TOL


@pytest.mark.parametrize("tag", ["src", "tgt", "odd"])
def test_gan_losses(golden, tag):
    g = golden("ganloss")
    x = g[tag + "_x"]
    for t in (0, 1):
        l, dx = O.bce_with_logits_const(x, t)
        assert abs(l - float(g[f"{tag}_bce{t}_loss"])) < 1e-6
        assert rel_err(dx, g[f"{tag}_bce{t}_dx"]) < TOL
        l, dx = O.mse_const(x, t)
        assert abs(l - float(g[f"{tag}_mse{t}_loss"])) < 1e-5 * max(1, abs(l))
        assert rel_err(dx, g[f"{tag}_mse{t}_dx"]) < TOL


def test_fast_hist_bit_exact(golden):
    g = golden("hist")
    h = O.fast_hist(g["a"].ravel(), g["b"].ravel(), 19)
    assert h.dtype == g["hist"].dtype and np.array_equal(h, g["hist"])
    assert np.array_equal(O.fast_hist(g["a_u8"].ravel(), g["b_u8"].ravel(), 19), g["hist_u8"])
    assert np.array_equal(O.fast_hist(g["a_spill"], g["b_spill"], 19), g["hist_spill"])
    iu = O.per_class_iu(h)
    assert np.array_equal(np.isnan(iu), np.isnan(g["iu"]))
    assert np.allclose(iu[~np.isnan(iu)], g["iu"][~np.isnan(iu)], rtol=0, atol=0)
    assert np.array_equal(O.label_mapping(g["map_in"], g["map_table"]), g["map_out"])
    with pytest.raises(ValueError):  # flat index >= n*n: the reference's reshape raises
        O.fast_hist(np.array([18]), np.array([19], dtype=np.uint8), 19)


def test_upsample_argmax_bit_exact(golden):
    g = golden("argmax")
    pred = O.upsample_argmax(g["x"], 64, 128)
    assert pred.dtype == np.uint8
    assert np.array_equal(pred, g["pred"])


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_two_stage_eval_chain_bit_exact(golden):
    """model/deeplab_multi.py:188-189 then evaluate_cityscapes.py:153,163,168-169 at shapes where the two bilinear
    stages do not compose ((out-1) % (in-1) != 0): the float32 restatement reproduces the reference's intermediate
    tensor bit for bit (sha256) and its uint8 prediction exactly."""
    g = golden("eval2")
    # the reference's own model: layer6 logits 9x17 -> 72x136 -> 144x272
    mh, mw = (int(v) for v in g["model_mid_hw"])
    assert _sha(O.upsample_bilinear_f32(g["model_low"], mh, mw)) == str(g["model_mid_sha"])
    assert np.array_equal(O.upsample_argmax(g["model_low"], 144, 272, mid=(mh, mw)), g["model_pred"])
    # fixed logits with exact ties: 33x65 -> 259x515 -> 518x1030
    assert _sha(O.upsample_bilinear_f32(g["mid_x"], 259, 515)) == str(g["mid_sha"])
    pred = O.upsample_argmax(g["mid_x"], 518, 1030, mid=(259, 515))
    assert np.array_equal(pred, g["mid_pred"])
    # a single-stage resize is NOT the reference's arithmetic at these shapes
    assert (O.upsample_argmax(g["mid_x"], 518, 1030) != g["mid_pred"]).any()


def test_input_pipeline_tail(golden):
    """dataset/gta5_dataset.py:58-71 (BGR flip, mean, CHW, id -> train id) against the reference's own dataset class run on
    PNG files of the crop size"""
    g = golden("preprocess")
    assert np.array_equal(O.gta5_image_to_tensor(g["rgb"], g["mean"]), g["image"])
    assert np.array_equal(O.gta5_label_to_trainid(g["ids"]), g["label"])
    assert tuple(g["size"]) == g["rgb"].shape
