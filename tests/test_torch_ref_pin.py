"""Pins oracle/torch_ref.py (the restated reference modules + training loop used as CPU baseline and
whole-step oracle): against tests/golden/step.npz everywhere, and against the reference's own modules
wherever /root/reference is mounted (the build container)."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from oracle import torch_ref as TR

REF = os.environ.get("ASN_REFERENCE", "/root/reference")
SEED = 1338


@pytest.mark.parametrize("level,gan,tag", [("multi-level", "Vanilla", "multi"), ("single-level", "LS", "single")])
def test_restated_step_reproduces_reference_golden(golden, level, gan, tag):
    """golden = the same loop driven over the REFERENCE's modules (make_golden.py::gold_step)"""
    g = golden("step")
    G = TR.seeded_init_(TR.RefDeeplabMulti(19), SEED)
    D1 = TR.seeded_init_(TR.RefFCDiscriminator(19), SEED + 1)
    D2 = TR.seeded_init_(TR.RefFCDiscriminator(19), SEED + 2)
    tr = TR.RefTrainer(level=level, gan=gan, model=G, model_D1=D1, model_D2=D2)
    src, lab, tgt = TR.synthetic_batch(SEED, (129, 257), (97, 193))
    losses = tr.step(src, lab, tgt, do_optimizer_step=False)
    for k, v in losses.items():
        assert abs(v - float(g[f"{tag}_{k}"])) <= 1e-6 * max(1.0, abs(v)), k
    mods = {"G": G, "D2": D2}
    if level == "multi-level":
        mods["D1"] = D1
    n_checked = 0
    for key in g:
        if not key.startswith(tag + "_") or not key.endswith("_head"):
            continue
        _, name, pn = key[:-5].split("_", 2)
        p = dict(mods[name].named_parameters())[pn]
        got = p.grad.numpy().reshape(-1)[:32]
        assert np.abs(got - g[key]).max() <= 1e-5 * max(np.abs(g[key]).max(), 1e-30), key
        n_checked += 1
    assert n_checked >= 20


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_modules_match_reference_bitwise():
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import importlib
    ref_dm = importlib.import_module("model.deeplab_multi")
    ref_d = importlib.import_module("model.discriminator")
    a = TR.seeded_init_(ref_dm.DeeplabMulti(19), 3).train()
    b = TR.RefDeeplabMulti(19).train()
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    b.load_state_dict(a.state_dict())
    x = torch.randn(1, 3, 65, 97)
    ya, yb = a(x, (97, 65)), b(x, (97, 65))
    assert torch.equal(ya[0], yb[0]) and torch.equal(ya[1], yb[1])
    da = TR.seeded_init_(ref_d.FCDiscriminator(19), 4)
    db = TR.RefFCDiscriminator(19)
    db.load_state_dict(da.state_dict())
    p = torch.softmax(torch.randn(1, 19, 64, 96), 1)
    assert torch.equal(da(p), db(p))
    # the product's modules expose the same state dict and parameter-group structure
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    from adaptsegnet_b200.model.discriminator import FCDiscriminator
    mine = DeeplabMulti(19)
    assert list(mine.state_dict().keys()) == list(a.state_dict().keys())
    assert all(mine.state_dict()[k].shape == v.shape for k, v in a.state_dict().items())
    assert [p.shape for p in mine.get_1x_lr_params_NOscale()] == [p.shape for p in a.get_1x_lr_params_NOscale()]
    assert list(FCDiscriminator(19).state_dict().keys()) == list(da.state_dict().keys())
