"""-m gpu parity: the whole training iteration (train_gta2cityscapes_multi.py:560-683 / :373-464) on the
new modules against (a) tests/golden/step.npz -- the restated loop over the reference's own modules --
and (b) oracle/torch_ref.RefTrainer run on the host CPU with the same weights and inputs.

Losses: 1e-2 relative in bf16 mode, 1e-4 in fp32 mode (cuDNN TF32 disabled for the trunk).
Gradients (before the optimizer step): head / trunk gradients within the same tolerances x10 (they
pass through ~100 cuDNN layers); discriminator gradients are compared across two different forwards
and so carry the LeakyReLU sign flips discussed in tests/test_gpu_fcd.py: bounded at 0.2."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref as TR
from conftest import rel_err
from gpu_util import gpu

pytestmark = gpu
SEED = 1338


def build(level, gan, mode):
    from adaptsegnet_b200.train_step import AdaptSegTrainer, TrainConfig
    G = TR.seeded_init_(TR.RefDeeplabMulti(19), SEED)
    D1 = TR.seeded_init_(TR.RefFCDiscriminator(19), SEED + 1)
    D2 = TR.seeded_init_(TR.RefFCDiscriminator(19), SEED + 2)
    ref = TR.RefTrainer(level=level, gan=gan, model=G, model_D1=D1, model_D2=D2)
    mine = AdaptSegTrainer(TrainConfig(level=level, gan=gan), device="cuda")
    mine.model.load_state_dict(G.state_dict())
    mine.model_D2.load_state_dict(D2.state_dict())
    if level == "multi-level":
        mine.model_D1.load_state_dict(D1.state_dict())
    return ref, mine


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("level,gan,tag", [("multi-level", "Vanilla", "multi"), ("single-level", "LS", "single")])
def test_train_step_parity(golden, mode, level, gan, tag):
    g = golden("step")
    tol = {"bf16": 1e-2, "fp32": 1e-4}[mode]
    os.environ["ASN_PRECISION"] = mode
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    try:
        ref, mine = build(level, gan, mode)
        src, lab, tgt = TR.synthetic_batch(SEED, (129, 257), (97, 193))
        ref_losses = ref.step(src, lab, tgt, do_optimizer_step=False)
        out = mine.step(src.cuda(), lab.cuda(), tgt.cuda(), do_optimizer_step=False)
        torch.cuda.synchronize()
        for k, v in ref_losses.items():
            got = float(out[k].item())
            assert abs(v - float(g[f"{tag}_{k}"])) <= 1e-5 * max(1.0, abs(v))      # CPU box reproduces the golden
            assert abs(got - v) <= tol * max(abs(v), 1e-3), (k, got, v)
        pairs = [("G", ref.model, mine.model), ("D2", ref.model_D2, mine.model_D2)]
        if level == "multi-level":
            pairs.append(("D1", ref.model_D1, mine.model_D1))
        for name, rm, mm in pairs:
            rp, mp = dict(rm.named_parameters()), dict(mm.named_parameters())
            for pn, p in rp.items():
                if p.grad is None:
                    continue
                bound = 10 * tol if name == "G" else (0.2 if mode == "bf16" else 2e-2)
                e = rel_err(mp[pn].grad.cpu().numpy(), p.grad.numpy())
                assert e < bound, (name, pn, e)
    finally:
        os.environ.pop("ASN_PRECISION", None)
        torch.backends.cudnn.allow_tf32 = tf32


def test_optimizer_step_moves_weights_like_reference():
    """one full iteration including SGD / Adam (duplicated parameter groups, Q11): post-step weights"""
    ref, mine = build("multi-level", "Vanilla", "bf16")
    src, lab, tgt = TR.synthetic_batch(SEED + 1, (129, 257), (97, 193))
    w0 = {k: v.clone() for k, v in ref.model.state_dict().items() if v.is_floating_point()}
    ref.step(src, lab, tgt)
    mine.step(src.cuda(), lab.cuda(), tgt.cuda())
    torch.cuda.synchronize()
    for key in ("layer5.conv2d_list.0.weight", "layer6.conv2d_list.3.bias", "layer4.2.conv3.weight", "conv1.weight"):
        dr = (ref.model.state_dict()[key] - w0[key]).numpy()
        dm = (mine.model.state_dict()[key].cpu() - w0[key]).numpy()
        assert np.abs(dr).max() > 0 and rel_err(dm, dr) < 0.1, key
    for key in ("conv1.weight", "classifier.bias"):
        a = ref.model_D2.state_dict()[key].numpy()
        b = mine.model_D2.state_dict()[key].cpu().numpy()
        assert rel_err(b, a) < 1e-2, key
