"""CPU: host-side logic of the flat parameter storage and the fused optimizers (no kernel is launched): layout,
views, how often the reference's parameter groups name each parameter, the label-mapping LUT."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import np_oracle as O


def test_flat_params_layout_and_views():
    from adaptsegnet_b200.optim import FlatParams
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 5, 3), nn.Conv2d(5, 7, 1), nn.Linear(3, 2))
    net[0].to(memory_format=torch.channels_last)
    net[1].bias.requires_grad = False
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(list(net.parameters()) + list(net.parameters()))      # duplicates collapse
    assert len(flat.params) == 5 and flat.numel % 4 == 0
    assert all(b % 4 == 0 for b in flat.begin)                               # 16-byte aligned segments
    assert flat.numel >= sum(p.numel() for p in flat.params)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k                                  # values moved, not changed
    base = flat.values.data_ptr()
    for p in flat.params:
        assert base <= p.data_ptr() < base + flat.values.numel() * 4
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.stride() == p.stride()
    x = torch.randn(2, 3, 6, 6)
    net[1](net[0](x)).sum().backward()
    assert flat.flat.abs().sum() > 0                                          # autograd wrote into the flat buffer
    assert net[1].bias.grad is None
    flat.zero()
    assert flat.flat.abs().sum() == 0
    with torch.no_grad():
        flat.values.mul_(2.0)                                                 # one op on the buffer moves every parameter
    assert torch.allclose(net[0].weight, before["0.weight"] * 2)


def test_fused_sgd_counts_the_reference_duplicates():
    """optim_parameters() names the trunk parameters once per enclosing module: the fused step must
    repeat exactly as often as torch.optim.SGD's parameter list does."""
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    from adaptsegnet_b200.optim import FlatParams, FusedSGD

    class Args:
        learning_rate = 2.5e-4

    model = DeeplabMulti(19)
    flat = FlatParams(model.parameters())
    groups = model.optim_parameters(Args())
    groups = [{"params": list(g["params"]), "lr": g["lr"]} for g in groups]
    opt = FusedSGD(flat, groups, lr=2.5e-4, momentum=0.9, weight_decay=5e-4)
    counts = {}
    for g in groups:
        for p in g["params"]:
            counts[id(p)] = counts.get(id(p), 0) + 1
    assert sum(counts.values()) == 330 and len(counts) == 120                 # what the reference's optimizer sees
    for p, rep in zip(flat.params, opt.repeat):
        assert rep == counts[id(p)]
    # conv1 once, block convolutions 3x (layer, block, conv), the four downsample convolutions 4x, heads once
    assert sorted(set(opt.repeat)) == [1, 3, 4] and opt.repeat.count(4) == 4 and opt.repeat.count(3) == 99
    assert opt.repeat[-1] == 1
    assert [g["lr"] for g in opt.param_groups] == [2.5e-4, 2.5e-3]
    head_ids = {id(p) for p in model.layer5.parameters()} | {id(p) for p in model.layer6.parameters()}
    grp = opt.seg_group.tolist()
    for i, p in enumerate(flat.params):
        assert grp[i] == (1 if id(p) in head_ids else 0)


def test_mapping_lut_matches_label_mapping():
    from adaptsegnet_b200.ops import mapping_lut
    mapping = [[0, 255], [7, 0], [8, 1], [33, 18], [7, 2], [-1, 255]]        # a duplicate source: the later row wins
    lut = mapping_lut(mapping, "cpu").numpy()
    ids = np.arange(256)
    assert np.array_equal(lut, O.label_mapping(ids, mapping).astype(np.uint8))
    with pytest.raises(ValueError):
        mapping_lut([[3, 300]], "cpu")
