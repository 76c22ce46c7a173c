"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/asn_b200.h declares, the ctypes binding mirrors the header one to one, size queries work without a
GPU, and the product path fails loudly (no CPU fallback) when handed CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "asn_b200.h")


@pytest.fixture(scope="module")
def lib():
    from adaptsegnet_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    return re.findall(r"ASN_API\s+[\w\s\*]+?\b(asn_\w+)\s*\(", text)


def test_header_symbols_exported_and_bound(lib):
    from adaptsegnet_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/asn_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.asn_abi_version() == _lib.ABI_VERSION == 2


def test_size_queries_without_gpu(lib):
    assert lib.asn_aspp_np(19, 4) == 688 and lib.asn_aspp_np(19, 2) == 352
    assert lib.asn_aspp_workspace_bytes(1, 2048, 90, 160, 19, 4) > 0
    assert lib.asn_upsample_bwd_workspace_bytes(1, 19, 720, 1280, 90, 160) == 19 * 720 * 160 * 4
    assert lib.asn_fcd_wpack_bytes(19, 64) > 2 * 2781121  # both bf16 packs of every conv weight
    assert lib.asn_fcd_acts_bytes(1, 19, 64, 512, 1024) > 0
    assert lib.asn_fcd_wpack_bytes(19, 48) == 0            # outside the tensor-core path -> 0 + error text
    assert b"ndf" in lib.asn_last_error()
    lay = (ctypes.c_int64 * 20)()
    assert lib.asn_fcd_act_layout(1, 19, 64, 720, 1280, lay) == 0
    assert [lay[4 * l + 3] for l in range(5)] == [32, 64, 128, 256, 512]
    assert (lay[5], lay[6]) == (360, 640) and (lay[17], lay[18]) == (45, 80)


def test_bad_arguments_return_error_codes(lib):
    assert lib.asn_fast_hist(None, 2, None, 10, 19, None, None, None) == -1
    assert b"null" in lib.asn_last_error()
    assert lib.asn_gan_loss_fwd_bwd(None, 0, 0.0, 0, 1.0, None, None, None) == -1


def test_no_cpu_fallback():
    from adaptsegnet_b200 import _lib, ops
    from adaptsegnet_b200.model.discriminator import FCDiscriminator
    from adaptsegnet_b200.utils.loss import CrossEntropy2d
    with pytest.raises(_lib.AsnError, match="no CPU fallback"):
        ops.softmax_channels(torch.zeros(1, 19, 8, 8))
    with pytest.raises(_lib.AsnError):
        FCDiscriminator(19)(torch.zeros(1, 19, 64, 64))
    with pytest.raises(_lib.AsnError):
        CrossEntropy2d()(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(_lib.AsnError):
        ops.fast_hist(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8), 19)


def test_module_surface_matches_reference_signatures():
    """names, constructor arguments and state-dict keys a caller of the reference relies on"""
    import inspect
    from adaptsegnet_b200.model import deeplab_multi as dm
    from adaptsegnet_b200.model.discriminator import FCDiscriminator
    from adaptsegnet_b200.utils.loss import CrossEntropy2d
    assert list(inspect.signature(dm.DeeplabMulti).parameters) == ["num_classes"]
    assert list(inspect.signature(dm.ResNetMulti.forward).parameters)[:4] == ["self", "x", "input_size", "warper"]
    assert list(inspect.signature(dm.Classifier_Module.__init__).parameters)[:5] == \
        ["self", "inplanes", "dilation_series", "padding_series", "num_classes"]
    assert list(inspect.signature(FCDiscriminator.__init__).parameters) == ["self", "num_classes", "ndf"]
    assert list(inspect.signature(CrossEntropy2d.__init__).parameters) == ["self", "size_average", "ignore_label"]
    assert list(inspect.signature(CrossEntropy2d.forward).parameters) == ["self", "predict", "target", "weight"]
    m = dm.DeeplabMulti(19)
    keys = list(m.state_dict().keys())
    assert len(keys) == 640
    for b in range(4):
        assert f"layer5.conv2d_list.{b}.weight" in keys and f"layer6.conv2d_list.{b}.bias" in keys
    assert m.state_dict()["layer6.conv2d_list.0.weight"].shape == (19, 2048, 3, 3)
    d = FCDiscriminator(19)
    assert list(d.state_dict().keys()) == [f"{n}.{k}" for n in ("conv1", "conv2", "conv3", "conv4", "classifier")
                                           for k in ("weight", "bias")]
    assert sum(p.numel() for p in d.parameters()) == 2781121
    groups = m.optim_parameters(type("A", (), {"learning_rate": 2.5e-4})())
    one_x = list(groups[0]["params"])
    assert len(one_x) == 314 and len({id(p) for p in one_x}) == 104  # the reference's duplicated groups (Q11)
    assert groups[1]["lr"] == 2.5e-3
