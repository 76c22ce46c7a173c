"""-m gpu parity: ASPP classifier head (K1/K1b).
bf16 tensor-core mode: 1e-2 norm-wise vs the fp64 oracle (north star: 'within 1e-2 relative,
bf16 MMA with fp32 accumulate'); fp32 mode: 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from conftest import rel_err
from gpu_util import cuda, host, gpu, feature_like

pytestmark = gpu
TOL = {"bf16": 1e-2, "fp32": 1e-4}


def run_head(mode, x, ws, bs, dy, n_active):
    from adaptsegnet_b200 import ops
    os.environ["ASN_PRECISION"] = mode
    try:
        xt = cuda(x).requires_grad_(True)
        wts = [cuda(w).requires_grad_(True) for w in ws]
        bts = [cuda(b).requires_grad_(True) for b in bs]
        y = ops.aspp_head(xt, wts, bts, O.ASPP_DILATIONS, n_active)
        y.backward(cuda(dy))
        torch.cuda.synchronize()
        return host(y), host(xt.grad), [None if w.grad is None else host(w.grad) for w in wts], \
            [None if b.grad is None else host(b.grad) for b in bts]
    finally:
        os.environ.pop("ASN_PRECISION", None)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,n_active", [("multi", 4), ("twobranch", 2)])
def test_aspp_golden(golden, mode, tag, n_active):
    g = golden("aspp")
    ws = [g[f"{tag}_w{i}"] for i in range(4)]
    bs = [g[f"{tag}_b{i}"] for i in range(4)]
    y, dx, dws, dbs = run_head(mode, g[tag + "_x"], ws, bs, g[tag + "_dy"], n_active)
    tol = TOL[mode]
    assert rel_err(y, g[tag + "_y"]) < tol
    assert rel_err(dx, g[tag + "_dx"]) < tol
    for i in range(4):
        if i < n_active:
            assert rel_err(dws[i], g[f"{tag}_dw{i}"]) < tol
            assert rel_err(dbs[i], g[f"{tag}_db{i}"]) < tol
        else:
            assert dws[i] is None and dbs[i] is None  # unused branches get no gradient (Q9)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,cin,h,w", [(1, 128, 33, 65), (1, 256, 45, 80), (2, 64, 17, 23), (1, 64, 7, 130)])
def test_aspp_oracle(mode, N, cin, h, w):
    rng = np.random.default_rng(cin + h)
    x = feature_like(rng, (N, cin, h, w))
    ws = [(rng.standard_normal((19, cin, 3, 3)) * 0.01).astype(np.float32) for _ in range(4)]
    bs = [(rng.standard_normal(19) * 0.1).astype(np.float32) for _ in range(4)]
    dy = rng.standard_normal((N, 19, h, w)).astype(np.float32)
    y, dx, dws, dbs = run_head(mode, x, ws, bs, dy, 4)
    yr = O.aspp_head_fwd(x, ws, bs)
    dxr, dwr, dbr = O.aspp_head_bwd(x, ws, dy)
    tol = TOL[mode]
    assert rel_err(y, yr) < tol
    assert rel_err(dx, dxr) < tol
    for i in range(4):
        assert rel_err(dws[i], dwr[i]) < tol
        assert rel_err(dbs[i], dbr[i]) < tol


def test_aspp_full_size_vs_fp32_path():
    """config-2 source shape (Cin 2048, 90x160): tensor-core path against the library's own fp32
    CUDA-core path (itself pinned to the oracle above) -- the oracle is too slow at this size."""
    rng = np.random.default_rng(9)
    x = feature_like(rng, (1, 2048, 90, 160), 1.6, 0.29)
    ws = [(rng.standard_normal((19, 2048, 3, 3)) * 0.01).astype(np.float32) for _ in range(4)]
    bs = [(rng.standard_normal(19) * 0.1).astype(np.float32) for _ in range(4)]
    dy = (rng.standard_normal((1, 19, 90, 160)) * 1e-3).astype(np.float32)
    ref = run_head("fp32", x, ws, bs, dy, 4)
    got = run_head("bf16", x, ws, bs, dy, 4)
    assert rel_err(got[0], ref[0]) < 1e-2
    assert rel_err(got[1], ref[1]) < 1e-2
    for i in range(4):
        assert rel_err(got[2][i], ref[2][i]) < 1e-2
        assert rel_err(got[3][i], ref[3][i]) < 1e-2


def test_aspp_channels_last_input_matches_nchw():
    """features handed over by a channels_last trunk (NHWC in memory) take the transposition-free path and
    must give the same logits / gradients as the NCHW path; dx comes back in channels_last"""
    from adaptsegnet_b200 import ops
    rng = np.random.default_rng(21)
    N, cin, h, w = 2, 128, 19, 27
    x = feature_like(rng, (N, cin, h, w))
    ws = [(rng.standard_normal((19, cin, 3, 3)) * 0.01).astype(np.float32) for _ in range(4)]
    bs = [(rng.standard_normal(19) * 0.1).astype(np.float32) for _ in range(4)]
    dy = rng.standard_normal((N, 19, h, w)).astype(np.float32)
    ref = run_head("bf16", x, ws, bs, dy, 4)
    xt = cuda(x).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert ops._is_channels_last(xt)
    wts = [cuda(v).requires_grad_(True) for v in ws]
    bts = [cuda(v).requires_grad_(True) for v in bs]
    y = ops.aspp_head(xt, wts, bts, O.ASPP_DILATIONS, 4)
    y.backward(cuda(dy))
    torch.cuda.synchronize()
    assert y.is_contiguous() and xt.grad.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(host(y), ref[0]) < 1e-5          # same bf16 operands, same GEMM
    assert rel_err(host(xt.grad), ref[1]) < 1e-3    # dgrad runs as the transposed GEMM
    for i in range(4):
        assert rel_err(host(wts[i].grad), ref[2][i]) < 1e-5
    yr = O.aspp_head_fwd(x, ws, bs)
    dxr, _, _ = O.aspp_head_bwd(x, ws, dy)
    assert rel_err(host(y), yr) < 1e-2 and rel_err(host(xt.grad), dxr) < 1e-2
