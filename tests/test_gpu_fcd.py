"""-m gpu parity: FCDiscriminator (K5/K5b/K8) forward, input gradient (G-step) and parameter
gradients (D-step).  bf16 tensor-core mode 1e-2 (norm-wise vs the fp64 oracle), fp32 mode 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from conftest import rel_err
from gpu_util import cuda, host, gpu

pytestmark = gpu
TOL = {"bf16": 1e-2, "fp32": 1e-4}


def seeded_params(seed, n_cls, ndf):
    """same generator as tests/golden/make_golden.py::fcd_params_from_seed"""
    gen = torch.Generator().manual_seed(seed)
    chans = [n_cls, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    out = {}
    for i, name in enumerate(O.FCD_LAYERS):
        bound = 1.0 / np.sqrt(chans[i] * 16)
        w = (torch.rand((chans[i + 1], chans[i], 4, 4), generator=gen) * 2 - 1) * bound
        b = (torch.rand((chans[i + 1],), generator=gen) * 2 - 1) * bound
        out[name] = (w.numpy(), b.numpy())
    return out


def run_fcd(mode, x, params, dout, want_x=True, want_p=True, logits=False):
    from adaptsegnet_b200 import ops
    os.environ["ASN_PRECISION"] = mode
    try:
        xt = cuda(x).requires_grad_(want_x)
        pts = []
        for n in O.FCD_LAYERS:
            pts += [cuda(params[n][0]).requires_grad_(want_p), cuda(params[n][1]).requires_grad_(want_p)]
        out = ops.fcd_forward(xt, pts, x_is_logits=logits)
        out.backward(cuda(dout))
        torch.cuda.synchronize()
        return host(out), (host(xt.grad) if want_x else None), [None if p.grad is None else host(p.grad) for p in pts]
    finally:
        os.environ.pop("ASN_PRECISION", None)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_fcd_golden_ndf64(golden, mode):
    g = golden("fcd")
    params = seeded_params(1338 + 40, 19, 64)
    out, dx, dps = run_fcd(mode, g["ndf64_x"], params, g["ndf64_dout"])
    tol = TOL[mode]
    assert out.shape == g["ndf64_out"].shape
    assert rel_err(out, g["ndf64_out"]) < tol
    assert rel_err(dx, g["ndf64_dx"]) < tol
    for i, n in enumerate(O.FCD_LAYERS):
        for j, kind in enumerate(("weight", "bias")):
            got = dps[2 * i + j]
            l2 = float(g[f"ndf64_dl2_{n}.{kind}"])
            assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - l2) < tol * l2 + 1e-12
            head = g[f"ndf64_dhead_{n}.{kind}"]
            assert np.abs(got.reshape(-1)[:head.size] - head).max() < tol * max(np.abs(got).max(), 1e-30)


def test_fcd_golden_small_fp32(golden):
    """ndf = 16 fixture (weights stored): only the CUDA-core path covers ndf % 64 != 0"""
    g = golden("fcd")
    params = {n: (g[f"small_{n}.weight"], g[f"small_{n}.bias"]) for n in O.FCD_LAYERS}
    out, dx, dps = run_fcd("fp32", g["small_x"], params, g["small_dout"])
    assert rel_err(out, g["small_out"]) < 1e-4 and rel_err(dx, g["small_dx"]) < 1e-4
    for i, n in enumerate(O.FCD_LAYERS):
        assert rel_err(dps[2 * i], g[f"small_d_{n}.weight"]) < 1e-4
        assert rel_err(dps[2 * i + 1], g[f"small_d_{n}.bias"]) < 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,H,W", [(1, 64, 128), (1, 90, 162), (2, 45, 80), (1, 33, 47)])
def test_fcd_oracle(mode, N, H, W):
    rng = np.random.default_rng(H * W)
    params = seeded_params(7 + H, 19, 64)
    z = (rng.standard_normal((N, 19, H, W)) * 3).astype(np.float32)
    x = O.softmax_c(z).astype(np.float32)
    oref, acts = O.fcd_fwd(x, params)
    dout = rng.standard_normal(oref.shape).astype(np.float32)
    dxr, gr = O.fcd_bwd(x, params, acts, dout)
    out, dx, dps = run_fcd(mode, x, params, dout)
    tol = TOL[mode]
    assert rel_err(out, oref) < tol
    assert rel_err(dx, dxr) < tol
    for i, n in enumerate(O.FCD_LAYERS):
        assert rel_err(dps[2 * i], gr[n][0]) < tol, n
        assert rel_err(dps[2 * i + 1], gr[n][1]) < tol, n
    # G-step (parameters frozen) and D-step (input detached) give the same numbers
    _, dx_g, dps_g = run_fcd(mode, x, params, dout, want_x=True, want_p=False)
    assert all(p is None for p in dps_g) and rel_err(dx_g, dxr) < tol
    _, dx_d, dps_d = run_fcd(mode, x, params, dout, want_x=False, want_p=True)
    assert dx_d is None and rel_err(dps_d[0], gr["conv1"][0]) < tol
    # fused softmax: feeding logits gives the gradient w.r.t. the logits
    out_l, dz, _ = run_fcd(mode, z, params, dout, want_x=True, want_p=False, logits=True)
    assert rel_err(out_l, oref) < tol
    assert rel_err(dz, O.softmax_c_bwd(O.softmax_c(z), dxr)) < tol
