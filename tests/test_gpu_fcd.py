"""-m gpu parity: FCDiscriminator (K5/K5b/K8) forward, input gradient (G-step) and parameter
gradients (D-step).  Tolerances: fp32 mode 1e-4, bf16 tensor-core mode 1e-2 (north star), norm-wise.

Forward: logits AND every saved activation against the exact fp64 oracle.
Backward: against the oracle's backward evaluated on the activations the forward saved -- the
definition autograd uses for the reference too.  LeakyReLU's derivative is discontinuous at 0, so a
unit whose pre-activation sits within rounding error of zero (fp32: ~1e-6 of the units, bf16: ~1e-3)
may land on the other slope than in an fp64 forward; comparing gradients across two different
forwards therefore measures those coin flips, not the backward arithmetic (with the masks fixed the
bf16 path is within 5e-3).  The all-oracle deviation is still bounded (< 0.2)
so a wrong kernel cannot hide behind this."""
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
        acts = [host(a) for a in ops.fcd_saved_activations(out)]
        out.backward(cuda(dout))
        torch.cuda.synchronize()
        return host(out), (host(xt.grad) if want_x else None), \
            [None if p.grad is None else host(p.grad) for p in pts], acts
    finally:
        os.environ.pop("ASN_PRECISION", None)


def check_backward(mode, x, params, acts, dout, dx, dps):
    """backward parity on the saved activations (bf16 path: its input is stored as bf16 too)"""
    tol = TOL[mode]
    xin = O.bf16_round(x) if mode == "bf16" else x
    dxr, gr = O.fcd_bwd(xin, params, [a.astype(np.float64) for a in acts], dout)
    if dx is not None:
        assert rel_err(dx, dxr) < tol
    if dps is not None:
        for i, n in enumerate(O.FCD_LAYERS):
            assert rel_err(dps[2 * i], gr[n][0]) < tol, n
            assert rel_err(dps[2 * i + 1], gr[n][1]) < tol, n
    return dxr, gr


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_fcd_golden_ndf64(golden, mode):
    """reference outputs (torch CPU fp32) for ndf = 64; weights regenerated from the seed"""
    g = golden("fcd")
    params = seeded_params(1338 + 40, 19, 64)
    out, dx, dps, acts = run_fcd(mode, g["ndf64_x"], params, g["ndf64_dout"])
    tol = TOL[mode]
    assert out.shape == g["ndf64_out"].shape
    assert rel_err(out, g["ndf64_out"]) < tol
    gtol = tol if mode == "fp32" else 0.2  # across two forwards: see the module docstring
    assert rel_err(dx, g["ndf64_dx"]) < gtol
    for i, n in enumerate(O.FCD_LAYERS):
        for j, kind in enumerate(("weight", "bias")):
            got = dps[2 * i + j]
            l2 = float(g[f"ndf64_dl2_{n}.{kind}"])
            assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - l2) < gtol * l2 + 1e-12
            head = g[f"ndf64_dhead_{n}.{kind}"]
            assert np.abs(got.reshape(-1)[:head.size] - head).max() < gtol * max(np.abs(got).max(), 1e-30)
    check_backward(mode, g["ndf64_x"], params, acts, g["ndf64_dout"], dx, dps)


def test_fcd_golden_small_fp32(golden):
    """ndf = 16 fixture (weights stored): only the CUDA-core path covers ndf % 64 != 0"""
    g = golden("fcd")
    params = {n: (g[f"small_{n}.weight"], g[f"small_{n}.bias"]) for n in O.FCD_LAYERS}
    out, dx, dps, acts = run_fcd("fp32", g["small_x"], params, g["small_dout"])
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
    oref, aref = O.fcd_fwd(x, params)
    dout = rng.standard_normal(oref.shape).astype(np.float32)
    dx_all, _ = O.fcd_bwd(x, params, aref, dout)
    out, dx, dps, acts = run_fcd(mode, x, params, dout)
    tol = TOL[mode]
    # forward: logits and every saved activation
    assert rel_err(out, oref) < tol
    for a, r in zip(acts, aref):
        assert a.shape == r.shape and rel_err(a, r) < tol
    # backward on the saved activations; all-oracle deviation bounded
    dxr, gr = check_backward(mode, x, params, acts, dout, dx, dps)
    assert rel_err(dx, dx_all) < 0.2
    # G-step (parameters frozen) and D-step (input detached) run the same kernels
    _, dx_g, dps_g, _ = run_fcd(mode, x, params, dout, want_x=True, want_p=False)
    assert all(p is None for p in dps_g) and rel_err(dx_g, dxr) < tol
    _, dx_d, dps_d, _ = run_fcd(mode, x, params, dout, want_x=False, want_p=True)
    assert dx_d is None and rel_err(dps_d[0], gr["conv1"][0]) < tol
    # fused softmax: feeding logits gives the gradient w.r.t. the logits
    out_l, dz, _, acts_l = run_fcd(mode, z, params, dout, want_x=True, want_p=False, logits=True)
    assert rel_err(out_l, oref) < tol
    dxl, _ = O.fcd_bwd(O.bf16_round(x) if mode == "bf16" else x, params, [a.astype(np.float64) for a in acts_l], dout)
    assert rel_err(dz, O.softmax_c_bwd(O.softmax_c(z), dxl)) < tol


def test_backward_is_bit_reproducible():
    """every gradient of the tensor-core path comes out of fixed-order reductions (split-K partials, bias partials, and --
    since round 2 -- the classifier's weight gradient, which used fp32 atomics before): two runs are bit-identical"""
    rng = np.random.default_rng(21)
    x = rng.random((1, 19, 96, 160)).astype(np.float32)
    params = seeded_params(5, 19, 64)
    dout = rng.standard_normal((1, 1, 3, 5)).astype(np.float32)
    a = run_fcd("bf16", x, params, dout)
    b = run_fcd("bf16", x, params, dout)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for ga, gb in zip(a[2], b[2]):
        assert np.array_equal(ga, gb)
