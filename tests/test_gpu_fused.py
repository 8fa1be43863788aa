"""-m gpu parity of the fused AdaIN/activation -> tcgen05 conv -> residual/statistics kernels (conv_pipe.cu for
stride-1 convolutions, conv_fused.cu otherwise) against the numpy oracle, through the C ABI.

Reference: one half-step of AdaINResBlock1.forward, Modules/hifigan.py:67-73.
Tolerance: operands are rounded to bf16/fp16 exactly as the oracle's emulation does, but the kernel evaluates Snake with
sin.approx, so an operand can land on the neighbouring 16-bit value (1 ulp = 2^-8 / 2^-11 relative): bound the
error relative to the output scale at 4e-3 (bf16) / 6e-4 (fp16); the per-layer bar of BASELINE.json is 1e-2."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_np as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G

TOL = {"bf16": 4e-3, "fp16": 6e-4}


def _ref(x, h, alpha, act, slope, w, b, res, y_old, pad, dil, scale, prec):
    C_ = x.shape[1]
    v = x
    if h is not None:
        gamma, beta = h[:, :C_, None], h[:, C_:, None]
        v = ((1 + gamma) * O.instance_norm(x) + beta).astype(np.float32)
    if act == "snake":
        v = O.snake(v, alpha)
    elif act == "lrelu":
        v = np.where(v >= 0, v, v * np.float32(slope)).astype(np.float32)
    y = O.conv1d(v, w, b, padding=pad, dilation=dil, operand=prec)
    if res is not None:
        y = y + res
    if y_old is not None:
        y = y + y_old
    return (y * np.float32(scale)).astype(np.float32)


def _coef(y, h_next):
    C_ = y.shape[1]
    y64 = y.astype(np.float64)
    mean = y64.mean(axis=2)
    var = y64.var(axis=2)
    a = (1.0 + h_next[:, :C_].astype(np.float64)) / np.sqrt(var + 1e-5)
    b = h_next[:, C_:].astype(np.float64) - mean * a
    return np.stack([a, b], axis=1)


FUSED_CASES = [
    # C, k, dil, T, res, acc, scale, act
    (32, 3, 1, 1000, True, False, 1.0, "snake"), (32, 11, 5, 777, False, False, 1.0, "snake"),
    (32, 7, 3, 100, False, False, 1.0, "snake"), (32, 11, 1, 640, True, True, 1.0 / 3.0, "snake"),
    (64, 3, 5, 515, False, False, 1.0, "snake"), (64, 11, 1, 900, True, False, 1.0, "snake"),
    (64, 7, 1, 385, True, True, 1.0, "snake"), (64, 11, 5, 400, False, False, 1.0, "lrelu"),
    (128, 3, 1, 300, True, False, 1.0, "snake"), (128, 11, 3, 517, False, False, 1.0, "snake"),
    (128, 7, 1, 260, True, True, 1.0 / 3.0, "snake"), (256, 3, 3, 200, False, False, 1.0, "snake"),
    (256, 7, 1, 129, True, False, 1.0, "snake"), (256, 11, 1, 131, True, True, 1.0, "none"),
]


@pytest.mark.parametrize("path", ["pipe", "tile"])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("C_,k,dil,T,use_res,acc,scale,act", FUSED_CASES)
def test_adain_conv1d_fused_vs_oracle(C_, k, dil, T, use_res, acc, scale, act, prec, path):
    rng = np.random.default_rng(C_ * 131 + k * 17 + dil + T)
    B = 3
    x = (rng.standard_normal((B, C_, T)) * 1.5 + 0.3).astype(np.float32)
    h = (rng.standard_normal((B, 2 * C_)) * 0.3).astype(np.float32)
    h_next = (rng.standard_normal((B, 2 * C_)) * 0.3).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, C_, 1))).astype(np.float32)
    w = (rng.standard_normal((C_, C_, k)) / np.sqrt(C_ * k)).astype(np.float32)
    b = rng.standard_normal(C_).astype(np.float32)
    res = rng.standard_normal((B, C_, T)).astype(np.float32) if use_res else None
    y_old = rng.standard_normal((B, C_, T)).astype(np.float32) if acc else None
    pad = dil * (k - 1) // 2
    ref = _ref(x, h, alpha, act, 0.1, w, b, res, y_old, pad, dil, scale, prec)
    if path == "tile":
        G._lib.check(G._lib.load().st2_set_tuning(b"no_pipe", 1))
    try:
        got, coef = G.adain_conv1d_fused(G.cl(x), h, alpha if act == "snake" else None, act, w, b,
                                         None if res is None else G.cl(res), None if y_old is None else G.cl(y_old),
                                         h_next, pad, dil, scale=scale, slope=0.1, precision=prec)
    finally:
        G._lib.check(G._lib.load().st2_set_tuning(b"no_pipe", 0))
    got = G.cf(got)
    nan = int(np.isnan(got).sum())
    err = float(np.abs(np.nan_to_num(got) - ref).max() / np.abs(ref).max())
    cref = _coef(ref, h_next)
    cerr = float(np.abs(np.nan_to_num(coef) - cref).max() / np.abs(cref).max())
    G.log("adain_conv1d_fused", path=path, prec=prec, C=C_, k=k, dil=dil, T=T, res=use_res, acc=acc, relmax=err,
          coef_relmax=cerr, nan=nan)
    assert nan == 0
    assert err <= TOL[prec]
    assert cerr <= 2 * TOL[prec]


def test_fused_residual_in_place_and_batch_edges():
    """res aliases y (the running tensor of AdaINResBlock1 is updated in place) and T is far from a tile multiple."""
    rng = np.random.default_rng(5)
    B, C_, k, T = 5, 64, 7, 130
    x = rng.standard_normal((B, C_, T)).astype(np.float32)
    h = (rng.standard_normal((B, 2 * C_)) * 0.3).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, C_, 1))).astype(np.float32)
    w = (rng.standard_normal((C_, C_, k)) / np.sqrt(C_ * k)).astype(np.float32)
    res = rng.standard_normal((B, C_, T)).astype(np.float32)
    ref = _ref(x, h, alpha, "snake", 0.0, w, None, res, None, 3, 1, 1.0, "bf16")
    lib = G._lib.load()
    xd, hd, ad, wd = G.to_dev(G.cl(x)), G.to_dev(h), G.to_dev(alpha.reshape(-1)), G.to_dev(w)
    y = G.to_dev(G.cl(res))           # residual and output are the same buffer
    nbytes = G._lib.check(lib.st2_adain_conv1d_fused_scratch_bytes(B, T, C_, C_, k))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=G.dev())
    import ctypes as C
    G._lib.check(lib.st2_adain_conv1d_fused(G._lib.ptr(xd), G._lib.ptr(hd), G._lib.ptr(ad), 2, C.c_float(0.0), G._lib.ptr(wd),
                                            None, G._lib.ptr(y), G._lib.ptr(y), None, None, G._lib.ptr(scratch), B, T, C_, C_, k,
                                            3, 1, C.c_float(1.0), 0, 1, G.stream()), "fused in place")
    torch.cuda.synchronize()
    got = G.cf(y.cpu().numpy())
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    G.log("adain_conv1d_fused_inplace", relmax=err)
    assert err <= TOL["bf16"]


# ConvTranspose1d of the generator (`ups`): polyphase sub-convolutions stacked along N in conv_pipe.cu (stride*Cout <= 256),
# one launch per phase group in conv_fused.cu otherwise
UPS_CASES = [
    # Cin, Cout, k, stride, pad, out_pad, T, act
    (64, 32, 4, 2, 1, 0, 700, "snake"), (128, 64, 6, 3, 2, 1, 333, "snake"), (256, 128, 10, 5, 3, 1, 150, "snake"),
    (512, 256, 20, 10, 5, 0, 40, "snake"), (64, 32, 4, 2, 1, 0, 129, "lrelu"), (256, 128, 12, 6, 3, 0, 100, "lrelu"),
]


@pytest.mark.parametrize("path", ["pipe", "tile"])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("Cin,Cout,k,stride,pad,opad,T,act", UPS_CASES)
def test_act_conv_transpose1d_fused_vs_oracle(Cin, Cout, k, stride, pad, opad, T, act, prec, path):
    rng = np.random.default_rng(Cin + 3 * Cout + k + T)
    B = 3
    x = (rng.standard_normal((B, Cin, T)) * 1.2).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, Cin, 1))).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / stride)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    h_next = (rng.standard_normal((B, 2 * Cout)) * 0.3).astype(np.float32)
    v = O.snake(x, alpha) if act == "snake" else np.where(x >= 0, x, x * np.float32(0.1)).astype(np.float32)
    ref = O.conv_transpose1d(v, w, b, stride=stride, padding=pad, output_padding=opad, operand=prec)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    ref = (ref + res).astype(np.float32)
    if path == "tile":
        G._lib.check(G._lib.load().st2_set_tuning(b"no_pipe", 1))
    try:
        got, coef = G.act_conv_transpose1d_fused(G.cl(x), alpha if act == "snake" else None, act, w, b, G.cl(res), h_next,
                                                 stride, pad, opad, slope=0.1, precision=prec)
    finally:
        G._lib.check(G._lib.load().st2_set_tuning(b"no_pipe", 0))
    got = G.cf(got)
    assert got.shape == ref.shape
    nan = int(np.isnan(got).sum())
    err = float(np.abs(np.nan_to_num(got) - ref).max() / np.abs(ref).max())
    cref = _coef(ref, h_next)
    cerr = float(np.abs(np.nan_to_num(coef) - cref).max() / np.abs(cref).max())
    G.log("act_conv_transpose1d_fused", path=path, prec=prec, Cin=Cin, Cout=Cout, k=k, stride=stride, T=T, relmax=err,
          coef_relmax=cerr, nan=nan)
    assert nan == 0
    assert err <= TOL[prec]
    assert cerr <= 2 * TOL[prec]


# ---- conv_row.cu: the 32 / 64-channel half-steps with fp16 stage-private tensors (residual and partial stage sum added by the
# tensor core through an identity tile, row-per-thread epilogue, statistics per (CTA, utterance, warp)).  The inputs are made
# fp16-representable first, so the oracle sees exactly what the kernel reads; an fp16 output adds half an fp16 ulp (2^-12).
def _h(a):
    return a.astype(np.float16).astype(np.float32)


ROW_CASES = [
    # C, k, dil, T, res, old, scale, x16, y16, B
    (32, 3, 1, 1000, True, False, 1.0, True, True, 3), (32, 11, 5, 777, False, False, 1.0, True, True, 3),
    (32, 7, 3, 100, False, False, 1.0, False, True, 2), (32, 11, 1, 640, True, True, 1.0 / 3.0, True, False, 3),
    (32, 3, 5, 5000, True, False, 1.0, True, True, 7), (32, 7, 1, 513, True, True, 1.0, True, True, 1),
    (64, 3, 5, 515, False, False, 1.0, True, True, 3), (64, 11, 1, 900, True, False, 1.0, True, True, 3),
    (64, 7, 1, 385, True, True, 1.0, True, False, 2), (64, 11, 5, 400, False, False, 1.0, False, True, 3),
    (64, 3, 1, 2100, True, False, 1.0, True, True, 5), (64, 7, 3, 129, True, True, 1.0 / 3.0, True, True, 4),
    (32, 3, 1, 40000, True, False, 1.0, True, True, 2),
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("C_,k,dil,T,use_res,use_old,scale,x16,y16,B", ROW_CASES)
def test_adain_conv1d_row_vs_oracle(C_, k, dil, T, use_res, use_old, scale, x16, y16, B, prec):
    rng = np.random.default_rng(C_ * 977 + k * 31 + dil * 7 + T)
    x = (rng.standard_normal((B, C_, T)) * 1.5 + 0.3).astype(np.float32)
    if x16:
        x = _h(x)
    h = (rng.standard_normal((B, 2 * C_)) * 0.3).astype(np.float32)
    h_next = (rng.standard_normal((B, 2 * C_)) * 0.3).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, C_, 1))).astype(np.float32)
    w = (rng.standard_normal((C_, C_, k)) / np.sqrt(C_ * k)).astype(np.float32)
    b = rng.standard_normal(C_).astype(np.float32)
    res = _h(rng.standard_normal((B, C_, T)).astype(np.float32)) if use_res else None
    old = _h(rng.standard_normal((B, C_, T)).astype(np.float32)) if use_old else None
    pad = dil * (k - 1) // 2
    ref = _ref(x, h, alpha, "snake", 0.0, w, b, res, old, pad, dil, scale, prec)
    got, coef = G.adain_conv1d_row(G.cl(x), h, alpha, w, b, None if res is None else G.cl(res),
                                   None if old is None else G.cl(old), h_next, pad, dil, scale=scale, precision=prec,
                                   x16=x16, y16=y16)
    got = G.cf(got)
    nan = int(np.isnan(got).sum())
    err = float(np.abs(np.nan_to_num(got) - ref).max() / np.abs(ref).max())
    cref = _coef(ref, h_next)
    cerr = float(np.abs(np.nan_to_num(coef) - cref).max() / np.abs(cref).max())
    G.log("adain_conv1d_row", prec=prec, C=C_, k=k, dil=dil, T=T, B=B, res=use_res, old=use_old, x16=x16, y16=y16, relmax=err,
          coef_relmax=cerr, nan=nan)
    assert nan == 0
    assert err <= TOL[prec] + (2.5e-4 if y16 else 0.0)
    assert cerr <= 2 * TOL[prec]
