"""Helpers for the -m gpu tests: thin ctypes callers of the unit entry points + diagnostics log."""
import ctypes as C
import json
import os

import numpy as np
import torch

from styletts2_lite_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def log(name, **kv):
    """Append one JSON line of parity metrics to gpurun_out/parity.jsonl (read back by the author)."""
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "parity.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **{k: (float(v) if isinstance(v, (np.floating, float)) else v)
                                                  for k, v in kv.items()})) + "\n")
    except OSError:
        pass


def dev():
    return torch.device("cuda", 0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=dev(), dtype=dtype).contiguous()


def cl(x_bct):
    """[B,C,T] numpy -> channels-last contiguous [B,T,C]."""
    return np.ascontiguousarray(np.transpose(x_bct, (0, 2, 1)))


def cf(x_btc):
    return np.ascontiguousarray(np.transpose(x_btc, (0, 2, 1)))


def sinegen_phase(f0, scale):
    lib = _lib.load()
    B, L2 = f0.shape
    f0d = to_dev(f0)
    phase = torch.empty(B, L2 * scale, 9, device=dev())
    frames = torch.empty(B, L2, 9, device=dev())
    _lib.check(lib.st2_sinegen_phase(_lib.ptr(f0d), _lib.ptr(phase), _lib.ptr(frames), B, L2, scale, stream()), "sinegen_phase")
    torch.cuda.synchronize()
    return phase.cpu().numpy(), frames.cpu().numpy()


def har_source(f0, noise, lin_w, lin_b, scale, seed=0):
    lib = _lib.load()
    B, L2 = f0.shape
    f0d = to_dev(f0)
    nz = None if noise is None else to_dev(noise)
    w, b = to_dev(lin_w.reshape(-1)), to_dev(lin_b.reshape(-1))
    har = torch.empty(B, L2 * scale, device=dev())
    frames = torch.empty(B, L2, 9, device=dev())
    _lib.check(lib.st2_har_source(_lib.ptr(f0d), _lib.ptr(nz), C.c_uint64(seed), _lib.ptr(w), _lib.ptr(b), _lib.ptr(har),
                                  _lib.ptr(frames), B, L2, scale, stream()), "har_source")
    torch.cuda.synchronize()
    return har.cpu().numpy()


def adain_act(x_btc, h, alpha, act, slope=0.0, out_dtype="fp32", ld_pad=0):
    """x channels-last [B,T,C]; h [B,2C] or None; returns fp32 numpy [B,T,C]."""
    lib = _lib.load()
    B, T, Cc = x_btc.shape
    ld = Cc + ld_pad
    xd = torch.zeros(B, T, ld, device=dev())
    xd[:, :, :Cc] = to_dev(x_btc)
    hd = None if h is None else to_dev(h)
    ad = None if alpha is None else to_dev(alpha.reshape(-1))
    tdt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[out_dtype]
    y = torch.zeros(B, T, ld, device=dev(), dtype=tdt)
    nbytes = _lib.check(lib.st2_adain_scratch_bytes(B, T, Cc), "adain_scratch_bytes")
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(lib.st2_adain_act(_lib.ptr(xd), ld, _lib.ptr(hd), 0 if h is None else h.shape[1], _lib.ptr(ad),
                                 _lib.ACT[act], C.c_float(slope), _lib.ptr(y), ld, _lib.DTYPE[out_dtype], B, T, Cc,
                                 _lib.ptr(scratch), stream()), "adain_act")
    torch.cuda.synchronize()
    return y[:, :, :Cc].float().cpu().numpy()


def conv1d(x_btc, w, bias, stride=1, padding=0, dilation=1, output_padding=0, transposed=False, precision="fp32"):
    """x channels-last [B,T,Cin]; w in the reference layout; returns channels-last [B,Tout,Cout]."""
    lib = _lib.load()
    B, T, Cin = x_btc.shape
    k = w.shape[2]
    Cout = w.shape[1] if transposed else w.shape[0]
    if transposed:
        Tout = (T - 1) * stride - 2 * padding + (k - 1) + output_padding + 1
    else:
        Tout = (T + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    xd, wd = to_dev(x_btc), to_dev(w)
    bd = None if bias is None else to_dev(bias)
    y = torch.full((B, Tout, Cout), float("nan"), device=dev())
    prec = _lib.PREC[precision]
    nbytes = _lib.check(lib.st2_conv1d_scratch_bytes(B, T, Cin, Cout, k, prec), "conv1d_scratch_bytes")
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(lib.st2_conv1d(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y), _lib.ptr(scratch), B, T, Cin, Cout, k,
                              stride, padding, dilation, output_padding, 1 if transposed else 0, prec, stream()), "conv1d")
    torch.cuda.synchronize()
    return y.cpu().numpy()


def adain_conv1d_fused(x_btc, h, alpha, act, w, bias, res_btc, y_old_btc, h_next, padding, dilation, scale=1.0,
                       slope=0.0, precision="bf16"):
    """Fused half-step of AdaINResBlock1 on channels-last tensors; returns (y [B,T,Cout], coef_next [B,2,Cout] or None)."""
    lib = _lib.load()
    B, T, Cin = x_btc.shape
    Cout, _, k = w.shape
    xd, wd = to_dev(x_btc), to_dev(w)
    hd = None if h is None else to_dev(h)
    ad = None if alpha is None else to_dev(alpha.reshape(-1))
    bd = None if bias is None else to_dev(bias)
    rd = None if res_btc is None else to_dev(res_btc)
    hn = None if h_next is None else to_dev(h_next)
    cn = None if h_next is None else torch.full((B, 2, Cout), float("nan"), device=dev())
    if y_old_btc is None:
        y = torch.full((B, T, Cout), float("nan"), device=dev())
    else:
        y = to_dev(y_old_btc)
    nbytes = _lib.check(lib.st2_adain_conv1d_fused_scratch_bytes(B, T, Cin, Cout, k), "fused_scratch_bytes")
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(lib.st2_adain_conv1d_fused(_lib.ptr(xd), _lib.ptr(hd), _lib.ptr(ad), _lib.ACT[act], C.c_float(slope),
                                          _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), _lib.ptr(hn), _lib.ptr(cn),
                                          _lib.ptr(scratch), B, T, Cin, Cout, k, padding, dilation, C.c_float(scale),
                                          0 if y_old_btc is None else 1, _lib.PREC[precision], stream()),
               "adain_conv1d_fused")
    torch.cuda.synchronize()
    return y.cpu().numpy(), (None if cn is None else cn.cpu().numpy())


def adain_conv1d_row(x_btc, h, alpha, w, bias, res_btc, old_btc, h_next, padding, dilation, scale=1.0, precision="bf16",
                     x16=True, y16=True):
    """The same half-step on conv_row.cu (fp16 residual / old values added by the tensor core); returns (y, coef_next)."""
    lib = _lib.load()
    B, T, Cc = x_btc.shape
    k = w.shape[2]
    xd, wd, hd, ad = to_dev(x_btc), to_dev(w), to_dev(h), to_dev(alpha.reshape(-1))
    bd = None if bias is None else to_dev(bias)
    rd = None if res_btc is None else to_dev(res_btc)
    od = None if old_btc is None else to_dev(old_btc)
    hn = None if h_next is None else to_dev(h_next)
    cn = None if h_next is None else torch.full((B, 2, Cc), float("nan"), device=dev())
    y = torch.full((B, T, Cc), float("nan"), device=dev())
    nbytes = _lib.check(lib.st2_adain_conv1d_row_scratch_bytes(B, T, Cc, k), "row_scratch_bytes")
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(lib.st2_adain_conv1d_row(_lib.ptr(xd), _lib.ptr(hd), _lib.ptr(ad), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd),
                                        _lib.ptr(od), _lib.ptr(y), _lib.ptr(hn), _lib.ptr(cn), _lib.ptr(scratch), B, T, Cc, k,
                                        padding, dilation, C.c_float(scale), _lib.PREC[precision], 1 if x16 else 0,
                                        1 if y16 else 0, stream()), "adain_conv1d_row")
    torch.cuda.synchronize()
    return y.cpu().numpy(), (None if cn is None else cn.cpu().numpy())


def act_conv_transpose1d_fused(x_btc, alpha, act, w, bias, res_btc, h_next, stride, padding, output_padding, slope=0.0,
                               precision="bf16"):
    """Fused upsampling step on channels-last tensors; returns (y [B,Tout,Cout], coef_next [B,2,Cout] or None)."""
    lib = _lib.load()
    B, T, Cin = x_btc.shape
    _, Cout, k = w.shape
    Tout = (T - 1) * stride - 2 * padding + (k - 1) + output_padding + 1
    xd, wd = to_dev(x_btc), to_dev(w)
    ad = None if alpha is None else to_dev(alpha.reshape(-1))
    bd = None if bias is None else to_dev(bias)
    rd = None if res_btc is None else to_dev(res_btc)
    hn = None if h_next is None else to_dev(h_next)
    cn = None if h_next is None else torch.full((B, 2, Cout), float("nan"), device=dev())
    y = torch.full((B, Tout, Cout), float("nan"), device=dev())
    nbytes = _lib.check(lib.st2_act_conv_transpose1d_fused_scratch_bytes(B, T, Cin, Cout, k, stride), "fused_t_scratch_bytes")
    # the residual view of the kernel may read (never use) a few rows before / after the tensor: keep it inside one arena
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    _lib.check(lib.st2_act_conv_transpose1d_fused(_lib.ptr(xd), _lib.ptr(ad), _lib.ACT[act], C.c_float(slope), _lib.ptr(wd),
                                                  _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), _lib.ptr(hn), _lib.ptr(cn),
                                                  _lib.ptr(scratch), B, T, Cin, Cout, k, stride, padding, output_padding,
                                                  _lib.PREC[precision], stream()), "act_conv_transpose1d_fused")
    torch.cuda.synchronize()
    return y.cpu().numpy(), (None if cn is None else cn.cpu().numpy())
