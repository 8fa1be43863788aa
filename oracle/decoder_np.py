"""CPU oracle for the StyleTTS2-lite waveform Decoder hot path  --  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the reference algorithm.  Nothing in the product
package (`styletts2_lite_b200/`) may import it; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs use it, and only as the checker / the CPU arm -- never as the thing shipped.

Parity pinning: the reference ships no tests, golden vectors or fixtures for
this path, and its arithmetic lives in PyTorch ATen
(pinned torch 2.7.0 in uv.lock:2116; installed here: 2.11.0).  The oracle is
therefore pinned against outputs of the *reference itself* run in the authoring
container: tests/golden/make_golden.py imports /root/reference/Modules/{hifigan,
istftnet}.py, loads the synthetic state_dict of styletts2_lite_b200.synth into
the unmodified reference Decoder, replays a shared noise tape and commits the
results under tests/golden/*.npz.  tests/test_oracle.py checks this file against
those fixtures (bit-exact for the SineGen phase and the length regulator;
<= 5e-5 max-abs for waveforms, the reference's own fp32-vs-fp64 self-consistency
level).

Every function cites the reference lines it restates (paths relative to the
reference root).  Layout here follows the reference: activations [B, C, T].
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

F32 = np.float32
F64 = np.float64


# ----------------------------------------------------------------------------
# operand rounding emulation (used to study the bf16 / fp16 tensor-core path)
# ----------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    bias = ((u >> 16) & 1) + np.uint32(0x7FFF)
    r = ((u + bias) >> 16) << 16
    return r.astype(np.uint32).view(F32).reshape(x.shape)


def round_operand(x: np.ndarray, mode: Optional[str]) -> np.ndarray:
    if mode is None or mode == "fp32":
        return x
    if mode == "bf16":
        return round_bf16(x)
    if mode == "fp16":
        return x.astype(np.float16).astype(F32)
    raise ValueError(mode)


# ----------------------------------------------------------------------------
# primitives (ATen ops the reference dispatches to)
# ----------------------------------------------------------------------------
def fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """torch._weight_norm(v, g, dim=0): w = v * (g / ||v||), the norm taken over
    every dim except 0 (torch.nn.utils.weight_norm as applied at hifigan.py:30-45,
    292, 317, 373, 377-382, 434-439).  For ConvTranspose1d dim 0 is Cin."""
    v = v.astype(F32)
    nrm = np.sqrt((v.astype(F64) ** 2).reshape(v.shape[0], -1).sum(1)).astype(F32)
    nrm = nrm.reshape((v.shape[0],) + (1,) * (v.ndim - 1))
    return (v * (g.astype(F32) / nrm)).astype(F32)


def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, operand=None):
    """nn.Conv1d, groups=1.  x [B,Cin,T], w [Cout,Cin,k] -> [B,Cout,Tout]."""
    B, Cin, T = x.shape
    Cout, Cin2, k = w.shape
    assert Cin == Cin2
    x = round_operand(x, operand)
    w = round_operand(w, operand)
    Tout = (T + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    xp = np.zeros((B, Cin, T + 2 * padding), F32)
    xp[:, :, padding:padding + T] = x
    out = np.zeros((B, Cout, Tout), F32)
    for kk in range(k):
        st = kk * dilation
        xs = xp[:, :, st: st + (Tout - 1) * stride + 1: stride]
        out += np.matmul(w[None, :, :, kk], xs)
    if b is not None:
        out += b.reshape(1, -1, 1)
    return out


def conv_transpose1d(x, w, b=None, stride=1, padding=0, output_padding=0, groups=1, operand=None):
    """nn.ConvTranspose1d.  x [B,Cin,T], w [Cin,Cout/groups,k] -> [B,Cout,Tout]
    (groups is either 1 or Cin, the only two cases on the path: generator.ups.*
    hifigan.py:292-294 and the depthwise `pool` hifigan.py:373)."""
    B, Cin, T = x.shape
    k = w.shape[2]
    x = round_operand(x, operand)
    w = round_operand(w, operand)
    Tout = (T - 1) * stride - 2 * padding + (k - 1) + output_padding + 1
    Lfull = (T - 1) * stride + k + output_padding
    if groups == 1:
        Cout = w.shape[1]
        full = np.zeros((B, Cout, Lfull), F32)
        for kk in range(k):
            full[:, :, kk: kk + (T - 1) * stride + 1: stride] += np.matmul(w[None, :, :, kk].transpose(0, 2, 1), x)
    else:
        assert groups == Cin and w.shape[1] == 1
        Cout = Cin
        full = np.zeros((B, Cout, Lfull), F32)
        for kk in range(k):
            full[:, :, kk: kk + (T - 1) * stride + 1: stride] += w[None, :, 0, kk, None] * x
    out = full[:, :, padding: padding + Tout].copy()
    if b is not None:
        out += b.reshape(1, -1, 1)
    return out


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm1d(affine=False, track_running_stats=False) (hifigan.py:17):
    per-(b,c) mean and *biased* variance over T.  ATen's CPU native_batch_norm
    accumulates the statistics in double and applies y = x*invstd + (-mean*invstd)
    in fp32."""
    xd = x.astype(F64)
    mean = xd.mean(axis=2, keepdims=True)
    var = ((xd - mean) ** 2).mean(axis=2, keepdims=True)
    invstd = (1.0 / np.sqrt(var + eps)).astype(F32)
    beta = (-(mean.astype(F32)) * invstd).astype(F32)
    return (x * invstd + beta).astype(F32)


def adain(x, s, fc_w, fc_b):
    """AdaIN1d.forward (hifigan.py:20-24): h = fc(s); gamma,beta = chunk(h);
    (1 + gamma) * IN(x) + beta."""
    h = (s @ fc_w.T + fc_b).astype(F32)            # [B, 2C]
    C = x.shape[1]
    gamma = h[:, :C, None]
    beta = h[:, C:, None]
    return ((F32(1) + gamma) * instance_norm(x) + beta).astype(F32)


def snake(x, alpha):
    """Snake1D as written at hifigan.py:68: x + (1/a) * sin(a*x)**2, a [1,C,1]."""
    a = alpha.reshape(1, -1, 1).astype(F32)
    return (x + (F32(1) / a) * (np.sin(a * x) ** 2)).astype(F32)


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * F32(slope)).astype(F32)


def get_padding(k, d=1):
    """Modules/utils.py:13-14."""
    return int((k * d - d) / 2)


# ----------------------------------------------------------------------------
# weight access
# ----------------------------------------------------------------------------
class Weights:
    """Read-only view of a reference state_dict (numpy), folding weight-norm pairs."""

    def __init__(self, sd: Dict[str, np.ndarray]):
        self.sd = {k: np.asarray(v, dtype=F32) for k, v in sd.items()}
        self._cache: Dict[str, np.ndarray] = {}

    def w(self, name):
        if name in self._cache:
            return self._cache[name]
        if name + ".weight_g" in self.sd:
            w = fold_weight_norm(self.sd[name + ".weight_g"], self.sd[name + ".weight_v"])
        elif name + ".parametrizations.weight.original0" in self.sd:      # parametrizations.weight_norm (vocos.py:10)
            w = fold_weight_norm(self.sd[name + ".parametrizations.weight.original0"],
                                 self.sd[name + ".parametrizations.weight.original1"])
        else:
            w = self.sd[name + ".weight"]
        self._cache[name] = w
        return w

    def b(self, name):
        return self.sd.get(name + ".bias")

    def p(self, name):
        return self.sd[name]


# ----------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------
def adain_resblk1d(W: Weights, name, x, s, upsample, taps=None, operand=None):
    """AdainResBlk1d.forward (hifigan.py:359-403): residual(norm1 -> lrelu(0.2) ->
    pool -> conv1 -> norm2 -> lrelu(0.2) -> conv2) + shortcut([nearest x2] -> conv1x1),
    divided by sqrt(2).  dropout p=0 is the identity."""
    h = adain(x, s, W.p(name + ".norm1.fc.weight"), W.p(name + ".norm1.fc.bias"))
    h = leaky_relu(h, 0.2)
    if upsample:
        h = conv_transpose1d(h, W.w(name + ".pool"), W.b(name + ".pool"), stride=2, padding=1,
                             output_padding=1, groups=h.shape[1])
    h = conv1d(h, W.w(name + ".conv1"), W.b(name + ".conv1"), padding=1, operand=operand)
    if taps is not None:
        taps[name + ".conv1"] = h
    h = adain(h, s, W.p(name + ".norm2.fc.weight"), W.p(name + ".norm2.fc.bias"))
    h = leaky_relu(h, 0.2)
    h = conv1d(h, W.w(name + ".conv2"), W.b(name + ".conv2"), padding=1, operand=operand)
    sc = x
    if upsample:
        sc = np.repeat(sc, 2, axis=2)              # F.interpolate(scale_factor=2, 'nearest') hifigan.py:414
    if (name + ".conv1x1.weight_v") in W.sd or (name + ".conv1x1.parametrizations.weight.original1") in W.sd:
        sc = conv1d(sc, W.w(name + ".conv1x1"), None, operand=operand)
    out = ((h + sc) / F32(math.sqrt(2))).astype(F32)
    if taps is not None:
        taps[name] = out
    return out


def adain_resblock1(W: Weights, name, x, s, k, dilations=(1, 3, 5), taps=None, operand=None):
    """AdaINResBlock1.forward (hifigan.py:65-74)."""
    for j, d in enumerate(dilations):
        xt = adain(x, s, W.p("%s.adain1.%d.fc.weight" % (name, j)), W.p("%s.adain1.%d.fc.bias" % (name, j)))
        xt = snake(xt, W.p("%s.alpha1.%d" % (name, j)))
        xt = conv1d(xt, W.w("%s.convs1.%d" % (name, j)), W.b("%s.convs1.%d" % (name, j)),
                    padding=get_padding(k, d), dilation=d, operand=operand)
        if taps is not None:
            taps["%s.convs1.%d" % (name, j)] = xt
        xt = adain(xt, s, W.p("%s.adain2.%d.fc.weight" % (name, j)), W.p("%s.adain2.%d.fc.bias" % (name, j)))
        xt = snake(xt, W.p("%s.alpha2.%d" % (name, j)))
        xt = conv1d(xt, W.w("%s.convs2.%d" % (name, j)), W.b("%s.convs2.%d" % (name, j)),
                    padding=get_padding(k, 1), dilation=1, operand=operand)
        x = (xt + x).astype(F32)
        if taps is not None:
            taps["%s.iter%d" % (name, j)] = x
    return x


# ----------------------------------------------------------------------------
# NSF harmonic source (bit-exact phase)
# ----------------------------------------------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add emulated through float64 (a*b is exact in double)."""
    return (a.astype(F64) * b.astype(F64) + c.astype(F64)).astype(F32)


def sinegen_phase_frames(f0_curve: np.ndarray, upsample_scale: int, sr: int = 24000, harmonics: int = 9):
    """Per-frame phase pf[B,L2,9] fed to the x`upsample_scale` linear up-sampler.
    SineGen._f02sine (hifigan.py:117-157), steps 2-5 of SURVEY.md §8(a) "SineGen exact recipe":
      fn = f0*h (fp32); rad = fmod(fn/sr, 1);  the 1/scale linear down-sample reads two
      samples of the same frame with weights 1/2,1/2 -> rad of that frame exactly (the
      `rand_ini` added at sample 0, hifigan.py:126-129, is never read);
      cumsum accumulates in double and rounds to fp32 (ATen CPU cumsum);
      pf = ((cs*2)*pi32)*scale."""
    f0 = f0_curve.astype(F32)
    h = np.arange(1, harmonics + 1, dtype=F32)
    fn = (f0[:, :, None] * h[None, None, :]).astype(F32)
    rad = np.fmod((fn / F32(sr)).astype(F32), F32(1)).astype(F32)
    rad = np.where((rad != 0) & (rad < 0), (rad + F32(1)).astype(F32), rad)   # torch `%` == python modulo
    cs = np.cumsum(rad.astype(F64), axis=1).astype(F32)
    pf = (cs * F32(2)).astype(F32)
    pf = (pf * F32(np.pi)).astype(F32)
    pf = (pf * F32(upsample_scale)).astype(F32)
    return pf


def sinegen_phase(f0_curve: np.ndarray, upsample_scale: int, sr: int = 24000, harmonics: int = 9):
    """phase[B,S,9] = F.interpolate(pf, scale_factor=upsample_scale, mode='linear')
    (hifigan.py:155-156), step 6 of the recipe: align_corners=False source index with the
    fp32 rounding of ATen's CPU upsample_linear1d kernel."""
    pf = sinegen_phase_frames(f0_curve, upsample_scale, sr, harmonics)
    B, L2, H = pf.shape
    S = L2 * upsample_scale
    n = np.arange(S, dtype=F32)
    scale = F32(1.0 / upsample_scale)
    src = _fma32(np.full(S, scale, F32), (n + F32(0.5)).astype(F32), np.full(S, -0.5, F32))
    src = np.maximum(src, F32(0))
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, L2 - 1)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1) - l1).astype(F32)
    p0 = pf[:, i0, :]
    p1 = pf[:, i1, :]
    t = (l1[None, :, None] * p1).astype(F32)
    phase = _fma32(np.broadcast_to(l0[None, :, None], p0.shape), p0, t)
    return phase


def source_module(W: Weights, f0_curve, upsample_scale, noise):
    """SourceModuleHnNSF.forward (hifigan.py:254-268) -> har_source [B,1,S].
    `noise` is the `torch.randn_like(sine_waves)` draw of hifigan.py:213, shape [B,S,9]."""
    f0_up = np.repeat(f0_curve.astype(F32), upsample_scale, axis=1)[:, :, None]   # nearest, hifigan.py:284,323
    phase = sinegen_phase(f0_curve, upsample_scale)
    sines = np.sin(phase).astype(F32)
    sine_waves = (sines * F32(0.1)).astype(F32)
    uv = (f0_up > F32(10.0)).astype(F32)
    noise_amp = (uv * F32(0.003) + ((F32(1) - uv) * F32(0.1)).astype(F32) / F32(3)).astype(F32)
    nz = (noise_amp * noise.astype(F32)).astype(F32)
    sine_waves = (sine_waves * uv + nz).astype(F32)
    lw = W.p("generator.m_source.l_linear.weight")       # [1,9]
    lb = W.p("generator.m_source.l_linear.bias")
    merged = np.tanh((sine_waves @ lw.T + lb).astype(F32)).astype(F32)           # [B,S,1]
    return merged.transpose(0, 2, 1)


# ----------------------------------------------------------------------------
# CustomSTFT (istftnet.py:111-301)
# ----------------------------------------------------------------------------
def stft_transform(W: Weights, wave, n_fft=20, hop=5):
    """CustomSTFT.transform (istftnet.py:207-243): replicate pad n_fft/2, two strided
    convs with the windowed DFT basis, magnitude with +1e-14, atan2 phase with the
    (im==0 & re<0) -> pi fix-up."""
    pad = n_fft // 2
    xp = np.pad(wave.astype(F32), ((0, 0), (pad, pad)), mode="edge")[:, None, :]
    re = conv1d(xp, W.p("generator.stft.weight_forward_real"), None, stride=hop)
    im = conv1d(xp, W.p("generator.stft.weight_forward_imag"), None, stride=hop)
    mag = np.sqrt((re * re + im * im + F32(1e-14)).astype(F32)).astype(F32)
    ph = np.arctan2(im, re).astype(F32)
    ph[(im == 0) & (re < 0)] = F32(np.pi)
    return mag, ph


def stft_inverse(W: Weights, mag, phase, n_fft=20, hop=5):
    """CustomSTFT.inverse (istftnet.py:246-293): two transposed convs, real - imag,
    trim n_fft/2 each side.  No window-envelope normalisation (as in the reference)."""
    re = (mag * np.cos(phase)).astype(F32)
    im = (mag * np.sin(phase)).astype(F32)
    r = conv_transpose1d(re, W.p("generator.stft.weight_backward_real"), None, stride=hop)
    i = conv_transpose1d(im, W.p("generator.stft.weight_backward_imag"), None, stride=hop)
    wave = (r - i).astype(F32)
    pad = n_fft // 2
    return wave[:, :, pad:-pad]


# ----------------------------------------------------------------------------
# generators
# ----------------------------------------------------------------------------
def _operand_for(name, operand, hp_groups):
    if operand is None:
        return None
    for g, mode in (hp_groups or {}).items():
        if g in name:
            return mode
    return operand


def generator_hifigan(W: Weights, cfg, x, s, f0_curve, noise, taps=None, operand=None, hp_groups=None,
                      har_source=None):
    """Generator.forward (hifigan.py:321-347)."""
    if har_source is None:
        har_source = source_module(W, f0_curve, cfg.upsample_scale, noise)
    if taps is not None:
        taps["har_source"] = har_source
    nk = len(cfg.resblock_kernel_sizes)
    for i in range(cfg.num_stages):
        x = snake(x, W.p("generator.alphas.%d" % i))
        _, k, st, pd = cfg.noise_conv_geometry(i)
        nm = "generator.noise_convs.%d" % i
        xs_ = conv1d(har_source, W.w(nm), W.b(nm), stride=st, padding=pd)
        nm = "generator.noise_res.%d" % i
        xs_ = adain_resblock1(W, nm, xs_, s, cfg.noise_res_kernel(i), taps=taps,
                              operand=_operand_for(nm, operand, hp_groups))
        ku, u, pu, opu = cfg.ups_geometry(i)
        nm = "generator.ups.%d" % i
        x = conv_transpose1d(x, W.w(nm), W.b(nm), stride=u, padding=pu, output_padding=opu,
                             operand=_operand_for(nm, operand, hp_groups))
        x = (x + xs_).astype(F32)
        if taps is not None:
            taps["generator.stage%d.in" % i] = x
        acc = None
        for j, kr in enumerate(cfg.resblock_kernel_sizes):
            nm = "generator.resblocks.%d" % (i * nk + j)
            r = adain_resblock1(W, nm, x, s, kr, cfg.resblock_dilation_sizes[j], taps=taps,
                                operand=_operand_for(nm, operand, hp_groups))
            acc = r if acc is None else (acc + r).astype(F32)
        x = (acc / F32(nk)).astype(F32)
        if taps is not None:
            taps["generator.stage%d.out" % i] = x
    x = snake(x, W.p("generator.alphas.%d" % cfg.num_stages))
    x = conv1d(x, W.w("generator.conv_post"), W.b("generator.conv_post"), padding=3)
    return np.tanh(x).astype(F32)


def generator_istftnet(W: Weights, cfg, x, s, f0_curve, noise, taps=None, operand=None, hp_groups=None,
                       har_source=None):
    """Generator.forward (istftnet.py:542-573)."""
    n_fft, hop = cfg.gen_istft_n_fft, cfg.gen_istft_hop_size
    if har_source is None:
        har_source = source_module(W, f0_curve, cfg.upsample_scale, noise)
    if taps is not None:
        taps["har_source"] = har_source
    mag, ph = stft_transform(W, har_source[:, 0, :], n_fft, hop)
    har = np.concatenate([mag, ph], axis=1)
    if taps is not None:
        taps["har"] = har
    nk = len(cfg.resblock_kernel_sizes)
    for i in range(cfg.num_stages):
        x = leaky_relu(x, 0.1)
        _, k, st, pd = cfg.noise_conv_geometry(i)
        nm = "generator.noise_convs.%d" % i
        xs_ = conv1d(har, W.w(nm), W.b(nm), stride=st, padding=pd)
        nm = "generator.noise_res.%d" % i
        xs_ = adain_resblock1(W, nm, xs_, s, cfg.noise_res_kernel(i), taps=taps,
                              operand=_operand_for(nm, operand, hp_groups))
        ku, u, pu, opu = cfg.ups_geometry(i)
        nm = "generator.ups.%d" % i
        x = conv_transpose1d(x, W.w(nm), W.b(nm), stride=u, padding=pu, output_padding=opu,
                             operand=_operand_for(nm, operand, hp_groups))
        if i == cfg.num_stages - 1:
            x = np.concatenate([x[:, :, 1:2], x], axis=2)       # ReflectionPad1d((1,0)) istftnet.py:558-559
        x = (x + xs_).astype(F32)
        if taps is not None:
            taps["generator.stage%d.in" % i] = x
        acc = None
        for j, kr in enumerate(cfg.resblock_kernel_sizes):
            nm = "generator.resblocks.%d" % (i * nk + j)
            r = adain_resblock1(W, nm, x, s, kr, cfg.resblock_dilation_sizes[j], taps=taps,
                                operand=_operand_for(nm, operand, hp_groups))
            acc = r if acc is None else (acc + r).astype(F32)
        x = (acc / F32(nk)).astype(F32)
        if taps is not None:
            taps["generator.stage%d.out" % i] = x
    x = leaky_relu(x, 0.01)                                      # F.leaky_relu default slope, istftnet.py:569
    x = conv1d(x, W.w("generator.conv_post"), W.b("generator.conv_post"), padding=3)
    bins = n_fft // 2 + 1
    spec = np.exp(x[:, :bins, :]).astype(F32)
    phase = np.sin(x[:, bins:, :]).astype(F32)
    if taps is not None:
        taps["spec"] = spec
        taps["phase"] = phase
    return stft_inverse(W, spec, phase, n_fft, hop)


def gelu(x):
    """nn.GELU() (vocos.py:48), the exact erf form."""
    from scipy.special import erf
    x64 = x.astype(F64)
    return (0.5 * x64 * (1.0 + erf(x64 / math.sqrt(2.0)))).astype(F32)


def convnext_block(W: Weights, name, x, s, operand=None, taps=None):
    """ConvNeXtBlock.forward (vocos.py:57-69): depthwise Conv1d(k=7, pad 3) -> AdaIN1d -> Linear(dim, intermediate) -> GELU ->
    Linear(intermediate, dim) -> gamma * x -> + residual.  x [B, dim, T]."""
    B, C, T = x.shape
    w, b = W.p(name + ".dwconv.weight"), W.p(name + ".dwconv.bias")
    xp = np.zeros((B, C, T + 6), F32)
    xp[:, :, 3:3 + T] = x
    h = np.zeros((B, C, T), F32)
    for k in range(7):
        h += w[None, :, 0, k, None] * xp[:, :, k:k + T]
    h = (h + b[None, :, None]).astype(F32)
    if taps is not None:
        taps[name + ".dwconv"] = h
    h = adain(h, s, W.p(name + ".norm.fc.weight"), W.p(name + ".norm.fc.bias"))
    h = h.transpose(0, 2, 1)                                                      # [B, T, C]
    h = (round_operand(h, operand) @ round_operand(W.p(name + ".pwconv1.weight"), operand).T + W.p(name + ".pwconv1.bias")).astype(F32)
    h = gelu(h)
    h = (round_operand(h, operand) @ round_operand(W.p(name + ".pwconv2.weight"), operand).T + W.p(name + ".pwconv2.bias")).astype(F32)
    h = (W.p(name + ".gamma") * h).astype(F32)
    out = (x + h.transpose(0, 2, 1)).astype(F32)
    if taps is not None:
        taps[name] = out
    return out


def istft_same(W: Weights, spec, n_fft, hop):
    """ISTFT.forward with padding='same' (vocos.py:195-232): irfft (norm 'backward') per frame, times the window, overlap-add
    with stride hop, trim (win - hop) / 2 samples on both sides, divide by the overlap-added squared window.
    spec complex [B, n_fft/2+1, T] -> [B, T*hop]."""
    win = W.p("generator.stft.istft.window").astype(F32)
    B, _, T = spec.shape
    pad = (n_fft - hop) // 2
    frames = (np.fft.irfft(spec.astype(np.complex128), n=n_fft, axis=1).astype(F32) * win[None, :, None]).astype(F32)
    size = (T - 1) * hop + n_fft
    y = np.zeros((B, size), F32)
    env = np.zeros(size, F32)
    wsq = (win * win).astype(F32)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, :, t]
        env[t * hop:t * hop + n_fft] += wsq
    assert (env[pad:size - pad] > 1e-11).all()                                    # vocos.py:229
    return (y[:, pad:size - pad] / env[None, pad:size - pad]).astype(F32)


def generator_vocos(W: Weights, cfg, x, s, taps=None, operand=None):
    """Generator.forward + ISTFTHead.forward (vocos.py:159-164, :268-296).  x [B, dim, 2T] (the front half's output) ->
    waveform [B, 1, 2T * hop]."""
    for i in range(cfg.num_layers):
        x = convnext_block(W, "generator.convnext.%d" % i, x, s, operand, taps)
    h = x.transpose(0, 2, 1).astype(F64)                                          # final LayerNorm(dim, eps=1e-6), vocos.py:153
    mean, var = h.mean(axis=2, keepdims=True), h.var(axis=2, keepdims=True)
    h = ((h - mean) / np.sqrt(var + 1e-6)).astype(F32)
    h = (h * W.p("generator.final_layer_norm.weight") + W.p("generator.final_layer_norm.bias")).astype(F32)
    if taps is not None:
        taps["generator.final_layer_norm"] = h
    op = "fp16" if operand else None                                              # the head runs on fp16 operands in the 16-bit modes
    o = (round_operand(h, op) @ round_operand(W.p("generator.stft.out.weight"), op).T + W.p("generator.stft.out.bias")).astype(F32)
    o = o.transpose(0, 2, 1)                                                      # [B, n_fft + 2, T]
    half = o.shape[1] // 2
    mag = np.minimum(np.exp(o[:, :half]), F32(1e2)).astype(F32)                   # vocos.py:282-283
    p = o[:, half:]
    spec = (mag * np.cos(p)).astype(F32) + 1j * (mag * np.sin(p)).astype(F32)     # vocos.py:285-292
    if taps is not None:
        taps["generator.stft.out"] = o
    return istft_same(W, spec, cfg.gen_istft_n_fft, cfg.gen_istft_hop_size)[:, None, :]


def decoder_forward(sd: Dict[str, np.ndarray], cfg, asr, F0_curve, N, s, noise,
                    taps: Optional[dict] = None, operand: Optional[str] = None,
                    hp_groups: Optional[Dict[str, str]] = None, har_source=None):
    """Decoder.forward in eval mode (hifigan.py:446-475 / istftnet.py:692-721).
    asr [B,512,T], F0_curve [B,2T], N [B,2T], s [B,128], noise [B,S,9] -> [B,1,S].

    operand: None (fp32) | 'bf16' | 'fp16' rounds the operands of the dense convs
    (AdainResBlk1d convs, resblock convs, ups) to emulate the tensor-core path;
    hp_groups maps a name fragment to a different operand mode for that group."""
    W = sd if isinstance(sd, Weights) else Weights(sd)
    asr = asr.astype(F32)
    s = s.astype(F32)
    F0 = conv1d(F0_curve.astype(F32)[:, None, :], W.w("F0_conv"), W.b("F0_conv"), stride=2, padding=1)
    Nn = conv1d(N.astype(F32)[:, None, :], W.w("N_conv"), W.b("N_conv"), stride=2, padding=1)
    x = np.concatenate([asr, F0, Nn], axis=1)
    x = adain_resblk1d(W, "encode", x, s, False, taps, _operand_for("encode", operand, hp_groups))
    asr_res = conv1d(asr, W.w("asr_res.0"), W.b("asr_res.0"), operand=_operand_for("asr_res", operand, hp_groups))
    for i in range(4):
        x = np.concatenate([x, asr_res, F0, Nn], axis=1)
        x = adain_resblk1d(W, "decode.%d" % i, x, s, i == 3, taps, _operand_for("decode", operand, hp_groups))
    if taps is not None:
        taps["decode.out"] = x
    if getattr(cfg, "is_vocos", False):
        return generator_vocos(W, cfg, x, s, taps, operand)
    gen = generator_istftnet if cfg.is_istft else generator_hifigan
    return gen(W, cfg, x, s, F0_curve.astype(F32), noise, taps, operand, hp_groups, har_source)


# ----------------------------------------------------------------------------
# length regulator
# ----------------------------------------------------------------------------
def round_durations(duration: np.ndarray) -> np.ndarray:
    """inference.py:257: torch.round (half to even) then clamp(min=1) -> int64."""
    return np.maximum(np.rint(duration.astype(F32)), 1).astype(np.int64)


def replace_outliers_zscore(x: np.ndarray, threshold: float = 3.0, factor: float = 0.95) -> np.ndarray:
    """StyleTTS2.__replace_outliers_zscore (inference.py:134-148) on a 1-D fp32 slice: entries with |z| > threshold become
    mean + sign * threshold * std * factor; std is torch's unbiased one (NaN for a single element: nothing is replaced)."""
    x = x.astype(F32)
    if x.size == 0:
        return x
    mean = F32(x.astype(np.float64).mean())
    with np.errstate(invalid="ignore", divide="ignore"):
        std = F32(np.sqrt(((x.astype(np.float64) - np.float64(mean)) ** 2).sum() / (x.size - 1))) if x.size > 1 else F32(np.nan)
        z = ((x - mean) / std).astype(F32)
        mask = np.abs(z) > F32(threshold)
        repl = (mean + np.sign(x - mean).astype(F32) * F32(F32(F32(threshold) * std) * F32(factor))).astype(F32)
    out = x.copy()
    out[mask] = repl[mask]
    return out


def smooth_durations(duration: np.ndarray, noise: Optional[np.ndarray], t: float = 0.1, speed: float = 1.0,
                     prev_d_mean: float = 0.0):
    """inference.py:248-255 for ONE sentence, duration [L] fp32, noise [L] the N(0,1) tape in place of normal_'s own draw
    (dur_stats = noise * std + mean, what normal_(mean, std) computes): returns (duration for the rounding, its mean --
    inference.py:272)."""
    d = duration.astype(F32)
    n = d.size
    mean = F32(d.astype(np.float64).mean())
    with np.errstate(invalid="ignore", divide="ignore"):
        std = F32(np.sqrt(((d.astype(np.float64) - np.float64(mean)) ** 2).sum() / (n - 1))) if n > 1 else F32(np.nan)
    mu = F32(prev_d_mean) if prev_d_mean != 0 else mean                       # inference.py:248-251
    stats = (noise.astype(F32) * std + mu).astype(F32) if noise is not None else np.full(n, mu, F32)
    d = ((d * F32(1.0 - t)).astype(F32) + (stats * F32(t)).astype(F32)).astype(F32)      # inference.py:252
    d[1:-2] = replace_outliers_zscore(d[1:-2])                                # inference.py:253
    d = (d / F32(speed)).astype(F32)                                          # inference.py:255
    return d, F32(d.astype(np.float64).mean())


def alignment_matrix(pred_dur: np.ndarray) -> np.ndarray:
    """inference.py:258-262: one-hot [L,F], row i is 1 on [c_i, c_i + dur_i)."""
    L = pred_dur.shape[0]
    F = int(pred_dur.sum())
    A = np.zeros((L, F), F32)
    c = 0
    for i in range(L):
        A[i, c:c + int(pred_dur[i])] = 1
        c += int(pred_dur[i])
    return A


def length_regulate(src: np.ndarray, pred_dur: np.ndarray) -> np.ndarray:
    """inference.py:266,268: `src @ alignment` for one utterance, src [C,L] -> [C,F].
    Stated as the reference does (matmul with the one-hot matrix) so that this oracle
    is an independent check of the gather kernel."""
    return (src.astype(F32) @ alignment_matrix(pred_dur)).astype(F32)


def length_regulate_batch(src: np.ndarray, pred_dur: np.ndarray, Fmax: Optional[int] = None):
    """Batched form (ONNX/inference_onnx.py:155-175 is the reference's vectorised twin):
    src [B,C,L], pred_dur [B,L] (0 for padded tokens) -> [B,C,Fmax], zero beyond each
    utterance's own frame count."""
    B, C, L = src.shape
    tot = pred_dur.sum(1)
    Fmax = int(tot.max()) if Fmax is None else Fmax
    out = np.zeros((B, C, Fmax), F32)
    for b in range(B):
        d = pred_dur[b]
        o = length_regulate(src[b], d)
        out[b, :, :o.shape[1]] = o
    return out
