"""Decoder hyper-parameters and the reference state_dict schema.

The two shipped shape families are the `decoder:` blocks of the reference's
Configs/config_example.yaml:59-64 (hifigan) and :66-73 (istftnet).  The
parameter names / shapes enumerated by `param_specs` are exactly what the
reference's `Decoder.state_dict()` holds (Modules/hifigan.py:416-443,
Modules/istftnet.py:660-690): legacy weight-norm pairs `X.weight_g`/`X.weight_v`,
AdaIN `fc` linears, Snake `alpha` lists, the NSF source linear and (istftnet)
the five CustomSTFT buffers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

SAMPLE_RATE = 24000          # hifigan.py:280 / istftnet.py:507
HARMONICS = 9                # harmonic_num=8 -> dim 9 (hifigan.py:106,282)
VOICED_THRESHOLD = 10.0      # hifigan.py:282
SINE_AMP = 0.1               # hifigan.py:99
NOISE_STD = 0.003            # hifigan.py:99
STYLE_DIM = 128              # config_example.yaml:43
HIDDEN_DIM = 512             # config_example.yaml:38


@dataclass
class DecoderConfig:
    type: str = "hifigan"                       # 'hifigan' | 'istftnet' | 'vocos'
    dim_in: int = HIDDEN_DIM
    style_dim: int = STYLE_DIM
    resblock_kernel_sizes: List[int] = field(default_factory=lambda: [3, 7, 11])
    upsample_rates: List[int] = field(default_factory=lambda: [10, 5, 3, 2])
    upsample_initial_channel: int = 512
    resblock_dilation_sizes: List[List[int]] = field(
        default_factory=lambda: [[1, 3, 5], [1, 3, 5], [1, 3, 5]])
    upsample_kernel_sizes: List[int] = field(default_factory=lambda: [20, 10, 6, 4])
    gen_istft_n_fft: int = 20
    gen_istft_hop_size: int = 5
    intermediate_dim: int = 1536                # vocos only (config_example.yaml:76-77)
    num_layers: int = 8

    @staticmethod
    def vocos() -> "DecoderConfig":
        """config_example.yaml:75-79: ConvNeXt backbone + ISTFTHead(n_fft 1200, hop 300) behind the shared front half."""
        return DecoderConfig(type="vocos", upsample_rates=[], upsample_kernel_sizes=[], resblock_kernel_sizes=[],
                             resblock_dilation_sizes=[], gen_istft_n_fft=1200, gen_istft_hop_size=300)

    @property
    def is_vocos(self) -> bool:
        return self.type == "vocos"

    @staticmethod
    def hifigan() -> "DecoderConfig":
        return DecoderConfig()

    @staticmethod
    def istftnet() -> "DecoderConfig":
        return DecoderConfig(type="istftnet", upsample_rates=[10, 6],
                             upsample_kernel_sizes=[20, 12])

    @property
    def is_istft(self) -> bool:
        return self.type == "istftnet"

    @property
    def num_stages(self) -> int:
        return len(self.upsample_rates)

    @property
    def upsample_scale(self) -> int:
        """F0 frame -> sample factor (hifigan.py:281; istftnet.py:506)."""
        p = 1
        for u in self.upsample_rates:
            p *= u
        return p * (self.gen_istft_hop_size if (self.is_istft or self.is_vocos) else 1)

    @property
    def samples_per_frame(self) -> int:
        """Output samples per asr frame: two F0 frames per asr frame."""
        return 2 * self.upsample_scale

    def stage_channels(self, i: int) -> int:
        return self.upsample_initial_channel // (2 ** (i + 1))

    def noise_conv_geometry(self, i: int) -> Tuple[int, int, int, int]:
        """(cin, kernel, stride, padding) of generator.noise_convs[i]
        (hifigan.py:296-303, istftnet.py:526-533)."""
        cin = self.gen_istft_n_fft + 2 if self.is_istft else 1
        if i + 1 < self.num_stages:
            sf = 1
            for u in self.upsample_rates[i + 1:]:
                sf *= u
            return cin, 2 * sf, sf, (sf + 1) // 2
        return cin, 1, 1, 0

    def noise_res_kernel(self, i: int) -> int:
        return 7 if i + 1 < self.num_stages else 11

    def ups_geometry(self, i: int) -> Tuple[int, int, int, int]:
        """(kernel, stride, padding, output_padding) of generator.ups[i]
        (hifigan.py:292-294 vs istftnet.py:512-514)."""
        u, k = self.upsample_rates[i], self.upsample_kernel_sizes[i]
        if self.is_istft:
            return k, u, (k - u) // 2, 0
        return k, u, u // 2 + u % 2, u % 2


def _wn(specs, name, shape, bias=True, new_style=False):
    """weight-normed conv: g has shape [shape[0],1,1] (dim=0 norm).  new_style: the keys of
    torch.nn.utils.parametrizations.weight_norm (Modules/vocos.py:10): original0 = g, original1 = v."""
    gk, vk = (".parametrizations.weight.original0", ".parametrizations.weight.original1") if new_style else (".weight_g", ".weight_v")
    specs.append((name + gk, (shape[0], 1, 1), "g:" + name + vk))
    specs.append((name + vk, tuple(shape), "conv"))
    if bias:
        specs.append((name + ".bias", None, "bias:" + name + vk))


def _adain(specs, name, style_dim, c):
    specs.append((name + ".fc.weight", (2 * c, style_dim), "linear"))
    specs.append((name + ".fc.bias", (2 * c,), "linear_bias:%d" % style_dim))


def _adain_resblk1d(specs, name, cin, cout, style_dim, upsample, new_style=False):
    _wn(specs, name + ".conv1", (cout, cin, 3), new_style=new_style)
    _wn(specs, name + ".conv2", (cout, cout, 3), new_style=new_style)
    _adain(specs, name + ".norm1", style_dim, cin)
    _adain(specs, name + ".norm2", style_dim, cout)
    if cin != cout:
        _wn(specs, name + ".conv1x1", (cout, cin, 1), bias=False, new_style=new_style)
    if upsample:
        _wn(specs, name + ".pool", (cin, 1, 3), new_style=new_style)


def _adain_resblock1(specs, name, c, k, style_dim):
    for j in range(3):
        _wn(specs, "%s.convs1.%d" % (name, j), (c, c, k))
        _wn(specs, "%s.convs2.%d" % (name, j), (c, c, k))
        _adain(specs, "%s.adain1.%d" % (name, j), style_dim, c)
        _adain(specs, "%s.adain2.%d" % (name, j), style_dim, c)
        specs.append(("%s.alpha1.%d" % (name, j), (1, c, 1), "alpha"))
        specs.append(("%s.alpha2.%d" % (name, j), (1, c, 1), "alpha"))


def param_specs(cfg: DecoderConfig):
    """[(state_dict key, shape, init-kind)] for every *parameter* of the
    reference Decoder.  bias shapes are resolved from their conv (out dim:
    dim 0 for Conv1d, dim 1*groups for ConvTranspose1d)."""
    sd = cfg.style_dim
    specs = []
    ns = cfg.is_vocos                            # Modules/vocos.py uses the parametrizations weight_norm
    _adain_resblk1d(specs, "encode", cfg.dim_in + 2, 1024, sd, False, ns)
    for i in range(3):
        _adain_resblk1d(specs, "decode.%d" % i, 1024 + 2 + 64, 1024, sd, False, ns)
    _adain_resblk1d(specs, "decode.3", 1024 + 2 + 64, 512, sd, True, ns)
    _wn(specs, "F0_conv", (1, 1, 3), new_style=ns)
    _wn(specs, "N_conv", (1, 1, 3), new_style=ns)
    _wn(specs, "asr_res.0", (64, 512, 1), new_style=ns)
    g = "generator"
    if cfg.is_vocos:
        _vocos_generator(specs, cfg)
        return _resolve_bias_shapes(specs)
    specs.append((g + ".m_source.l_linear.weight", (1, HARMONICS), "linear"))
    specs.append((g + ".m_source.l_linear.bias", (1,), "linear_bias:%d" % HARMONICS))
    c0 = cfg.upsample_initial_channel
    if not cfg.is_istft:
        specs.append((g + ".alphas.0", (1, c0, 1), "alpha"))
    for i in range(cfg.num_stages):
        c = cfg.stage_channels(i)
        cin, k, _, _ = cfg.noise_conv_geometry(i)
        specs.append(("%s.noise_convs.%d.weight" % (g, i), (c, cin, k), "conv"))
        specs.append(("%s.noise_convs.%d.bias" % (g, i), (c,), "bias:%s.noise_convs.%d.weight" % (g, i)))
        _adain_resblock1(specs, "%s.noise_res.%d" % (g, i), c, cfg.noise_res_kernel(i), sd)
        ku = cfg.ups_geometry(i)[0]
        # ConvTranspose1d weight is [Cin, Cout, k]; weight_g is [Cin,1,1]
        _wn(specs, "%s.ups.%d" % (g, i), (2 * c, c, ku))
        if not cfg.is_istft:
            specs.append(("%s.alphas.%d" % (g, i + 1), (1, c, 1), "alpha"))
        for j, kr in enumerate(cfg.resblock_kernel_sizes):
            _adain_resblock1(specs, "%s.resblocks.%d" % (g, i * len(cfg.resblock_kernel_sizes) + j),
                             c, kr, sd)
    c_last = cfg.stage_channels(cfg.num_stages - 1)
    cpost = cfg.gen_istft_n_fft + 2 if cfg.is_istft else 1
    _wn(specs, g + ".conv_post", (cpost, c_last, 7))
    return _resolve_bias_shapes(specs)


def _vocos_generator(specs, cfg: DecoderConfig):
    """Generator of Modules/vocos.py:103-162: num_layers ConvNeXtBlocks (:27-69: depthwise Conv1d k=7, AdaIN1d, Linear dim ->
    intermediate, GELU, Linear back, layer scale gamma), final LayerNorm(dim, eps 1e-6), ISTFTHead.out = Linear(dim, n_fft + 2)."""
    dim, sd = cfg.dim_in, cfg.style_dim
    for i in range(cfg.num_layers):
        n = "generator.convnext.%d" % i
        specs.append((n + ".gamma", (dim,), "layer_scale:%g" % (1.0 / cfg.num_layers)))
        specs.append((n + ".dwconv.weight", (dim, 1, 7), "trunc02"))
        specs.append((n + ".dwconv.bias", (dim,), "small_bias"))
        specs.append((n + ".norm.fc.weight", (2 * dim, sd), "trunc02"))
        specs.append((n + ".norm.fc.bias", (2 * dim,), "small_bias"))
        specs.append((n + ".pwconv1.weight", (cfg.intermediate_dim, dim), "trunc02"))
        specs.append((n + ".pwconv1.bias", (cfg.intermediate_dim,), "small_bias"))
        specs.append((n + ".pwconv2.weight", (dim, cfg.intermediate_dim), "trunc02"))
        specs.append((n + ".pwconv2.bias", (dim,), "small_bias"))
    specs.append(("generator.final_layer_norm.weight", (dim,), "gamma"))
    specs.append(("generator.final_layer_norm.bias", (dim,), "beta"))
    specs.append(("generator.stft.out.weight", (cfg.gen_istft_n_fft + 2, dim), "linear"))
    specs.append(("generator.stft.out.bias", (cfg.gen_istft_n_fft + 2,), "linear_bias:%d" % dim))


def _resolve_bias_shapes(specs):
    shapes = {n: s for n, s, _ in specs if s is not None}
    out = []
    for n, s, kind in specs:
        if s is None:
            wname = kind.split(":", 1)[1]
            w = shapes[wname]
            transposed = (".ups." in n) or n.endswith(".pool.bias")
            if n.endswith(".pool.bias"):
                s = (w[0],)                      # depthwise: groups = Cin
            elif transposed:
                s = (w[1],)
            else:
                s = (w[0],)
        out.append((n, s, kind))
    return out


def buffer_specs(cfg: DecoderConfig):
    """CustomSTFT buffers registered by the istftnet Generator (istftnet.py:145-203); the Hann window of the vocos ISTFT
    (vocos.py:192-193)."""
    if cfg.is_vocos:
        return [("generator.stft.istft.window", (cfg.gen_istft_n_fft,))]
    if not cfg.is_istft:
        return []
    n, bins = cfg.gen_istft_n_fft, cfg.gen_istft_n_fft // 2 + 1
    p = "generator.stft."
    return [(p + "window", (n,)),
            (p + "weight_forward_real", (bins, 1, n)),
            (p + "weight_forward_imag", (bins, 1, n)),
            (p + "weight_backward_real", (bins, 1, n)),
            (p + "weight_backward_imag", (bins, 1, n))]


def is_transposed_conv(name: str) -> bool:
    """True for the ConvTranspose1d modules (generator.ups.*, decode.3.pool)."""
    return ".ups." in name or name.endswith(".pool") or ".pool." in name


# ---- §8(f) N1: the F0 / energy predictor that feeds the Decoder (models.py:448-461) -----------------
@dataclass
class PredictorConfig:
    """`ProsodyPredictor(style_dim, d_hid, ...)` as built at models.py:575 from config_example.yaml:38,43.
    Only the sub-modules `F0Ntrain` uses are covered: `shared`, `F0`, `N`, `F0_proj`, `N_proj`."""
    d_hid: int = HIDDEN_DIM
    style_dim: int = STYLE_DIM


F0N_PREFIXES = ("shared.", "F0.", "N.", "F0_proj.", "N_proj.")


def predictor_param_specs(cfg: PredictorConfig):
    """[(state_dict key, shape, init-kind)] of the parameters ProsodyPredictor.F0Ntrain reads
    (models.py:407-419: bidirectional nn.LSTM `shared`, 2 x 3 AdainResBlk1d, two 1x1 Conv1d)."""
    d, sd = cfg.d_hid, cfg.style_dim
    h = d // 2
    specs = []
    for suffix in ("", "_reverse"):
        specs.append(("shared.weight_ih_l0" + suffix, (4 * h, d + sd), "lstm:%d" % h))
        specs.append(("shared.weight_hh_l0" + suffix, (4 * h, h), "lstm:%d" % h))
        specs.append(("shared.bias_ih_l0" + suffix, (4 * h,), "lstm:%d" % h))
        specs.append(("shared.bias_hh_l0" + suffix, (4 * h,), "lstm:%d" % h))
    for br in ("F0", "N"):
        _adain_resblk1d(specs, br + ".0", d, d, sd, False)
        _adain_resblk1d(specs, br + ".1", d, h, sd, True)
        _adain_resblk1d(specs, br + ".2", h, h, sd, False)
    for br in ("F0_proj", "N_proj"):
        specs.append((br + ".weight", (1, h, 1), "conv"))
        specs.append((br + ".bias", None, "bias:" + br + ".weight"))
    shapes = {n: s for n, s, _ in specs if s is not None}
    out = []
    for n, s, kind in specs:
        if s is None:
            w = shapes[kind.split(":", 1)[1]]
            s = (w[0],)                              # Conv1d bias; depthwise pool: groups = Cin = w[0]
        out.append((n, s, kind))
    return out


DUR_PREFIXES = ("text_encoder.", "lstm.", "duration_proj.")


def _lstm(specs, name, d_in, h):
    for suffix in ("", "_reverse"):
        specs.append((name + ".weight_ih_l0" + suffix, (4 * h, d_in), "lstm:%d" % h))
        specs.append((name + ".weight_hh_l0" + suffix, (4 * h, h), "lstm:%d" % h))
        specs.append((name + ".bias_ih_l0" + suffix, (4 * h,), "lstm:%d" % h))
        specs.append((name + ".bias_hh_l0" + suffix, (4 * h,), "lstm:%d" % h))


def duration_param_specs(cfg: PredictorConfig, nlayers: int = 3, max_dur: int = 50):
    """§8(f) N2: the duration half of ProsodyPredictor -- DurationEncoder `text_encoder` (models.py:468-483: nlayers x
    [bidirectional LSTM, AdaLayerNorm]), `lstm` (models.py:404) and `duration_proj` = LinearNorm(d_hid, max_dur)
    (models.py:405, :152-162)."""
    d, sd = cfg.d_hid, cfg.style_dim
    specs = []
    for i in range(nlayers):
        _lstm(specs, "text_encoder.lstms.%d" % (2 * i), d + sd, d // 2)
        _adain(specs, "text_encoder.lstms.%d" % (2 * i + 1), sd, d)        # AdaLayerNorm.fc: Linear(style, 2*channels)
    _lstm(specs, "lstm", d + sd, d // 2)
    specs.append(("duration_proj.linear_layer.weight", (max_dur, d), "linear"))
    specs.append(("duration_proj.linear_layer.bias", (max_dur,), "linear_bias:%d" % d))
    return specs


# ---- §8(f) N3: TextEncoder (models.py:238-285) ---------------------------------------------------------------------
@dataclass
class TextEncoderConfig:
    """`TextEncoder(channels=hidden_dim, kernel_size=5, depth=n_layer, n_symbols=n_token)` (models.py:563,
    config_example.yaml:38-41; inference.py:82 sets n_token = 178)."""
    channels: int = HIDDEN_DIM
    kernel_size: int = 5
    depth: int = 3
    n_symbols: int = 178


def text_encoder_param_specs(cfg: TextEncoderConfig):
    """[(state_dict key, shape, init-kind)] of TextEncoder: nn.Embedding, depth x [weight-normed Conv1d, LayerNorm],
    bidirectional nn.LSTM(channels, channels // 2) (models.py:241-256)."""
    c = cfg.channels
    specs = [("embedding.weight", (cfg.n_symbols, c), "normal")]
    for i in range(cfg.depth):
        _wn(specs, "cnn.%d.0" % i, (c, c, cfg.kernel_size))
        specs.append(("cnn.%d.1.gamma" % i, (c,), "gamma"))
        specs.append(("cnn.%d.1.beta" % i, (c,), "beta"))
    _lstm(specs, "lstm", c, c // 2)
    shapes = {n: s for n, s, _ in specs if s is not None}
    return [(n, s if s is not None else (shapes[k.split(":", 1)[1]][0],), k) for n, s, k in specs]
