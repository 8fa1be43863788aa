"""F0Ntrain on torch's own CPU kernels  --  TEST / BASELINE INFRASTRUCTURE ONLY (see decoder_torch.py).

The reference's ProsodyPredictor.F0Ntrain (models.py:448-461) executes ATen's LSTM, convolution and batch-norm
kernels on a CPU; this restates it over a flat state_dict with the same ops so that tools/bench_predictor.py can time
what the reference really runs on the host cores.  tests/test_oracle.py pins it to the reference fixtures.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .decoder_torch import TorchWeights, adain_resblk1d


def bilstm(W: TorchWeights, name: str, x: torch.Tensor) -> torch.Tensor:
    """nn.LSTM(640, 256, 1, batch_first=True, bidirectional=True) (models.py:407) through torch's functional LSTM."""
    flat = [W.p(name + "." + k + sfx) for sfx in ("", "_reverse") for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    H = flat[1].shape[1]
    z = x.new_zeros(2, x.shape[0], H)
    out, _, _ = torch.lstm(x, (z, z), flat, True, 1, 0.0, False, True, True)
    return out


def f0n_train(sd: Dict[str, torch.Tensor], en: torch.Tensor, s: torch.Tensor):
    """ProsodyPredictor.F0Ntrain(x, s), models.py:448-461."""
    W = sd if isinstance(sd, TorchWeights) else TorchWeights(sd)
    x = bilstm(W, "shared", en.transpose(-1, -2))
    outs = []
    for br in ("F0", "N"):
        h = x.transpose(-1, -2)
        for i, up in enumerate((False, True, False)):
            h = adain_resblk1d(W, "%s.%d" % (br, i), h, s, up)
        h = F.conv1d(h, W.w(br + "_proj"), W.b(br + "_proj"))
        outs.append(h.squeeze(1))
    return outs[0], outs[1]


def predict_duration(sd: Dict[str, torch.Tensor], t_en: torch.Tensor, s: torch.Tensor, nlayers: int = 3):
    """inference.py:242-245 for equal-length batches: DurationEncoder.forward (models.py:485-520), predictor.lstm,
    duration_proj, sigmoid-sum.  Returns (d [B, L, 640], duration [B, L])."""
    W = sd if isinstance(sd, TorchWeights) else TorchWeights(sd)
    L = t_en.shape[2]
    sty = s[:, None, :].expand(-1, L, -1)
    x = torch.cat([t_en.transpose(1, 2), sty], dim=2)
    for i in range(nlayers):
        y = bilstm(W, "text_encoder.lstms.%d" % (2 * i), x)
        n = "text_encoder.lstms.%d" % (2 * i + 1)
        h = F.linear(s, W.p(n + ".fc.weight"), W.p(n + ".fc.bias"))
        gamma, beta = torch.chunk(h[:, None, :], 2, dim=2)
        y = (1 + gamma) * F.layer_norm(y, (y.shape[2],), eps=1e-5) + beta
        x = torch.cat([y, sty], dim=2)
    z = bilstm(W, "lstm", x)
    dur = torch.sigmoid(F.linear(z, W.p("duration_proj.linear_layer.weight"), W.p("duration_proj.linear_layer.bias"))).sum(-1)
    return x, dur


def text_encoder(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, depth: int = 3):
    """TextEncoder.forward (models.py:258-285) for equal-length batches on torch's CPU kernels."""
    W = sd if isinstance(sd, TorchWeights) else TorchWeights(sd)
    x = F.embedding(tokens, W.p("embedding.weight")).transpose(1, 2)
    for i in range(depth):
        n = "cnn.%d" % i
        w = W.w(n + ".0")
        x = F.conv1d(x, w, W.b(n + ".0"), padding=(w.shape[2] - 1) // 2)
        x = F.layer_norm(x.transpose(1, -1), (x.shape[1],), W.p(n + ".1.gamma"), W.p(n + ".1.beta"), 1e-5).transpose(1, -1)
        x = F.leaky_relu(x, 0.2)
    return bilstm(W, "lstm", x.transpose(1, 2)).transpose(1, 2)
