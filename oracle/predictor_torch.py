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
