"""CPU oracle for SURVEY.md §8(f) N1, `ProsodyPredictor.F0Ntrain`  --  TEST INFRASTRUCTURE ONLY.

numpy restatement of the step right before the Decoder (reference models.py:448-461): the shared
bidirectional LSTM, two stacks of three AdainResBlk1d and the two 1x1 projections that produce
`F0_pred` / `N_pred` (inference.py:268).  Same rules as oracle/decoder_np.py: only tests/, smoke() and
bench.py's CPU legs may import it, never the product package.

Parity pinning: the reference has no tests or fixtures for this function either; the oracle is pinned against
the reference itself run in the authoring container (tests/golden/make_golden_predictor.py imports
/root/reference/models.py through a 4-line `munch` shim and commits tests/golden/f0n_*.npz).

The arithmetic of nn.LSTM lives in PyTorch ATen (torch/nn/modules/rnn.py + aten/native/RNN.cpp, torch 2.7.0 in
uv.lock:2116): gates = x W_ih^T + b_ih + h W_hh^T + b_hh, chunked (i, f, g, o);
c' = sigmoid(f) c + sigmoid(i) tanh(g); h' = sigmoid(o) tanh(c').  Layout follows the reference: [B, C, T].
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .decoder_np import F32, Weights, adain_resblk1d, conv1d


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool, lengths=None):
    """One direction of nn.LSTM(batch_first=True), zero initial state.  x [B, T, I] -> [B, T, H].
    lengths [B]: pack_padded_sequence -> LSTM -> pad_packed_sequence (models.py:271-277, :503-509, :426-435): utterance b is
    the LSTM over its first lengths[b] rows (the reverse direction starts at row lengths[b] - 1), later rows are zero."""
    B, T, _ = x.shape
    if lengths is not None:
        out = np.zeros((B, T, w_hh.shape[1]), F32)
        for b in range(B):
            n = int(lengths[b])
            out[b, :n] = lstm_direction(x[b:b + 1, :n], w_ih, w_hh, b_ih, b_hh, reverse)[0]
        return out
    H = w_hh.shape[1]
    gin = (x.reshape(B * T, -1).astype(F32) @ w_ih.T.astype(F32) + b_ih).reshape(B, T, 4 * H).astype(F32)
    h = np.zeros((B, H), F32)
    c = np.zeros((B, H), F32)
    out = np.zeros((B, T, H), F32)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = (gin[:, t] + (h @ w_hh.T.astype(F32) + b_hh)).astype(F32)
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = (_sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)).astype(F32)
        h = (_sigmoid(o) * np.tanh(c)).astype(F32)
        out[:, t] = h
    return out


def bilstm(W: Weights, name: str, x, lengths=None):
    """nn.LSTM(d_hid + style_dim, d_hid // 2, 1, batch_first=True, bidirectional=True) (models.py:407).
    x [B, T, I] -> [B, T, 2H] = cat(forward, reverse); lengths as in lstm_direction."""
    outs = []
    for suffix, rev in (("", False), ("_reverse", True)):
        outs.append(lstm_direction(x, W.p(name + ".weight_ih_l0" + suffix), W.p(name + ".weight_hh_l0" + suffix),
                                   W.p(name + ".bias_ih_l0" + suffix), W.p(name + ".bias_hh_l0" + suffix), rev, lengths))
    return np.concatenate(outs, axis=2)


def f0n_train(sd: Dict[str, np.ndarray], en, s, taps: Optional[dict] = None, operand: Optional[str] = None):
    """ProsodyPredictor.F0Ntrain(x, s) (models.py:448-461).  en [B, d_hid+style_dim, T], s [B, style_dim]
    -> (F0 [B, 2T], N [B, 2T]).  Dropout (p=0.2, models.py:409-416) is the identity in eval mode."""
    W = Weights(sd)
    en = np.asarray(en, F32)
    s = np.asarray(s, F32)
    x = bilstm(W, "shared", en.transpose(0, 2, 1))            # models.py:449
    if taps is not None:
        taps["shared"] = x
    outs = []
    for br in ("F0", "N"):
        h = x.transpose(0, 2, 1)                                # models.py:451 / :456
        for i, up in enumerate((False, True, False)):           # models.py:409-416 (`upsample=True` on block 1)
            h = adain_resblk1d(W, "%s.%d" % (br, i), h, s, up, taps=taps, operand=operand)
        h = conv1d(h, W.w(br + "_proj"), W.b(br + "_proj"))     # models.py:454 / :459
        outs.append(h[:, 0, :])                                 # .squeeze(1)
    return outs[0], outs[1]


# ----------------------------------------------------------------------------
# the duration half (inference.py:242-245)
# ----------------------------------------------------------------------------
def ada_layer_norm(x, s, fc_w, fc_b, eps=1e-5):
    """AdaLayerNorm.forward (models.py:372-392) on x [B, L, C]: LayerNorm over C (biased variance, eps 1e-5, no affine)
    then (1 + gamma) * x + beta with gamma | beta = fc(s) [B, 2C]."""
    h = (s.astype(F32) @ fc_w.T.astype(F32) + fc_b).astype(F32)
    C = x.shape[2]
    gamma, beta = h[:, None, :C], h[:, None, C:]
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=2, keepdims=True)
    var = x64.var(axis=2, keepdims=True)
    xn = ((x64 - mean) / np.sqrt(var + eps)).astype(F32)
    return ((1 + gamma) * xn + beta).astype(F32)


def _pad_mask(lengths, B, L):
    """length_to_mask (models.py:463-466): True behind each utterance; all False without lengths."""
    if lengths is None:
        return np.zeros((B, L), bool)
    return np.arange(L)[None, :] >= np.asarray(lengths).reshape(B, 1)


def duration_encoder(W: Weights, t_en, s, nlayers=3, taps: Optional[dict] = None, lengths=None):
    """DurationEncoder.forward (models.py:485-520).  t_en [B, d_hid, L], s [B, style] -> d [B, L, d_hid + style].
    lengths [B] (None = no padding: every mask is False, the pack / pad round trip is the identity)."""
    B, _, L = t_en.shape
    m = _pad_mask(lengths, B, L)
    sty = np.repeat(s[:, None, :], L, axis=1).astype(F32)                    # style broadcast over tokens (models.py:489)
    x = np.concatenate([t_en.transpose(0, 2, 1), sty], axis=2).astype(F32)   # [B, L, 640] (models.py:490)
    x[m] = 0                                                                 # models.py:491
    for i in range(nlayers):
        y = bilstm(W, "text_encoder.lstms.%d" % (2 * i), x, lengths)         # models.py:503-509
        if taps is not None:
            taps["text_encoder.lstms.%d" % (2 * i)] = y
        n = "text_encoder.lstms.%d" % (2 * i + 1)
        y = ada_layer_norm(y, s, W.p(n + ".fc.weight"), W.p(n + ".fc.bias"))  # models.py:498
        x = np.concatenate([y, sty], axis=2).astype(F32)                     # models.py:499
        x[m] = 0                                                             # models.py:500
    return x


def predict_duration(sd: Dict[str, np.ndarray], t_en, s, taps: Optional[dict] = None, lengths=None):
    """inference.py:242-245: d = predictor.text_encoder(t_en, s, lengths, mask); x, _ = predictor.lstm(d);
    duration = sigmoid(predictor.duration_proj(x)).sum(-1).  Returns (d [B, L, 640], duration [B, L]).
    With lengths (a padded batch) predictor.lstm is the packed one of ProsodyPredictor.forward (models.py:426-439), so a
    padded token's duration is sigmoid(duration_proj.bias).sum()."""
    W = Weights(sd)
    t_en = np.asarray(t_en, F32)
    s = np.asarray(s, F32)
    d = duration_encoder(W, t_en, s, taps=taps, lengths=lengths)
    x = bilstm(W, "lstm", d, lengths)
    if taps is not None:
        taps["lstm"] = x
    logits = (x @ W.p("duration_proj.linear_layer.weight").T.astype(F32) + W.p("duration_proj.linear_layer.bias")).astype(F32)
    duration = _sigmoid(logits).sum(axis=2).astype(F32)
    return d, duration


# ----------------------------------------------------------------------------
# TextEncoder (models.py:238-285), equal-length batches
# ----------------------------------------------------------------------------
def text_encoder(sd: Dict[str, np.ndarray], tokens, depth=3, taps: Optional[dict] = None, operand: Optional[str] = None,
                 lengths=None):
    """TextEncoder.forward(x, input_lengths, m) (models.py:258-285); lengths None = all-False mask: embedding -> depth x
    [weight-normed Conv1d(k=5, 'same') -> LayerNorm over channels (models.py:224-236) -> LeakyReLU(0.2) -> Dropout = id]
    -> bidirectional LSTM -> [B, channels, L]."""
    from .decoder_np import leaky_relu
    W = Weights(sd)
    x = W.p("embedding.weight")[np.asarray(tokens)]                          # [B, L, C]  (models.py:259)
    x = x.transpose(0, 2, 1).astype(F32)                                     # [B, C, L]
    m = _pad_mask(lengths, x.shape[0], x.shape[2])[:, None, :]               # models.py:261
    x = np.where(m, F32(0), x)                                               # models.py:262
    for i in range(depth):
        n = "cnn.%d" % i
        k = W.w(n + ".0").shape[2]
        x = conv1d(x, W.w(n + ".0"), W.b(n + ".0"), padding=(k - 1) // 2, operand=operand)
        x64 = x.astype(np.float64)
        mean, var = x64.mean(axis=1, keepdims=True), x64.var(axis=1, keepdims=True)
        xn = ((x64 - mean) / np.sqrt(var + 1e-5)).astype(F32)
        x = (xn * W.p(n + ".1.gamma")[None, :, None] + W.p(n + ".1.beta")[None, :, None]).astype(F32)
        x = leaky_relu(x, 0.2)
        x = np.where(m, F32(0), x)                                           # models.py:266
        if taps is not None:
            taps[n] = x
    y = bilstm(W, "lstm", x.transpose(0, 2, 1), lengths)                     # models.py:271-277
    return np.where(m, F32(0), y.transpose(0, 2, 1))                         # models.py:279-283
