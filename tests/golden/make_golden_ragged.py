"""Golden fixtures for PADDED batches of the text modules, from the UNMODIFIED reference.
Authoring container only (needs /root/reference):

    python tests/golden/make_golden_ragged.py

The reference synthesises one sentence per call (inference.py:231-245), but its modules take padded batches: TextEncoder.forward
masks with m and packs the LSTM (models.py:258-285), DurationEncoder.forward likewise (models.py:485-520), and
ProsodyPredictor.forward packs predictor.lstm (models.py:426-439).  Here those same modules run on a batch of three utterances of
9, 6 and 4 tokens padded to 9, and every utterance is ALSO run alone through the inference.py call sequence (B = 1, its own
length): the fixture keeps both, and the script asserts that the padded-batch result equals the one-at-a-time result on the
valid tokens -- which is what makes a padded batch a drop-in for the reference's sentence loop.

Fixtures
  text_ragged_B3_L9_w0_i5101.npz   tokens, lengths, out [3, 512, 9] (padded columns zero), single-utterance outputs
  dur_ragged_B3_L9_w0_i4101.npz    t_en, s, lengths, d [3, 9, 640], duration [3, 9], tap of text_encoder.lstms.0 and of lstm
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_predictor import build_full_predictor, synth, PredictorConfig  # noqa: E402  (sets up sys.path + the munch shim)

LENGTHS = [9, 6, 4]


def text_fixture():
    from models import TextEncoder
    from styletts2_lite_b200.config import TextEncoderConfig
    tsd = synth.make_text_state_dict(TextEncoderConfig(), seed=0, perturb=True)
    te = TextEncoder(channels=512, kernel_size=5, depth=3, n_symbols=178)                 # models.py:563
    te.load_state_dict(tsd)
    te = te.eval()
    B, L = len(LENGTHS), max(LENGTHS)
    tok = synth.make_tokens(B, L, seed=5101)
    lengths = torch.tensor(LENGTHS, dtype=torch.long)
    for b, n in enumerate(LENGTHS):
        tok[b, n:] = 0                                                                    # the pad id of a collated batch
    taps = {}
    hk = te.cnn[0].register_forward_hook(lambda mod, i, o: taps.__setitem__("cnn.0", o.detach().clone()))
    with torch.no_grad():
        m = te.length_to_mask(lengths)
        out = te(tok, lengths, m)                                                         # padded batch, models.py:258-285
        hk.remove()
        singles = []
        for b, n in enumerate(LENGTHS):                                                   # inference.py:235-239, one sentence
            ln = torch.tensor([n], dtype=torch.long)
            singles.append(te(tok[b:b + 1, :n], ln, te.length_to_mask(ln)))
    for b, n in enumerate(LENGTHS):
        assert torch.allclose(out[b, :, :n], singles[b][0], atol=1e-5), (b, float((out[b, :, :n] - singles[b][0]).abs().max()))
        assert float(out[b, :, n:].abs().max()) == 0.0 if n < L else True
    # the hook sees the block's output BEFORE the masked_fill_ of models.py:266 touches it in place -- but masked_fill_ is in
    # place on the same tensor, so the clone above is the unmasked value; mask it here as the reference does next
    t0 = taps["cnn.0"].masked_fill(m.unsqueeze(1), 0.0)
    np.savez_compressed(os.path.join(HERE, "text_ragged_B3_L9_w0_i5101.npz"), tokens=tok.numpy(), lengths=np.array(LENGTHS, np.int32),
                        out=out.numpy(), **{"tap:cnn.0": t0.numpy()},
                        **{"single%d" % b: singles[b].numpy() for b in range(B)})
    print("text ragged", out.shape, float(out.abs().max()))


def duration_fixture():
    sd = synth.make_predictor_state_dict(PredictorConfig(), seed=0, perturb=True, duration=True)
    p = build_full_predictor(sd)
    B, L = len(LENGTHS), max(LENGTHS)
    inp = synth.make_duration_inputs(B, L, seed=4101)
    t_en, s = inp["t_en"], inp["s"]
    lengths = torch.tensor(LENGTHS, dtype=torch.long)
    taps = {}
    mods = dict(p.named_modules())

    def grab(o):
        o = o[0] if isinstance(o, tuple) else o
        if isinstance(o, torch.nn.utils.rnn.PackedSequence):
            o = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=L)[0]
        return o.detach().clone()

    hooks = [mods[n].register_forward_hook(lambda mod, i, o, n=n: taps.__setitem__(n, grab(o)))
             for n in ("text_encoder.lstms.0", "lstm")]
    with torch.no_grad():
        m = p.length_to_mask(lengths)
        d = p.text_encoder(t_en, s, lengths, m)                                            # models.py:423
        x = torch.nn.utils.rnn.pack_padded_sequence(d, lengths, batch_first=True, enforce_sorted=False)   # models.py:426-428
        x, _ = p.lstm(x)                                                                   # models.py:432-433
        x, _ = torch.nn.utils.rnn.pad_packed_sequence(x, batch_first=True)                 # models.py:434-435
        x_pad = torch.zeros([x.shape[0], m.shape[-1], x.shape[-1]])                        # models.py:437-440
        x_pad[:, :x.shape[1], :] = x
        duration = torch.sigmoid(p.duration_proj(x_pad)).sum(axis=-1)                      # models.py:442 / inference.py:245
        for h in hooks:
            h.remove()
        sd_, sdur = [], []
        for b, n in enumerate(LENGTHS):                                                    # inference.py:242-245, one sentence
            ln = torch.tensor([n], dtype=torch.long)
            d1 = p.text_encoder(t_en[b:b + 1, :, :n], s[b:b + 1], ln, p.length_to_mask(ln))
            x1, _ = p.lstm(d1)
            sd_.append(d1)
            sdur.append(torch.sigmoid(p.duration_proj(x1)).sum(axis=-1))
    for b, n in enumerate(LENGTHS):
        assert torch.allclose(d[b, :n], sd_[b][0], atol=1e-5), (b, float((d[b, :n] - sd_[b][0]).abs().max()))
        assert torch.allclose(duration[b, :n], sdur[b][0], atol=1e-4), (b, float((duration[b, :n] - sdur[b][0]).abs().max()))
        if n < L:
            assert float(d[b, n:].abs().max()) == 0.0
    np.savez_compressed(os.path.join(HERE, "dur_ragged_B3_L9_w0_i4101.npz"), t_en=t_en.numpy(), s=s.numpy(),
                        lengths=np.array(LENGTHS, np.int32), d=d.numpy(), duration=duration.numpy(),
                        **{"tap:" + k: v.numpy() for k, v in taps.items()},
                        **{"single_dur%d" % b: sdur[b].numpy() for b in range(B)})
    print("duration ragged", d.shape, duration.numpy().round(3))


if __name__ == "__main__":
    torch.set_num_threads(8)
    text_fixture()
    duration_fixture()
