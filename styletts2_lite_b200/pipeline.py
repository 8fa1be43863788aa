"""Token ids -> waveform for the sentences of one text: the device-side mirror of `StyleTTS2.generate` / `__inference`
(inference.py:224-272, :303-319) after the phonemiser / text cleaner and with a given style vector.

The reference walks the sentences one at a time.  Everything before the length regulator is independent per sentence apart from one
scalar (the mean duration of sentence i - 1 steadies sentence i, inference.py:248-252, :312-313), so here the B sentences run as ONE
padded batch through the TextEncoder, the duration half of the predictor (both with the reference's masking / packed-sequence
semantics for padded batches), the chained duration smoothing and the rounding.  From there on the reference's modules have no
length masking (the InstanceNorms of F0Ntrain and of the Decoder run over whatever frames they are given), so each sentence is
regulated, sent through F0Ntrain and decoded with its own frame count, exactly as the reference does (optionally with graph
replay for frame counts that repeat).  The waveforms are then trimmed / concatenated / padded / peak-normalised / quantised on the device
(postprocess.assemble, inference.py:314-319 + Demo/infer.py:51-54).

Every step is one of the C-ABI calls of include/st2_b200.h; there is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import length_regulator as LR
from . import postprocess as PP


class B200Synthesizer:
    """text_encoder: B200TextEncoder, predictor: B200F0NPredictor(duration=True), decoder: B200Decoder (any variant); all on the
    same CUDA device, eval mode."""

    def __init__(self, text_encoder, predictor, decoder, precision: Optional[str] = None, cuda_graph: bool = False):
        """cuda_graph=True replays one captured graph per (frame count, precision) for the per-sentence forwards: worth it when
        frame counts repeat (the caches hold 8 shapes each); free-running text rarely repeats them, so the default is eager."""
        self.text_encoder, self.predictor, self.decoder = text_encoder, predictor, decoder
        self.precision = precision
        self.cuda_graph = cuda_graph

    @torch.no_grad()
    def infer_sentences(self, sentences: Sequence[torch.Tensor], s: torch.Tensor, speed: float = 1.0, t: float = 0.2,
                        prev_d_mean: float = 0.0, duration_noise: Optional[torch.Tensor] = None,
                        decoder_seeds: Optional[Sequence[int]] = None) -> Tuple[List[torch.Tensor], torch.Tensor, torch.Tensor]:
        """sentences: token-id tensors [L_i] (with the leading / trailing pad token 0 of inference.py:230-231); s [128] or [1,128]
        the style (inference.py:240).  duration_noise [B, Lmax]: the N(0,1) tape for the duration smoothing (drawn on the device
        when omitted and t > 0).  Returns (waveforms: B tensors [600 * frames_i], pred_dur [B, Lmax] int32, mean durations [B])."""
        dev = s.device
        B = len(sentences)
        lens = torch.tensor([int(x.numel()) for x in sentences], dtype=torch.int64)
        L = int(lens.max())
        tokens = torch.zeros(B, L, dtype=torch.int64)
        for b, x in enumerate(sentences):
            tokens[b, :x.numel()] = x.reshape(-1).to("cpu")
        tokens = tokens.to(dev)
        s1 = s.reshape(1, -1).float()
        sB = s1.expand(B, -1).contiguous()
        t_en = self.text_encoder(tokens, lens, precision=self.precision)                       # inference.py:239
        d, duration = self.predictor.predict_duration(t_en, sB, precision=self.precision, input_lengths=lens)   # inference.py:242-245
        duration, means = LR.smooth_durations(duration, duration_noise, t=t, speed=speed, prev_d_mean=prev_d_mean,
                                              n_tokens=lens, chained=True)                      # inference.py:248-255 (+ :312-313)
        pred_dur, frames = LR.round_durations(duration, lens)                                   # inference.py:257
        frames_h = [int(f) for f in frames.cpu()]                                               # the one host sync: frame counts
        Fmax = max(frames_h)
        asr = LR.length_regulate(t_en, pred_dur, Fmax)                                          # inference.py:268
        en = LR.length_regulate(d.transpose(1, 2).contiguous(), pred_dur, Fmax)                 # inference.py:266
        waves = []
        for b in range(B):                                                                      # no length masking downstream
            F = frames_h[b]
            en_b, asr_b = en[b:b + 1, :, :F].contiguous(), asr[b:b + 1, :, :F].contiguous()
            f0, n = self.predictor.F0Ntrain(en_b, s1, precision=self.precision, cuda_graph=self.cuda_graph)   # inference.py:267
            seed = None if decoder_seeds is None else int(decoder_seeds[b])
            w = self.decoder(asr_b, f0, n, s1, seed=seed, precision=self.precision, cuda_graph=self.cuda_graph)   # inference.py:270
            waves.append(w.reshape(-1))
        return waves, pred_dur, means

    @torch.no_grad()
    def generate(self, sentences: Sequence[torch.Tensor], s: torch.Tensor, speed: float = 1.0, stabilize: bool = True,
                 duration_noise: Optional[torch.Tensor] = None, decoder_seeds: Optional[Sequence[int]] = None,
                 trim: int = 4000, pad: int = 4000):
        """StyleTTS2.generate (inference.py:303-319) + the normalisation / PCM_16 conversion of Demo/infer.py:51-54 for the
        already-tokenised sentences of one text.  Returns (float64 waveform as handed to soundfile, int16 PCM), both on the device."""
        waves, _, _ = self.infer_sentences(sentences, s, speed=speed, t=0.2 if stabilize else 0.0,       # inference.py:304-305
                                           duration_noise=duration_noise, decoder_seeds=decoder_seeds)
        lengths = torch.tensor([w.numel() for w in waves], dtype=torch.int32, device=s.device)
        S = int(lengths.max())
        batch = torch.zeros(len(waves), S, dtype=torch.float32, device=s.device)
        for b, w in enumerate(waves):
            batch[b, :w.numel()] = w
        return PP.assemble(batch, lengths, trim=trim, pad=pad, want_float=True)
