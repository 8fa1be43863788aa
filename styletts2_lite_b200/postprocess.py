"""Host side of SURVEY.md 8(f) N4: the reference's waveform post-processing and wire format on the device.

Mirrors what follows the decoder in the reference: `StyleTTS2.generate` (inference.py:314-319 -- trim 4000 samples at both ends of
every sentence, concatenate, pad 4000 zeros) and the demo (Demo/infer.py:51-54 -- peak-normalise, write 24 kHz PCM_16).
`assemble()` is the drop-in for those lines when the sentences were decoded as ONE batch on the GPU: it returns the float64
utterance `r` the demo hands to soundfile and its PCM_16 samples, both computed by `st2_postprocess` (csrc/postprocess.cu),
bit-exact to the numpy restatement.  There is no CPU fallback."""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import torch

from . import _lib

TRIM = 4000      # inference.py:315
PAD = 4000       # inference.py:319
SAMPLE_RATE = 24000


def assemble(wav: torch.Tensor, lengths: Optional[torch.Tensor] = None, trim: int = TRIM, pad: int = PAD,
             want_float: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """wav [B, 1, S] or [B, S] fp32 CUDA tensor (one decoder batch, sentence i valid on its first lengths[i] samples).
    Returns (r float64 [N] or None, pcm int16 [N]) with N = sum(max(len_i - 2*trim, 0)) + 2*pad."""
    lib = _lib.load()
    if not wav.is_cuda:
        raise _lib.St2Error("postprocess.assemble needs CUDA tensors (there is no CPU path)")
    if wav.dim() == 3:
        wav = wav[:, 0, :]
    if wav.dim() != 2 or wav.dtype != torch.float32:
        raise ValueError("wav must be fp32 [B, 1, S] or [B, S]")
    wav = wav.contiguous()
    B, S = wav.shape
    if lengths is not None:
        if lengths.shape != (B,):
            raise ValueError("lengths must be [B]")
        lengths = lengths.to(device=wav.device, dtype=torch.int32).contiguous()
        if B and int(lengths.max()) > S:
            raise ValueError("a sentence length exceeds the waveform tensor")
    nmax = _lib.check(lib.st2_postprocess_max_samples(B, S, trim, pad), "postprocess_max_samples")
    dev = wav.device
    r = torch.empty(nmax, dtype=torch.float64, device=dev) if want_float else None
    pcm = torch.empty(nmax, dtype=torch.int16, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(_lib.check(lib.st2_postprocess_scratch_bytes(B)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.st2_postprocess(_lib.ptr(wav), _lib.ptr(lengths), B, S, trim, pad, _lib.ptr(r), _lib.ptr(pcm),
                                       _lib.ptr(total), _lib.ptr(scratch), stream), "postprocess")
    n = int(total.item())
    return (None if r is None else r[:n]), pcm[:n]


def wav_bytes(pcm: torch.Tensor, sample_rate: int = SAMPLE_RATE) -> bytes:
    """The 24 kHz mono PCM_16 RIFF/WAVE file soundfile.write produces for these samples (Demo/infer.py:53-54)."""
    data = pcm.to(torch.int16).cpu().numpy().astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate,
                                                                                      sample_rate * 2, 2, 16)
    return hdr + b"data" + struct.pack("<I", len(data)) + data
