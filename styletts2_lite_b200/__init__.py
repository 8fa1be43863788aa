"""B200-native StyleTTS2-lite waveform decoder hot path."""
from .config import DecoderConfig  # noqa: F401

__all__ = ["DecoderConfig", "build_decoder"]


def build_decoder(cfg: DecoderConfig, precision: str = "fp32"):
    """Mirror of the class selection in models.py:538-561 / inference.py:95-111."""
    from .decoder import B200Decoder
    assert cfg.type in ("hifigan", "istftnet", "vocos"), "Decoder type unknown"
    return B200Decoder(cfg, precision)
