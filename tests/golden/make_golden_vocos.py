"""Golden fixtures for the Vocos decoder variant (Modules/vocos.py:364-422), from the UNMODIFIED reference.
Authoring container only (needs /root/reference):

    python tests/golden/make_golden_vocos.py

`Decoder(dim_in=512, style_dim=128, dim_out=80, intermediate_dim=1536, num_layers=8, gen_istft_n_fft=1200,
gen_istft_hop_size=300)` as models.py:555-562 builds it from config_example.yaml:75-79, loaded (strict) with
styletts2_lite_b200.synth.make_state_dict(DecoderConfig.vocos(), seed 0, perturb) -- the variant has no random draw, so no tape.

Fixtures
  vocos_B2_T6_w0_i1000.npz     output [2,1,3600] + taps: decode.3 output, convnext.0 / .7 outputs, final LayerNorm, ISTFTHead.out
  vocos_B1_T120_w0_i1001.npz   3 s (cfg 1 size), output only
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200 import synth  # noqa: E402


def build():
    warnings.simplefilter("ignore")
    from Modules.vocos import Decoder
    cfg = DecoderConfig.vocos()
    d = Decoder(dim_in=cfg.dim_in, style_dim=cfg.style_dim, dim_out=80, intermediate_dim=cfg.intermediate_dim,
                num_layers=cfg.num_layers, gen_istft_n_fft=cfg.gen_istft_n_fft, gen_istft_hop_size=cfg.gen_istft_hop_size)
    sd = synth.make_state_dict(cfg, 0, True)
    ref = d.state_dict()
    assert set(ref) == set(sd), set(ref) ^ set(sd)
    d.load_state_dict(sd)
    return cfg, d.eval()


def run(cfg, d, B, T, seed, tap_names=()):
    inp = synth.make_inputs(B, T, seed, cfg, with_noise=False)
    taps, hooks = {}, []
    mods = dict(d.named_modules())
    for n in tap_names:
        hooks.append(mods[n].register_forward_hook(lambda mod, i, o, n=n: taps.__setitem__(n, o.detach().clone())))
    with torch.no_grad():
        out = d(inp["asr"], inp["F0_curve"], inp["N"], inp["s"])
    for h in hooks:
        h.remove()
    return out, taps


def main():
    torch.set_num_threads(8)
    cfg, d = build()
    out, taps = run(cfg, d, 2, 6, 1000, ["decode.3", "generator.convnext.0", "generator.convnext.7", "generator.final_layer_norm",
                                          "generator.stft.out"])
    np.savez_compressed(os.path.join(HERE, "vocos_B2_T6_w0_i1000.npz"), out=out.numpy(),
                        **{"tap:" + k: v.numpy() for k, v in taps.items()})
    print("vocos small", out.shape, float(out.abs().max()))
    out, _ = run(cfg, d, 1, 120, 1001)
    np.savez_compressed(os.path.join(HERE, "vocos_B1_T120_w0_i1001.npz"), out=out.numpy())
    print("vocos 3 s", out.shape, float(out.abs().max()), float(out.std()))
    import json
    path = os.path.join(HERE, "state_dict_schema.json")
    schema = json.load(open(path))
    sd = d.state_dict()
    schema["vocos"] = {"num_params": sum(v.numel() for k, v in sd.items() if not k.endswith("istft.window")),
                       "state_dict": {k: list(v.shape) for k, v in sd.items()}}
    json.dump(schema, open(path, "w"), sort_keys=True)


if __name__ == "__main__":
    main()
