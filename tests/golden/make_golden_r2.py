"""Round-2 fixtures from the UNMODIFIED reference (authoring container only: needs /root/reference):
    python tests/golden/make_golden_r2.py

  hifigan_B1_T400_w0_i1003        10 s, the longest single-utterance case of BASELINE.json (fp32 bar 1e-4 where SineGen phases
                                  reach 1e5 rad)
  hifigan_B1_T120_w3_i1007_plain  3 s with the reference's TRUE initialisation (see reference_init_stats.json)
  hifigan_B1_T203_w0_i1009        ragged length: no tile size of any kernel divides 203 frames
  istftnet_B1_T120_w0_i1006       3 s iSTFTNet
  istftnet_B1_T203_w0_i1008       ragged iSTFTNet
  reference_init_stats.json       std / abs-max of the parameters of a freshly constructed reference Decoder
                                  (torch.manual_seed(0)).  It shows that `init_weights` N(0, 0.01) (hifigan.py:37,47,318-319)
                                  never reaches the forward: under the legacy weight_norm the call rewrites the derived
                                  `.weight` tensor only, which the pre-forward hook recomputes from weight_g / weight_v, and those
                                  keep PyTorch's default U(+-1/sqrt(fan_in)) with g = ||v||.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402


def init_stats():
    import warnings
    warnings.simplefilter("ignore")
    from Modules.hifigan import Decoder
    cfg = DecoderConfig.hifigan()
    torch.manual_seed(0)
    m = Decoder(dim_in=cfg.dim_in, style_dim=cfg.style_dim, dim_out=80, resblock_kernel_sizes=cfg.resblock_kernel_sizes,
                upsample_rates=cfg.upsample_rates, upsample_initial_channel=cfg.upsample_initial_channel,
                resblock_dilation_sizes=cfg.resblock_dilation_sizes, upsample_kernel_sizes=cfg.upsample_kernel_sizes).eval()
    sd = m.state_dict()
    out = {}
    for k, v in sd.items():
        if v.numel() < 2:
            out[k] = {"shape": list(v.shape), "std": 0.0, "absmax": float(v.abs().max()), "mean": float(v.mean())}
        else:
            out[k] = {"shape": list(v.shape), "std": float(v.std()), "absmax": float(v.abs().max()), "mean": float(v.mean())}
    # the effective weight the forward uses: recomputed from (g, v), not the N(0, 0.01) draw
    c = m.generator.resblocks[0].convs1[0]
    c(torch.randn(1, 256, 20))
    out["__effective__generator.resblocks.0.convs1.0.weight"] = {"std": float(c.weight.std()), "absmax": float(c.weight.abs().max())}
    with open(os.path.join(HERE, "reference_init_stats.json"), "w") as f:
        json.dump(out, f, sort_keys=True)
    print("init stats:", len(out), "entries; resblocks.0.convs1.0.weight_v std %.4f absmax %.4f" %
          (out["generator.resblocks.0.convs1.0.weight_v"]["std"], out["generator.resblocks.0.convs1.0.weight_v"]["absmax"]))


def main():
    torch.set_num_threads(8)
    init_stats()
    hg, ig = DecoderConfig.hifigan(), DecoderConfig.istftnet()
    for name, cfg, B, T, ws, iseed, perturb in (
            ("hifigan_B1_T120_w3_i1007_plain", hg, 1, 120, 3, 1007, False),
            ("hifigan_B1_T203_w0_i1009", hg, 1, 203, 0, 1009, True),
            ("istftnet_B1_T120_w0_i1006", ig, 1, 120, 0, 1006, True),
            ("istftnet_B1_T203_w0_i1008", ig, 1, 203, 0, 1008, True),
            ("hifigan_B1_T400_w0_i1003", hg, 1, 400, 0, 1003, True)):
        out, _, cap, _ = MG.run_case(cfg, B, T, ws, iseed, perturb)
        d = {"out": out.numpy()}
        if "phase" in cap:
            d["phase_sha256"] = np.array(MG.sha(cap["phase"].numpy()))
            d["phase_absmax"] = np.array(np.abs(cap["phase"].numpy()).max())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, tuple(out.shape), "absmax %.4f" % float(out.abs().max()), flush=True)


if __name__ == "__main__":
    main()
