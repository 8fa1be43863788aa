"""Stage the UNMODIFIED reference decoder modules under baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot)
so that `bench.py --impl reference` and the cpu_baseline / eager_gpu legs can time the reference's own code:

    python oracle/stage_reference.py            # authoring container only: needs /root/reference

Copies Modules/{__init__,hifigan,istftnet,vocos,utils}.py byte for byte (they are pure Python over torch / numpy / scipy, all present
on the GPU box) and writes their SHA-256 next to them.  Nothing is staged into tracked paths; the repo never contains reference
sources.  TEST / BASELINE INFRASTRUCTURE ONLY -- nothing in styletts2_lite_b200/ imports baseline/_ref."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/Modules"
DST = os.path.join(ROOT, "baseline", "_ref", "Modules")
FILES = ("__init__.py", "hifigan.py", "istftnet.py", "vocos.py", "utils.py")


def stage() -> bool:
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    digests = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        digests[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    with open(os.path.join(DST, "SHA256.json"), "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)
    return True


def load_reference_decoder(cfg):
    """The staged reference Decoder class for `cfg` (hifigan / istftnet / vocos), or None when nothing is staged."""
    base = os.path.join(ROOT, "baseline", "_ref")
    name = "vocos" if getattr(cfg, "is_vocos", False) else ("istftnet" if cfg.is_istft else "hifigan")
    if not os.path.exists(os.path.join(base, "Modules", name + ".py")):
        return None
    if base not in sys.path:
        sys.path.insert(0, base)
    import importlib
    return importlib.import_module("Modules." + name).Decoder


def build_reference(cfg, state_dict):
    """Reference Decoder(**config_example.yaml decoder block) with `state_dict` loaded through its own load_state_dict."""
    import warnings
    Decoder = load_reference_decoder(cfg)
    if Decoder is None:
        return None
    if getattr(cfg, "is_vocos", False):
        kw = dict(dim_in=cfg.dim_in, style_dim=cfg.style_dim, dim_out=80, intermediate_dim=cfg.intermediate_dim,
                  num_layers=cfg.num_layers, gen_istft_n_fft=cfg.gen_istft_n_fft, gen_istft_hop_size=cfg.gen_istft_hop_size)
    else:
        kw = dict(dim_in=cfg.dim_in, style_dim=cfg.style_dim, dim_out=80, resblock_kernel_sizes=cfg.resblock_kernel_sizes,
                  upsample_rates=cfg.upsample_rates, upsample_initial_channel=cfg.upsample_initial_channel,
                  resblock_dilation_sizes=cfg.resblock_dilation_sizes, upsample_kernel_sizes=cfg.upsample_kernel_sizes)
    if cfg.is_istft:
        kw.update(gen_istft_n_fft=cfg.gen_istft_n_fft, gen_istft_hop_size=cfg.gen_istft_hop_size)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Decoder(**kw)
        m.load_state_dict(state_dict)
    return m.eval()


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here: nothing staged")
