"""Generate golden input/output vectors by running the REFERENCE itself (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

/root/reference is imported unmodified; its only missing dependency
(rotary-embedding-torch==0.6.1, not installable offline) is satisfied by registering
oracle/rope.py's stand-in as ``rotary_embedding_torch`` -- so every vector here is pinned
to the reference EXCEPT the RoPE arithmetic, which stays "parity unpinned" (oracle/rope.py).
The reference cannot travel to the GPU box, so the small vectors are committed.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import rope as _rope  # noqa: E402

shim = types.ModuleType("rotary_embedding_torch")
shim.RotaryEmbedding = _rope.RotaryEmbedding
sys.modules["rotary_embedding_torch"] = shim

from models.mss_tflocoformer import TFLocoformerMSS  # noqa: E402
from standalone.tflocoformer_separator import TFLocoformerSeparator  # noqa: E402
from standalone.bslocoformer_separator import BSLocoformerSeparator  # noqa: E402


def perturb(model, seed=1):
    """Give every 1-D parameter (gamma, biases, gLN affine) a non-trivial value."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim == 1 and not name.endswith("rope.freqs"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def mixture(n_samples, batch, seed=1234):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples) / 44100.0
    x = 0.1 * torch.randn(batch, n_samples, generator=g)
    for f0 in (55.0, 220.0, 880.0, 3520.0, 7040.0):
        x = x + 0.05 * torch.sin(2 * torch.pi * f0 * t)[None]
    return x.clamp(-1, 1)


def save(name, cfg, model, arrays):
    out = {"config": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.detach().numpy()
    for k, v in arrays.items():
        out[k] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def keep_in(n):
    return n.endswith("attn_norm") or ".ffn." in n


def capture(model, names):
    store, hooks = {}, []
    mods = dict(model.named_modules())
    for n in names:
        def hook(mod, inp, out, n=n):
            if keep_in(n):
                store[n + ":in"] = inp[0].detach().clone()
            store[n + ":out"] = out.detach().clone()
        hooks.append(mods[n].register_forward_hook(hook))
    return store, hooks


def mss_case(name, cfg, n_samples, batch, seed):
    torch.manual_seed(seed)
    model = TFLocoformerMSS(**cfg).eval()
    perturb(model, seed + 1)
    x = mixture(n_samples, batch, 1234 + seed)
    stage_names = ["conv", "blocks.0", "blocks.0.freq_path.attn_norm", "blocks.0.freq_path.attn",
                   "blocks.0.frame_path.ffn.0", "deconv"]
    store, hooks = capture(model, stage_names)
    with torch.no_grad():
        td = model(x, return_time_domain=True)
        for h in hooks:
            h.remove()
        sp = model(x, return_time_domain=False) if cfg["n_sources"] >= 4 else None
        ref_stft = model.transform.stft(x)
    arrays = {"mixture": x, "stft": ref_stft}
    for k, v in td.items():
        arrays["out/" + k] = v
    if sp is not None:
        for k, v in sp.items():
            arrays["spec/" + k] = v
    for k, v in store.items():
        arrays["stage/" + k] = v.contiguous()
    save(name, cfg, model, arrays)


def sep_case(name, cfg, shape, seed):
    torch.manual_seed(seed)
    model = TFLocoformerSeparator(**cfg).eval()
    perturb(model, seed + 1)
    g = torch.Generator().manual_seed(99 + seed)
    x = torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g))
    with torch.no_grad():
        y = model(x)
    save(name, cfg, model, {"spec_in": x, "spec_out": y})


def bs_case(name, cfg, shape, seed):
    torch.manual_seed(seed)
    model = BSLocoformerSeparator(**cfg).eval()
    perturb(model, seed + 1)
    g = torch.Generator().manual_seed(7 + seed)
    x = torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g))
    with torch.no_grad():
        y = model(x)
    save(name, cfg, model, {"spec_in": x, "spec_out": y})


if __name__ == "__main__":
    torch.set_num_threads(8)
    mac = ["swiglu_conv1d", "swiglu_conv1d"]
    # hop = n_fft/2 like BASELINE Variant D; macaron; 4 sources; head_dim 8
    mss_case("mss_hop2_macaron", dict(n_fft=256, hop_length=128, n_sources=4, n_layers=2, emb_dim=32,
             norm_type="rmsgroupnorm", num_groups=4, tf_order="ft", n_heads=4, flash_attention=False,
             attention_dim=32, pos_enc="rope", ffn_type=mac, ffn_hidden_dim=[64, 64], conv1d_kernel=4,
             conv1d_shift=1, dropout=0.0, eps=1e-5), n_samples=1700, batch=2, seed=0)
    # hop = n_fft/4 like configs/musdb18.yaml (Variant Y); single FFN; 'tf' order; head_dim 6; 2 sources
    mss_case("mss_hop4_single_tf", dict(n_fft=128, hop_length=32, n_sources=2, n_layers=1, emb_dim=24,
             norm_type="rmsgroupnorm", num_groups=2, tf_order="tf", n_heads=4, flash_attention=False,
             attention_dim=24, pos_enc="rope", ffn_type="swiglu_conv1d", ffn_hidden_dim=40, conv1d_kernel=4,
             conv1d_shift=1, dropout=0.0, eps=1e-5), n_samples=601, batch=1, seed=3)
    # the reference's own test shapes (tests/test_tflocoformer.py:48-72): [2, 50, 65], emb 32
    sep_case("sep_rope_k4", dict(num_spk=2, n_layers=2, emb_dim=32, norm_type="rmsgroupnorm", num_groups=4,
             tf_order="ft", n_heads=4, attention_dim=32, pos_enc="rope", ffn_type=mac,
             ffn_hidden_dim=[32, 32], conv1d_kernel=4, conv1d_shift=1, dropout=0.0, eps=1e-5),
             (2, 50, 65), seed=5)
    sep_case("sep_nope_k1", dict(num_spk=1, n_layers=1, emb_dim=32, norm_type="rmsgroupnorm", num_groups=1,
             tf_order="tf", n_heads=1, attention_dim=32, pos_enc="nope", ffn_type=mac,
             ffn_hidden_dim=[32, 32], conv1d_kernel=1, conv1d_shift=1, dropout=0.0, eps=1e-5),
             (2, 50, 65), seed=6)
    # ESPnet-recipe regime: kernel 8 (egs2/whamr/enh1/conf/tuning/train_enh_tflocoformer.yaml)
    sep_case("sep_rope_k8", dict(num_spk=2, n_layers=1, emb_dim=32, norm_type="rmsgroupnorm", num_groups=4,
             tf_order="ft", n_heads=4, attention_dim=32, pos_enc="rope", ffn_type=mac,
             ffn_hidden_dim=[48, 48], conv1d_kernel=8, conv1d_shift=1, dropout=0.0, eps=1e-5),
             (1, 20, 33), seed=7)
    # BS-Locoformer (tests/test_bslocoformer.py): stft 2048 -> 1025 bins, 62 bands; stereo + masking
    bs_case("bs_stereo_mask", dict(num_spk=2, n_layers=1, emb_dim=8, norm_type="rmsgroupnorm", num_groups=2,
            tf_order="ft", n_heads=2, attention_dim=8, pos_enc="rope", ffn_type=mac, ffn_hidden_dim=[16, 16],
            conv1d_kernel=4, conv1d_shift=1, dropout=0.0, eps=1e-5, sample_rate=44100, stft_size=2048,
            masking=True, stereo=True), (1, 2, 5, 1025), seed=8)
    bs_case("bs_mono_map", dict(num_spk=1, n_layers=1, emb_dim=8, norm_type="rmsgroupnorm", num_groups=2,
            tf_order="ft", n_heads=2, attention_dim=8, pos_enc="rope", ffn_type="swiglu_conv1d",
            ffn_hidden_dim=16, conv1d_kernel=4, conv1d_shift=1, dropout=0.0, eps=1e-5, sample_rate=48000,
            stft_size=2048, masking=False, stereo=False), (1, 4, 1025), seed=9)
