"""CPU-side checks of the drop-in boundary: constructor/assert/state_dict parity with the
reference, C-ABI symbol export, and loud failure without a GPU.  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

import mss_tf_locoformer_b200 as pkg
from mss_tf_locoformer_b200 import _lib, build as build_mod
from mss_tf_locoformer_b200.engine import weight_keys
from conftest import load_golden, ROOT

MAC = ["swiglu_conv1d", "swiglu_conv1d"]


@pytest.fixture(scope="module", autouse=True)
def built_library():
    build_mod.build()


@pytest.mark.parametrize("name,cls", [("mss_hop2_macaron", "TFLocoformerMSS"), ("mss_hop4_single_tf", "TFLocoformerMSS"),
                                      ("sep_rope_k4", "TFLocoformerSeparator"), ("sep_nope_k1", "TFLocoformerSeparator"),
                                      ("sep_rope_k8", "TFLocoformerSeparator")])
def test_reference_state_dict_loads_strict(name, cls):
    cfg, sd, _ = load_golden(name)
    model = getattr(pkg, cls)(**cfg)
    own = model.state_dict()
    assert list(own.keys()) == list(sd.keys())          # same names, same order as the reference
    for k in sd:
        assert own[k].shape == sd[k].shape and own[k].dtype == sd[k].dtype, k
    model.load_state_dict(sd, strict=True)
    prefixed = {"separator." + k: v for k, v in sd.items()}
    model.load_state_dict(pkg.strip_prefix(prefixed), strict=True)


@pytest.mark.parametrize("name", ["bs_stereo_mask", "bs_mono_map"])
def test_bs_state_dict_loads_strict(name):
    cfg, sd, _ = load_golden(name)
    model = pkg.BSLocoformerSeparator(**cfg)
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    assert len(model.band_split_module.bands) == (62 if cfg["sample_rate"] == 44100 else 61)


def test_seeded_init_matches_reference_init():
    """Same RNG consumption order as the reference constructor => same random-init weights."""
    cfg, sd, _ = load_golden("sep_nope_k1")
    torch.manual_seed(6)                                  # seed used by tests/golden/make_golden.py
    model = pkg.TFLocoformerSeparator(**cfg)
    w = model.state_dict()
    for k in ("conv.0.weight", "blocks.0.freq_path.attn.qkv.weight", "blocks.0.frame_path.ffn.1.conv1d.weight",
              "deconv.weight"):
        assert torch.equal(w[k], sd[k]), k                # 2-D+ weights are not perturbed by make_golden


def test_constructor_errors_match_reference():
    kw = dict(n_layers=1, emb_dim=32, attention_dim=32, n_heads=4, ffn_type=MAC, ffn_hidden_dim=[32, 32],
              norm_type="rmsgroupnorm")
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, n_heads=3))                        # attention_dim % n_heads
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, tf_order="xx"))
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, norm_type="rmsgrouporm"))          # the reference default's typo
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, ffn_hidden_dim=32))                # macaron needs a 2-list
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, num_groups=5))
    with pytest.raises(AssertionError):
        pkg.TFLocoformerSeparator(**dict(kw, ffn_type=["swiglu_conv1d", "nope"]))
    with pytest.raises(ValueError):
        pkg.TFLocoformerSeparator(**dict(kw, pos_enc="abs"))
    with pytest.raises(NotImplementedError):
        pkg.TFLocoformerSeparator(**dict(kw, conv1d_shift=2))
    with pytest.raises(NotImplementedError):
        pkg.TFLocoformerSeparator(**dict(kw, norm_type="layernorm"))
    with pytest.raises(NotImplementedError):
        pkg.TFLocoformerSeparator(**dict(kw, ffn_type=["conv1d", "conv1d"]))


@pytest.mark.parametrize("num_spk", [1, 2])
@pytest.mark.parametrize("n_layers", [1, 4])
@pytest.mark.parametrize("num_groups", [1, 4])
@pytest.mark.parametrize("tf_order", ["tf", "ft"])
@pytest.mark.parametrize("n_heads", [1, 4])
@pytest.mark.parametrize("pos_enc", ["rope", "nope"])
@pytest.mark.parametrize("conv1d_kernel", [1, 4])
def test_reference_test_grid_constructs(num_spk, n_layers, num_groups, tf_order, n_heads, pos_enc, conv1d_kernel):
    """The 128-combination grid of /root/reference/tests/test_tflocoformer.py:11-26: every combination must
    construct, expose one weight pointer per packed tensor, and be accepted by tfl_plan_create."""
    model = pkg.TFLocoformerSeparator(num_spk=num_spk, n_layers=n_layers, emb_dim=32, norm_type="rmsgroupnorm",
                                      num_groups=num_groups, tf_order=tf_order, n_heads=n_heads, attention_dim=32,
                                      pos_enc=pos_enc, ffn_type=MAC, ffn_hidden_dim=[32, 32],
                                      conv1d_kernel=conv1d_kernel, conv1d_shift=1, dropout=0.1, eps=1e-5)
    keys = weight_keys(model._engine_cfg)
    sd = model.state_dict()
    assert all(k in sd for k in keys)
    rope_dupes = [k for k in sd if k.endswith("rope.freqs")]
    assert len(keys) == len(sd) and len(rope_dupes) == (2 * n_layers if pos_enc == "rope" else 0)
    eng = pkg.Engine(model._engine_cfg)                    # plan creation is host-only
    assert eng.lib.tfl_packed_bytes(eng.plan) > 0
    assert eng.lib.tfl_workspace_bytes(eng.plan, 2, 50, 65, 0) > 0


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tfl.h")).read()
    declared = set(re.findall(r"\b(tfl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().tfl_version() >= 100


def test_plan_rejects_bad_configs():
    base = dict(n_fft=256, hop=128, n_src=4, n_layers=1, emb_dim=32, num_groups=4, tf_order=0, n_heads=4,
                attention_dim=32, rope=1, macaron=1, ffn_hidden0=64, ffn_hidden1=64, conv_kernel=4, enc_in_ch=2, eps=1e-5)
    pkg.Engine(base)
    for bad in (dict(n_fft=300), dict(emb_dim=30), dict(n_heads=3), dict(conv_kernel=0), dict(n_src=5), dict(hop=0)):
        with pytest.raises(_lib.TflError):
            pkg.Engine(dict(base, **bad))


def test_no_cpu_fallback():
    cfg, sd, arr = load_golden("sep_rope_k4")
    model = pkg.TFLocoformerSeparator(**cfg).eval()
    model.load_state_dict(sd)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(arr["spec_in"])
    with pytest.raises(NotImplementedError, match="forward-only"):
        model(arr["spec_in"].clone().requires_grad_(True))
    drop = pkg.TFLocoformerSeparator(**dict(cfg, dropout=0.1)).train()
    with pytest.raises(NotImplementedError, match="dropout"):
        drop(arr["spec_in"])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "mss_tf_locoformer_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
