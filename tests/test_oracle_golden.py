"""Pin the CPU oracle (oracle/) against vectors produced by the reference itself.

tests/golden/*.npz come from tests/golden/make_golden.py, which runs /root/reference
unmodified (RoPE via the flagged stand-in, oracle/rope.py).  CPU-only.
"""
import pytest
import torch

import oracle
from conftest import load_golden

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order


def maxdiff(a, b):
    return float((a - b).abs().max())


@pytest.mark.parametrize("name", ["mss_hop2_macaron", "mss_hop4_single_tf"])
def test_mss_forward_matches_reference(name):
    cfg, sd, arr = load_golden(name)
    out = oracle.mss_forward(sd, cfg, arr["mixture"])
    for k, v in out.items():
        ref = arr["out/" + k]
        assert v.shape == ref.shape
        assert maxdiff(v, ref) < TOL, (k, maxdiff(v, ref))
        assert oracle.si_sdr_db(v, ref) > 90.0
    if cfg["n_sources"] >= 4:
        sp = oracle.mss_forward(sd, cfg, arr["mixture"], return_time_domain=False)
        for k, v in sp.items():
            assert maxdiff(torch.view_as_real(v), torch.view_as_real(arr["spec/" + k])) < 5e-5


@pytest.mark.parametrize("name", ["mss_hop2_macaron", "mss_hop4_single_tf"])
def test_stft_istft_match_torch(name):
    cfg, _, arr = load_golden(name)
    n_fft, hop = cfg["n_fft"], cfg["hop_length"]
    x = arr["mixture"]
    spec = oracle.stft(x, n_fft, hop)                       # [B, Tf, F]
    assert maxdiff(torch.view_as_real(spec.transpose(1, 2).contiguous()),
                   torch.view_as_real(arr["stft"].contiguous())) < 2e-5
    win = torch.hann_window(n_fft)
    back = torch.istft(arr["stft"], n_fft, hop, n_fft, win, length=x.shape[-1])
    mine = oracle.istft(spec, n_fft, hop, x.shape[-1])
    assert maxdiff(mine, back) < 1e-5
    if x.shape[-1] % hop != hop - 1:   # NOLA round trip holds away from the truncated tail
        n_ok = (x.shape[-1] // hop) * hop - n_fft
        assert maxdiff(mine[:, :n_ok], x[:, :n_ok]) < 1e-5


def test_stage_vectors():
    cfg, sd, arr = load_golden("mss_hop2_macaron")
    n_fft, hop = cfg["n_fft"], cfg["hop_length"]
    spec = oracle.stft(arr["mixture"], n_fft, hop)
    x = torch.stack([spec.real, spec.imag], -1)
    enc = oracle.encoder(x, sd["conv.0.weight"], sd["conv.0.bias"], sd["conv.1.weight"], sd["conv.1.bias"], cfg["eps"])
    ref_enc = arr["stage/conv:out"].permute(0, 2, 3, 1)     # [B,C,Tf,F] -> [B,Tf,F,C]
    assert maxdiff(enc, ref_enc) < TOL
    blk = oracle.tf_block(ref_enc.contiguous(), sd, "blocks.0", cfg)
    assert maxdiff(blk, arr["stage/blocks.0:out"].permute(0, 2, 3, 1)) < 5e-5
    # RMSGroupNorm
    nin, nout = arr["stage/blocks.0.freq_path.attn_norm:in"], arr["stage/blocks.0.freq_path.attn_norm:out"]
    got = oracle.rms_group_norm(nin, sd["blocks.0.freq_path.attn_norm.gamma"], cfg["num_groups"], cfg["eps"])
    assert maxdiff(got, nout) < 1e-5
    # attention on the frequency path: [B*Tf, F, C]
    b, tf, f, c = nout.shape
    att = oracle.attention(nout.reshape(b * tf, f, c), sd["blocks.0.freq_path.attn.qkv.weight"],
                           sd["blocks.0.freq_path.attn.aggregate_heads.0.weight"], cfg["n_heads"],
                           sd["blocks.0.freq_path.attn.rope.freqs"])
    assert maxdiff(att, arr["stage/blocks.0.freq_path.attn:out"]) < 1e-5
    # ConvSwiGLU FFN on the time path: input [B, F, Tf, C]
    fin, fout = arr["stage/blocks.0.frame_path.ffn.0:in"], arr["stage/blocks.0.frame_path.ffn.0:out"]
    b, s1, s2, c = fin.shape
    p = "blocks.0.frame_path.ffn.0."
    got = oracle.swiglu_conv_deconv(fin.reshape(b * s1, s2, c), sd[p + "conv1d.weight"], sd[p + "conv1d.bias"],
                                    sd[p + "deconv1d.weight"], sd[p + "deconv1d.bias"]).reshape(fout.shape)
    assert maxdiff(got, fout) < 1e-5
    # decoder
    dec = oracle.decoder(arr["stage/blocks.0:out"].permute(0, 2, 3, 1) * 0 + oracle.blocks_forward(ref_enc.contiguous(), sd, cfg),
                         sd["deconv.weight"], sd["deconv.bias"])
    assert maxdiff(dec, arr["stage/deconv:out"].permute(0, 2, 3, 1)) < 1e-4


@pytest.mark.parametrize("name", ["sep_rope_k4", "sep_nope_k1", "sep_rope_k8"])
def test_separator_matches_reference(name):
    cfg, sd, arr = load_golden(name)
    out = oracle.separator_forward(sd, cfg, arr["spec_in"])
    assert out.shape == arr["spec_out"].shape
    assert maxdiff(torch.view_as_real(out), torch.view_as_real(arr["spec_out"])) < 5e-5


@pytest.mark.parametrize("name", ["bs_stereo_mask", "bs_mono_map"])
def test_bs_matches_reference(name):
    cfg, sd, arr = load_golden(name)
    out = oracle.bs_forward(sd, cfg, arr["spec_in"])
    assert out.shape == arr["spec_out"].shape
    assert maxdiff(torch.view_as_real(out), torch.view_as_real(arr["spec_out"])) < 1e-4


def test_bands():
    assert len(oracle.bs_bands(44100, 2048)) == 62
    assert len(oracle.bs_bands(48000, 2048)) == 61


def test_stitch_is_partition_of_unity():
    seg = 64
    for n in (64, 65, 96, 97, 200, 33):
        starts = oracle.segment_starts(n, seg)
        ident = oracle.separate_track(lambda x: {"a": x, "b": 2 * x}, torch.arange(1, n + 1, dtype=torch.float32), seg, batch=3)
        assert torch.allclose(ident["a"], torch.arange(1, n + 1, dtype=torch.float32), atol=1e-4), (n, starts)
        assert torch.allclose(ident["b"], 2 * torch.arange(1, n + 1, dtype=torch.float32), atol=2e-4)
    assert len(oracle.segment_starts(10_584_000, 264_600)) == 79


def test_rope_matches_independent_gptj_implementation():
    """oracle/rope.py against an implementation nobody here wrote: transformers' GPT-J rotary embedding
    (`create_sinusoidal_positions`, `rotate_every_two`, `apply_rotary_pos_emb`), which is the same published
    convention rotary-embedding-torch follows for `RotaryEmbedding(dim).rotate_queries_or_keys` (interleaved pairs,
    theta = 10000, positions from 0, whole head rotated).  rotary-embedding-torch==0.6.1 itself is absent from the image
    (requirements.txt:23; `pip download` fails offline), so this is the strongest pin available for that boundary."""
    gptj = pytest.importorskip("transformers.models.gptj.modeling_gptj")
    from oracle import rope
    for hd, L in ((32, 1025), (24, 517), (12, 259), (8, 50)):
        g = torch.Generator().manual_seed(hd)
        t = torch.randn(2, 4, L, hd, generator=g)                       # [B, H, L, hd]; the reference's layout
        got = rope.rope_rotate(t, rope.rope_freqs(hd))
        sincos = gptj.create_sinusoidal_positions(L, hd)               # [L, hd]: sin | cos
        sin, cos = sincos[None, :, : hd // 2], sincos[None, :, hd // 2:]
        want = gptj.apply_rotary_pos_emb(t.transpose(1, 2), sin, cos).transpose(1, 2)   # GPT-J wants [B, L, H, hd]
        assert torch.allclose(got, want, atol=2e-5, rtol=0), float((got - want).abs().max())
