"""GPU tests of the tcgen05 / TMEM path (TFL_PRECISION_BF16)."""
import math

import pytest
import torch

import oracle
from conftest import load_golden
from test_gpu_parity import VARIANT_D, VARIANT_Y, _check, _mixture, _random_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import mss_tf_locoformer_b200 as m
    assert torch.cuda.is_available()
    return m


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("N,Kd,taps", [(64, 128, 1), (64, 128, 4), (128, 32, 4), (96, 96, 4), (256, 64, 2), (16, 16, 8)])
def test_tc_selftest_kmajor_taps(pkg, N, Kd, taps):
    """Row-shifted A descriptors (implicit-GEMM conv taps), K-major B staged by a bulk copy."""
    from mss_tf_locoformer_b200.engine import tc_selftest
    g = torch.Generator().manual_seed(N * 1000 + Kd + taps)
    A = torch.randn(128 + taps - 1, Kd, generator=g)
    B = torch.randn(taps, N, Kd, generator=g)
    D = tc_selftest(A.cuda(), B.cuda(), taps, 0).cpu()
    Ab, Bb = _bf16(A).double(), _bf16(B).double()
    want = sum(Ab[t:t + 128] @ Bb[t].T for t in range(taps)).float()
    assert float((D - want).abs().max()) < 2e-3 * (Kd * taps) ** 0.5


@pytest.mark.parametrize("N,Kd", [(32, 128), (64, 64), (16, 256)])
def test_tc_selftest_mn_major_b(pkg, N, Kd):
    """MN-major B (V in P.V): same chunk-major tile, rows indexed by K."""
    from mss_tf_locoformer_b200.engine import tc_selftest
    g = torch.Generator().manual_seed(N + Kd)
    A = torch.randn(128, Kd, generator=g)
    V = torch.randn(Kd, N, generator=g)
    D = tc_selftest(A.cuda(), V.cuda(), 1, 1).cpu()
    want = (_bf16(A).double() @ _bf16(V).double()).float()
    assert float((D - want).abs().max()) < 2e-3 * Kd ** 0.5


@pytest.mark.parametrize("N,Kd", [(32, 64), (32, 32), (16, 16), (64, 128)])
def test_tc_selftest_a_in_tmem(pkg, N, Kd):
    """P.V form of attn_tc2_kernel: A staged in TMEM by tcgen05.st as packed bf16 pairs, V MN-major in smem."""
    from mss_tf_locoformer_b200.engine import tc_selftest
    g = torch.Generator().manual_seed(7 * N + Kd)
    A = torch.randn(128, Kd, generator=g)
    V = torch.randn(Kd, N, generator=g)
    D = tc_selftest(A.cuda(), V.cuda(), 1, 2).cpu()
    want = (_bf16(A).double() @ _bf16(V).double()).float()
    assert float((D - want).abs().max()) < 2e-3 * Kd ** 0.5


def _ffn_oracle(sd, cfg, xin, layer, axis, j):
    path = "freq_path" if axis == 0 else "frame_path"
    p = f"blocks.{layer}.{path}"
    xa = xin if axis == 0 else xin.transpose(1, 2).contiguous()
    b, s1, s2, c = xa.shape
    xn = oracle.rms_group_norm(xa, sd[f"{p}.ffn_norm.{j}.gamma"], cfg["num_groups"], cfg["eps"])
    y = oracle.swiglu_conv_deconv(xn.reshape(b * s1, s2, c), sd[f"{p}.ffn.{j}.conv1d.weight"],
                                  sd[f"{p}.ffn.{j}.conv1d.bias"], sd[f"{p}.ffn.{j}.deconv1d.weight"],
                                  sd[f"{p}.ffn.{j}.deconv1d.bias"]).reshape(xa.shape)
    return (y if axis == 0 else y.transpose(1, 2)), xin


@pytest.mark.parametrize("name,cfg,shape", [("golden32", None, None), ("D", VARIANT_D, (1, 5, 300)),
                                            ("Y", VARIANT_Y, (2, 7, 131)), ("D_long", VARIANT_D, (1, 3, 1025))])
def test_ffn_tc_vs_oracle(pkg, name, cfg, shape):
    """Fused norm + ConvSwiGLU + residual tcgen05 kernel on both axes; error judged on the FFN branch itself."""
    if cfg is None:
        cfg, sd, arr = load_golden("mss_hop2_macaron")
        model = pkg.TFLocoformerMSS(**cfg)
        model.load_state_dict(sd)
        xin = arr["stage/conv:out"].permute(0, 2, 3, 1).contiguous()
    else:
        model = _random_model(pkg, cfg)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        g = torch.Generator().manual_seed(3)
        xin = torch.randn(*shape, cfg["emb_dim"], generator=g)
    model = model.cuda().eval()
    eng = model._ready()
    for axis in (0, 1):
        for j in (0, 1):
            branch, x0 = _ffn_oracle(sd, cfg, xin, 0, axis, j)
            got = eng.ffn_(0, axis, j, xin.cuda().clone(), 1).cpu()
            got_branch = got - x0
            sdr = oracle.si_sdr_db(got_branch, branch)
            assert sdr > 40.0, (name, axis, j, sdr)
            assert float((got_branch - branch).abs().max()) < 0.05 * float(branch.abs().max()), (name, axis, j)


@pytest.mark.parametrize("name", ["mss_hop2_macaron"])
def test_mss_forward_golden_bf16(pkg, name):
    cfg, sd, arr = load_golden(name)
    model = pkg.TFLocoformerMSS(**cfg)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    model.precision = "bf16"
    with torch.no_grad():
        out = model(arr["mixture"].cuda())
    for k, v in out.items():
        assert oracle.si_sdr_db(v.cpu(), arr["out/" + k]) >= 40.0, k
    with torch.autocast("cuda", dtype=torch.bfloat16):          # the reference's way of entering bf16 mode
        model.precision = None
        with torch.no_grad():
            out2 = model(arr["mixture"].cuda())
    for k in out:
        assert torch.equal(out[k], out2[k])


@pytest.mark.parametrize("name,cfg,n_samples", [("D", VARIANT_D, 30000), ("Y", VARIANT_Y, 20000)])
def test_real_width_one_layer_bf16(pkg, name, cfg, n_samples):
    model = _random_model(pkg, cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mix = _mixture(n_samples, 1)
    want = oracle.mss_forward(sd, dict(cfg), mix)
    model = model.cuda()
    model.precision = "bf16"
    with torch.no_grad():
        got = model(mix.cuda())
    for k in want:
        assert oracle.si_sdr_db(got[k].cpu(), want[k]) >= 40.0, (name, k)


@pytest.mark.parametrize("name,cfg,shape", [("golden32", None, None), ("D", VARIANT_D, (1, 5, 300)),
                                            ("Y", VARIANT_Y, (2, 7, 131)), ("D_long", VARIANT_D, (1, 3, 1025)),
                                            ("D_time", VARIANT_D, (1, 259, 6)),
                                            ("D_time_spill3", VARIANT_D, (1, 259, 100)),     # 100 sequences x 3 tail rows
                                            ("D_freq_spill2", VARIANT_D, (1, 150, 1025))])   # 150 sequences x 1 tail row
def test_attention_tc_vs_oracle(pkg, name, cfg, shape):
    """tcgen05 attention (S = QK^T and P.V on the tensor core, exp2 softmax) on both axes."""
    if cfg is None:
        cfg, sd, arr = load_golden("mss_hop2_macaron")
        model = pkg.TFLocoformerMSS(**cfg)
        model.load_state_dict(sd)
        xin = arr["stage/conv:out"].permute(0, 2, 3, 1).contiguous()
    else:
        model = _random_model(pkg, cfg)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        g = torch.Generator().manual_seed(3)
        xin = torch.randn(*shape, cfg["emb_dim"], generator=g)
    model = model.cuda().eval()
    eng = model._ready()
    for axis, path in ((0, "freq_path"), (1, "frame_path")):
        p = f"blocks.0.{path}"
        xa = xin if axis == 0 else xin.transpose(1, 2).contiguous()
        b, s1, s2, c = xa.shape
        xn = oracle.rms_group_norm(xa, sd[f"{p}.attn_norm.gamma"], cfg["num_groups"], cfg["eps"])
        y = oracle.attention(xn.reshape(b * s1, s2, c), sd[f"{p}.attn.qkv.weight"],
                             sd[f"{p}.attn.aggregate_heads.0.weight"], cfg["n_heads"],
                             sd.get(f"{p}.attn.rope.freqs")).reshape(xa.shape)
        branch = y if axis == 0 else y.transpose(1, 2)
        got = eng.attn_(0, axis, xin.cuda().clone(), 1).cpu() - xin
        sdr = oracle.si_sdr_db(got, branch)
        assert sdr > 35.0, (name, axis, sdr)
        assert float((got - branch).abs().max()) < 0.06 * float(branch.abs().max()) + 1e-3, (name, axis)


def test_variant_y_all_layers_bf16(pkg):
    """configs/musdb18.yaml as committed (hop 512, 4 layers, emb 96, head_dim 24 -> padded to 32): 1.5-s segment, all layers."""
    cfg = dict(VARIANT_Y, n_layers=4)
    model = _random_model(pkg, cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mix = _mixture(66150, 1)
    want = oracle.mss_forward(sd, dict(cfg), mix)
    model = model.cuda()
    model.precision = "bf16"
    with torch.no_grad():
        got = model(mix.cuda())
    for k in want:
        assert oracle.si_sdr_db(got[k].cpu(), want[k]) >= 40.0, k


def test_ffn_tc_wide_single_tile_variant(pkg):
    """emb_dim 256 / hidden 1024 (musdb18_rtx5090_xlarge.yaml widths): the NT = 1 instantiation of the FFN kernel."""
    cfg = dict(VARIANT_D, emb_dim=256, attention_dim=256, n_heads=16, num_groups=8, ffn_hidden_dim=[1024, 1024])
    model = _random_model(pkg, cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    xin = torch.randn(1, 4, 200, 256, generator=g)
    model = model.cuda().eval()
    eng = model._ready()
    for axis in (0, 1):
        branch, x0 = _ffn_oracle(sd, cfg, xin, 0, axis, 0)
        got = eng.ffn_(0, axis, 0, xin.cuda().clone(), 1).cpu() - x0
        assert oracle.si_sdr_db(got, branch) > 40.0, axis


def test_ffn_tc_kernel_variants_agree(pkg):
    """TFL_OPT_FFN_KERNEL: the 1-CTA (two tiles per CTA) and the 2-CTA (cta_group::2) kernels compute the same FFN."""
    from mss_tf_locoformer_b200 import _lib
    lib = _lib.load()
    model = _random_model(pkg, VARIANT_D).cuda().eval()
    eng = model._ready()
    g = torch.Generator().manual_seed(11)
    xin = torch.randn(2, 9, 300, VARIANT_D["emb_dim"], generator=g).cuda()
    out = {}
    try:
        for opt in (1, 2):
            assert lib.tfl_debug_set_option(1, opt) == 0
            out[opt] = [eng.ffn_(0, axis, 1, xin.clone(), 1) for axis in (0, 1)]
    finally:
        lib.tfl_debug_set_option(1, 2)
    for a, b in zip(out[1], out[2]):
        branch = (a - xin).float()
        assert float((a - b).abs().max()) <= 2e-2 * float(branch.abs().max())   # the 2-CTA kernel's order of accumulation differs
        assert oracle.si_sdr_db((b - xin).cpu(), branch.cpu()) > 50.0


def test_attention_tc_reference_raise(pkg):
    """The lazy softmax reference of attn_tc2_kernel: with q / k weights scaled by 6 the scores spread over ~18 log2
    units, so that in most rows a later key exceeds the reference taken from the first 16 keys by more than
    2^ATT2_TH -- the rare path that raises the reference, rescales O and the row sums and (when the second 16-key
    quarter of a half triggers it) redoes the first quarter's probabilities.  Checked against the oracle (sharp
    softmax amplifies the bf16 rounding of q and k: 20 dB gate) and against attn_tc_kernel, which keeps a true running
    maximum and reads the same bf16 images (40 dB)."""
    from mss_tf_locoformer_b200 import _lib
    from oracle.locoformer_oracle import rope_rotate
    lib = _lib.load()
    cfg = dict(VARIANT_D)
    model = _random_model(pkg, cfg)
    with torch.no_grad():
        for path in ("freq_path", "frame_path"):
            getattr(model.blocks[0], path).attn.qkv.weight[:2 * cfg["attention_dim"]] *= 6.0
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    xin = torch.randn(1, 130, 300, cfg["emb_dim"], generator=g)
    model = model.cuda().eval()
    eng = model._ready()
    hd = cfg["attention_dim"] // cfg["n_heads"]
    for axis, path in ((0, "freq_path"), (1, "frame_path")):
        p = f"blocks.0.{path}"
        xa = xin if axis == 0 else xin.transpose(1, 2).contiguous()
        b, s1, s2, c = xa.shape
        xn = oracle.rms_group_norm(xa, sd[f"{p}.attn_norm.gamma"], cfg["num_groups"], cfg["eps"]).reshape(b * s1, s2, c)
        # the case must exercise the path: rows whose later keys exceed the first quarter's maximum by > 16 log2 units
        qkv = (xn[:8] @ sd[f"{p}.attn.qkv.weight"].t()).reshape(8, s2, 3, cfg["n_heads"], hd).permute(2, 0, 3, 1, 4)
        q, k = rope_rotate(qkv[0], sd[f"{p}.attn.rope.freqs"]), rope_rotate(qkv[1], sd[f"{p}.attn.rope.freqs"])
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd) * math.log2(math.e)
        jump = s[..., 16:].max(-1).values - s[..., :16].max(-1).values
        assert float((jump > 16.0).float().mean()) > 0.3
        y = oracle.attention(xn, sd[f"{p}.attn.qkv.weight"], sd[f"{p}.attn.aggregate_heads.0.weight"], cfg["n_heads"],
                             sd.get(f"{p}.attn.rope.freqs")).reshape(xa.shape)
        branch = y if axis == 0 else y.transpose(1, 2)
        got = {}
        try:
            for opt in (2, 1):
                assert lib.tfl_debug_set_option(0, opt) == 0
                got[opt] = eng.attn_(0, axis, xin.cuda().clone(), 1).cpu() - xin
        finally:
            lib.tfl_debug_set_option(0, 2)
        assert oracle.si_sdr_db(got[2], branch) > 20.0, (axis, oracle.si_sdr_db(got[2], branch))
        assert oracle.si_sdr_db(got[2], got[1]) > 40.0, (axis, oracle.si_sdr_db(got[2], got[1]))
