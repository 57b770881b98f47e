"""One TFLocoformerMSS forward (Variant D, batch 8, bf16) + one stand-alone RMSGroupNorm call: the ncu target for the
HBM-bound kernels (stft, enc_conv + gLN, rms_group_norm, dec_conv, istft_ola).

    ncu --set full --clock-control none -k regex:'stft_kernel|enc_conv|gln_|rms_group_norm|dec_conv|istft_ola' \
        -o gpurun_out/r02_hbm python profiles/run_forward.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANTS, SEG, make_mixture, make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(VARIANTS["D"])
model = make_state_dict(cfg).cuda()
model.precision = "bf16"
mix = make_mixture(B, SEG).cuda()
with torch.no_grad():
    out = model(mix)
eng = model._ready()
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")
y = eng.rms_group_norm(0, 0, 2, x)
torch.cuda.synchronize()
print("ok", float(out["vocals"].abs().mean()), float(y.abs().mean()))
