"""ncu target: one FFN call and one attention sub-block call (frequency axis, Variant D, batch 8, bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda()
eng = model._ready()
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")
y = torch.empty_like(x)
eng.ffn_out(0, 0, 0, x, y, 1)
eng.attn_(0, 0, x, 1)
torch.cuda.synchronize()
print("ok")
