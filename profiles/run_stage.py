"""Run one stage kernel of Variant D a few times (target for ncu captures).

    [TFL_OPTS=key:value,...] python profiles/run_stage.py ffn|attn [batch] [axis]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_state_dict  # noqa: E402

stage = sys.argv[1] if len(sys.argv) > 1 else "ffn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
axis = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda()
eng = model._ready()
for kv in os.environ.get("TFL_OPTS", "").split(","):          # e.g. TFL_OPTS=6:2 -> tfl_debug_set_option(6, 2)
    if kv:
        k, v = kv.split(":")
        assert eng.lib.tfl_debug_set_option(int(k), int(v)) == 0
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")
for _ in range(4):
    if stage == "ffn":
        eng.ffn_(0, axis, 0, x, 1)
    else:
        eng.attn_(0, axis, x, 1)
torch.cuda.synchronize()
print("ok", stage, B, axis)
