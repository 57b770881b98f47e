"""A/B timing of the attention sub-block (qkv + attention + projection launches) for both attention kernels.

    python profiles/time_attn.py [batch]

Prints CUDA-event ms per tfl_rope_attn call per axis and kernel version, the max difference of the results, and the
bounded-wait diagnostic (a protocol bug shows up as a timeout record instead of a hang).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_state_dict  # noqa: E402
from mss_tf_locoformer_b200 import _lib  # noqa: E402
from mss_tf_locoformer_b200.engine import debug_timeout  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda()
eng = model._ready()
lib = _lib.load()
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x0 = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")
res = {}
for axis in (0, 1):
    for ver in (1, 2):
        lib.tfl_debug_set_option(0, ver)
        x = x0.clone()
        eng.attn_(0, axis, x, 1)
        torch.cuda.synchronize()
        res[(axis, ver)] = x.clone()
        t = debug_timeout(True)
        if t[0]:
            print("TIMEOUT", axis, ver, t)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 10
        ev[0].record()
        for _ in range(n):
            eng.attn_(0, axis, x, 1)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"axis {axis} kernel v{ver}: {ev[0].elapsed_time(ev[1]) / n:.3f} ms per sub-block call")
    d = (res[(axis, 1)] - res[(axis, 2)]).abs().max().item()
    ref = (res[(axis, 1)] - x0).abs().max().item()
    print(f"axis {axis}: max |v1 - v2| = {d:.3e} (update magnitude {ref:.3e})")
lib.tfl_debug_set_option(0, 2)
