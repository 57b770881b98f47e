"""A/B timing of the stage calls (FFN and attention sub-block, both axes, Variant D): 30 calls each, CUDA events.
Run once per library build (TFL_LIB selects the .so); alternate the builds on the same box."""
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


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


tag = os.path.basename(os.environ.get("TFL_LIB", "default"))
out = [f"{tag:16s}"]
for axis in (0, 1):
    out.append(f"ffn{axis} {timeit(lambda: eng.ffn_out(0, axis, 0, x, y, 1)):.3f}")
    out.append(f"attn{axis} {timeit(lambda: eng.attn_(0, axis, x, 1)):.3f}")
print("  ".join(out))
