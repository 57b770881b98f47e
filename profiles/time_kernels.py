"""Per-kernel CUDA-event timing of one stage call, kernels separated by running the stage under torch profiler-free
event brackets is not possible from outside the library, so this script times whole stage calls:

    python profiles/time_kernels.py [batch]

ffn (both axes) and the attention sub-block (both axes), 10 calls each, CUDA events.
"""
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


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


for axis in (0, 1):
    print(f"axis {axis}: ffn {timeit(lambda: eng.ffn_(0, axis, 0, x, 1)):.3f} ms   attention sub-block "
          f"{timeit(lambda: eng.attn_(0, axis, x, 1)):.3f} ms")

spec = torch.randn(B, Tf, F, 2, device="cuda")
print(f"decoder conv {timeit(lambda: eng.dec_conv(x)):.3f} ms   encoder conv + gLN {timeit(lambda: eng.enc_conv_gln(spec)):.3f} ms")
