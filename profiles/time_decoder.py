"""A/B timing of the bf16-mode decoder kernels (TFL_OPT_DEC_KERNEL: 1 = 9-tap gather, 2 = scatter form) at the bench
shape, CUDA events, 20 calls each after a warm-up; x (1.09 GB at batch 8) exceeds the L2.

    python profiles/time_decoder.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import VARIANT_D, SEG, make_state_dict  # noqa: E402
from mss_tf_locoformer_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(VARIANT_D)
model = make_state_dict(cfg).cuda()
eng = model._ready()
lib = _lib.load()
Tf, F = 1 + SEG // cfg["hop_length"], cfg["n_fft"] // 2 + 1
x = torch.randn(B, Tf, F, cfg["emb_dim"], device="cuda")


def timeit(fn, n=20):
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


out = {}
for opt, name in ((1, "dec_conv_mma_kernel (gather)"), (2, "dec_conv_scatter_kernel")):
    lib.tfl_debug_set_option(6, opt)
    out[opt] = eng.dec_conv(x, 1)
    ms = timeit(lambda: eng.dec_conv(x, 1))
    gb = x.numel() * 4 / 1e9
    print(f"{name}: {ms:.3f} ms per call, {gb / ms * 1e3:.0f} GB/s of x ({gb:.2f} GB read once algorithmically)")
lib.tfl_debug_set_option(6, 2)
print("max |gather - scatter| =", float((out[1] - out[2]).abs().max()), " output scale", float(out[1].abs().max()))
