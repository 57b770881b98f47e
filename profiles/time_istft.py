"""CUDA-event timing of the iSTFT + overlap-add stage at the bench shape (batch 8, 4 sources, 259 frames of 2048).

    python profiles/time_istft.py [batch]
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
est = torch.randn(B, cfg["n_sources"], Tf, F, 2, device="cuda")
est[..., 0, 1] = 0.0    # DC and Nyquist of a real signal's spectrum are real (irfft ignores their imaginary parts on the
est[..., -1, 1] = 0.0   # CPU; cuFFT's C2R does not define what it does with them)
want = torch.istft(torch.view_as_complex(est).reshape(-1, Tf, F).transpose(1, 2), cfg["n_fft"], cfg["hop_length"],
                   window=torch.hann_window(cfg["n_fft"], device="cuda"), length=SEG)
got = eng.istft(est, SEG)                               # [S, B, T]
want = want.reshape(B, cfg["n_sources"], SEG).transpose(0, 1)
print("max |istft_ola - torch.istft| =", float((got - want).abs().max()), " scale", float(want.abs().max()))
for _ in range(3):
    eng.istft(est, SEG)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(20):
    eng.istft(est, SEG)
ev[1].record()
torch.cuda.synchronize()
print(f"istft_ola: {ev[0].elapsed_time(ev[1]) / 20:.3f} ms per call")
