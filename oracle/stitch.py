"""Full-track chunked separation: oracle stitch -- TEST INFRASTRUCTURE ONLY.

The reference has NO chunked inference (inference/separate.py:135-148 runs the whole
track in one forward), so this defines the behaviour of BASELINE config 3
once, for the checker: fixed-length segments at 50 % overlap, each an independent
forward, cross-faded with a periodic-Hann partition of unity.

  starts   s_i = i * (seg_len // 2), i = 0 .. n-1, n = max(1, ceil((T - seg_len) / hop) + 1)
  window   w_i[m] = 0.5 - 0.5 cos(2 pi m / seg_len), except the first segment's first
           half and the last segment's second half are 1 (no partner to fade with)
  output   y[t] = sum_i w_i[t - s_i] * f(x[s_i : s_i + seg_len])[t - s_i], cut to T
           (the last segment's input is zero-extended to seg_len).
"""
import math
from typing import Callable, Dict, List

import torch


def segment_starts(n_samples: int, seg_len: int) -> List[int]:
    hop = seg_len // 2
    n = max(1, math.ceil((n_samples - seg_len) / hop) + 1)
    return [i * hop for i in range(n)]


def segment_window(seg_len: int, first: bool, last: bool, dtype=torch.float32) -> torch.Tensor:
    m = torch.arange(seg_len, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2.0 * math.pi * m / seg_len)
    half = seg_len // 2
    if first:
        w[:half] = 1.0
    if last:
        w[half:] = 1.0
    return w.to(dtype)


def stitch_segments(seg_out: torch.Tensor, n_samples: int) -> torch.Tensor:
    """seg_out [n_seg, n_src, seg_len] -> [n_src, n_samples]."""
    n_seg, n_src, seg_len = seg_out.shape
    starts = segment_starts(n_samples, seg_len)
    assert len(starts) == n_seg
    total = starts[-1] + seg_len
    y = torch.zeros(n_src, total, dtype=seg_out.dtype)
    for i, s in enumerate(starts):
        w = segment_window(seg_len, i == 0, i == n_seg - 1, seg_out.dtype)
        y[:, s:s + seg_len] += seg_out[i] * w
    return y[:, :n_samples]


def separate_track(forward: Callable[[torch.Tensor], Dict[str, torch.Tensor]], track: torch.Tensor,
                   seg_len: int, batch: int = 1) -> Dict[str, torch.Tensor]:
    """track [T] mono -> {name: [T]} using ``forward([b, seg_len]) -> {name: [b, seg_len]}``."""
    n_samples = track.shape[-1]
    starts = segment_starts(n_samples, seg_len)
    padded = torch.zeros(starts[-1] + seg_len, dtype=track.dtype)
    padded[:n_samples] = track
    segs = torch.stack([padded[s:s + seg_len] for s in starts])
    outs, names = [], None
    for i in range(0, len(starts), batch):
        res = forward(segs[i:i + batch])
        names = list(res.keys())
        outs.append(torch.stack([res[k] for k in names], dim=1))
    seg_out = torch.cat(outs, dim=0)
    y = stitch_segments(seg_out, n_samples)
    return {k: y[i] for i, k in enumerate(names)}
