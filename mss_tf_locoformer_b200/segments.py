"""Full-track separation by overlapping fixed-length segments, sharded across the ranks of one node.

The reference runs a whole track in one forward (inference/separate.py:135-148) and documents the OOMs that
follow (MEMORY_ANALYSIS.md:7-11); BASELINE config 3 instead cuts the track into 6-s segments at 50 % overlap,
each an independent forward, cross-faded with a periodic-Hann partition of unity (the first / last segment keep
a flat outer edge).  Segments are the sharding unit: rank r owns a contiguous run, overlap-adds its outputs
into its own zeroed track buffer (tfl_segment_ola) and ONE sum-reduction merges the ranks -- only the
half-segment halos between neighbouring ranks hold non-zero data from two ranks.
"""
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch

from .engine import segment_ola

SEGMENT_SAMPLES = 264600  # 6 s at 44.1 kHz


def segment_starts(n_samples: int, seg_len: int) -> List[int]:
    hop = seg_len // 2
    n = max(1, math.ceil((n_samples - seg_len) / hop) + 1)
    return [i * hop for i in range(n)]


def partition(n_seg: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous block partition, larger blocks first: 79 over 8 -> 10,10,10,10,10,10,10,9."""
    base, extra = divmod(n_seg, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def separate_track(model: Callable[[torch.Tensor], Dict[str, torch.Tensor]], track: torch.Tensor,
                   seg_len: int = SEGMENT_SAMPLES, batch: int = 8, group=None,
                   ola: Optional[Callable] = None) -> Dict[str, torch.Tensor]:
    """track [T] mono (every rank holds it) -> {source: [T]} on every rank.

    ``model([b, seg_len]) -> {name: [b, seg_len]}``; ``ola(seg_out[S, b, L], first_index, n_seg, acc[S, T'])``
    accumulates windowed segments (default: the CUDA kernel behind tfl_segment_ola).
    """
    import torch.distributed as dist
    if seg_len % 2:
        raise ValueError("segment length must be even (50 % overlap)")
    ola = ola or segment_ola
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_samples = track.shape[-1]
    starts = segment_starts(n_samples, seg_len)
    n_seg = len(starts)
    lo, hi = partition(n_seg, world)[rank]
    total = starts[-1] + seg_len
    padded = torch.zeros(total, dtype=track.dtype, device=track.device)
    padded[:n_samples] = track
    acc, names = None, None
    for i0 in range(lo, hi, batch):
        idx = range(i0, min(i0 + batch, hi))
        segs = torch.stack([padded[starts[i]:starts[i] + seg_len] for i in idx])
        out = model(segs)
        names = list(out.keys())
        seg_out = torch.stack([out[k] for k in names])            # [S, b, L]
        if acc is None:
            acc = torch.zeros(len(names), total, dtype=seg_out.dtype, device=seg_out.device)
        ola(seg_out, i0, n_seg, acc)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, names, group=group)      # a rank with no segment learns the source names
        if acc is None:
            names = next(n for n in gathered if n is not None)
            acc = torch.zeros(len(names), total, dtype=torch.float32, device=track.device)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return {k: acc[i, :n_samples] for i, k in enumerate(names)}
