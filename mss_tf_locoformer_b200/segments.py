"""Full-track separation by overlapping fixed-length segments, sharded across the ranks of one node.

The reference runs a whole track in one forward (inference/separate.py:135-148) and documents the OOMs that
follow (MEMORY_ANALYSIS.md:7-11); BASELINE config 3 instead cuts the track into 6-s segments at 50 % overlap,
each an independent forward, cross-faded with a periodic-Hann partition of unity (the first / last segment keep
a flat outer edge; the behaviour is defined once, in oracle/stitch.py).

Segments are the sharding unit.  Rank r owns a contiguous run of segments [lo, hi) and therefore the samples
[lo * half, hi * half) of the track (half = seg_len / 2; the last non-empty rank owns up to the end).  It
overlap-adds its segments (tfl_segment_ola) into a LOCAL buffer covering its own range plus a right halo of
`half` samples -- the second half of its last segment, which belongs to the next rank.  The exchange step is then
  1. halo:   rank r sends its right halo [S, half] (2.1 MB for 4 sources at 6-s segments) to the next non-empty
             rank, which adds it onto the first `half` samples of its own range (point-to-point, NCCL send/recv);
  2. gather: all-gather of the disjoint own ranges (padded to the longest) -- every rank ends with the full track.
No whole-track reduction, no pickled Python objects, no host synchronisation.
"""
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

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


def owned_ranges(n_seg: int, seg_len: int, world: int) -> List[Tuple[int, int]]:
    """Sample range [a, b) of the padded track each rank owns (empty for a rank without segments)."""
    half = seg_len // 2
    total = (n_seg - 1) * half + seg_len
    parts = partition(n_seg, world)
    last = max(r for r, (lo, hi) in enumerate(parts) if hi > lo)
    out = []
    for r, (lo, hi) in enumerate(parts):
        if hi <= lo:
            out.append((0, 0))
        else:
            out.append((lo * half, total if r == last else hi * half))
    return out


def separate_track(model: Callable[[torch.Tensor], Dict[str, torch.Tensor]], track: torch.Tensor,
                   seg_len: int = SEGMENT_SAMPLES, batch: int = 8, group=None, ola: Optional[Callable] = None,
                   names: Optional[Sequence[str]] = None) -> Dict[str, torch.Tensor]:
    """track [T] mono (every rank holds it) -> {source: [T]} on every rank.

    ``model([b, seg_len]) -> {name: [b, seg_len]}``; ``ola(seg_out[S, b, L], first_index, n_seg, acc[S, n], origin)``
    accumulates windowed segments into a buffer that starts at sample ``origin`` of the track (default: the CUDA
    kernel behind tfl_segment_ola).  ``names``: output keys, needed only by a rank that owns no segment (more ranks
    than segments); such a rank otherwise learns them from one forward of a silent segment.
    """
    import torch.distributed as dist
    if seg_len % 2:
        raise ValueError("segment length must be even (50 % overlap)")
    ola = ola or segment_ola
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_samples = track.shape[-1]
    half = seg_len // 2
    starts = segment_starts(n_samples, seg_len)
    n_seg = len(starts)
    total = starts[-1] + seg_len
    lo, hi = partition(n_seg, world)[rank]
    own = owned_ranges(n_seg, seg_len, world)
    a, b = own[rank]
    padded = torch.zeros(total, dtype=track.dtype, device=track.device)
    padded[:n_samples] = track
    non_empty = [r for r in range(world) if own[r][1] > own[r][0]]
    is_last = rank == non_empty[-1]
    local_len = (b - a) + (0 if (is_last or hi <= lo) else half)          # own range + right halo
    acc = None
    for i0 in range(lo, hi, batch):
        idx = range(i0, min(i0 + batch, hi))
        segs = torch.stack([padded[starts[i]:starts[i] + seg_len] for i in idx])
        out = model(segs)
        names = list(out.keys())
        seg_out = torch.stack([out[k] for k in names])            # [S, b, L]
        if acc is None:
            acc = torch.zeros(len(names), local_len, dtype=seg_out.dtype, device=seg_out.device)
        ola(seg_out, i0, n_seg, acc, a)
    if world == 1:
        return {k: acc[i, :n_samples] for i, k in enumerate(names)}
    if names is None:                                             # a rank without segments (world > n_seg)
        names = list(model(padded.new_zeros(1, seg_len)).keys())
    S = len(names)
    if acc is None:
        acc = torch.zeros(S, 0, dtype=torch.float32, device=track.device)
    # ---- 1. halo exchange between neighbouring non-empty ranks (point to point) ----
    if rank in non_empty:
        k = non_empty.index(rank)
        ops, recv = [], None
        if k + 1 < len(non_empty):
            send = acc[:, b - a:].contiguous()                    # [S, half]: belongs to the next rank
            ops.append(dist.P2POp(dist.isend, send, non_empty[k + 1], group))
        if k > 0:
            recv = torch.empty(S, half, dtype=acc.dtype, device=acc.device)
            ops.append(dist.P2POp(dist.irecv, recv, non_empty[k - 1], group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if recv is not None:
            acc[:, :half] += recv
    # ---- 2. all-gather of the disjoint own ranges (padded to the longest) ----
    longest = max(hi_ - lo_ for lo_, hi_ in own)
    mine = torch.zeros(S, longest, dtype=acc.dtype, device=acc.device)
    mine[:, : b - a] = acc[:, : b - a]
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = torch.empty(S, total, dtype=acc.dtype, device=acc.device)
    for r, (ra, rb) in enumerate(own):
        if rb > ra:
            full[:, ra:rb] = parts[r][:, : rb - ra]
    return {k: full[i, :n_samples] for i, k in enumerate(names)}
