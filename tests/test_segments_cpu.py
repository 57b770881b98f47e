"""Host logic of the full-track segment scheduler, including the N > 1 path on gloo (world_size 2, CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mss_tf_locoformer_b200 import segments


def test_starts_and_partition_match_survey():
    assert segments.segment_starts(10_584_000, 264_600) == oracle.segment_starts(10_584_000, 264_600)
    assert len(segments.segment_starts(10_584_000, 264_600)) == 79
    assert [hi - lo for lo, hi in segments.partition(79, 8)] == [10] * 7 + [9]
    assert [hi - lo for lo, hi in segments.partition(79, 2)] == [40, 39]
    assert [hi - lo for lo, hi in segments.partition(79, 4)] == [20, 20, 20, 19]
    assert segments.partition(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    for n in (1, 63, 64, 65, 1000):
        assert segments.segment_starts(n, 64) == oracle.segment_starts(n, 64)


def _cpu_ola(seg_out, i0, n_seg, acc, origin=0):
    """CPU stand-in for tfl_segment_ola in host-logic tests (the product default is the CUDA kernel)."""
    s, b, length = seg_out.shape
    for k in range(b):
        gi = i0 + k
        w = oracle.segment_window(length, gi == 0, gi == n_seg - 1, seg_out.dtype)
        start = gi * (length // 2) - origin
        acc[:, start:start + length] += seg_out[:, k] * w


def _fake_model(x):
    return {"a": x * 2.0 + 1.0, "b": -x}


def _worker(rank, world, port, n, seg, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        track = torch.sin(torch.arange(n, dtype=torch.float32) * 0.01)
        out = segments.separate_track(_fake_model, track, seg_len=seg, batch=3, ola=_cpu_ola)
        want = oracle.separate_track(_fake_model, track, seg, batch=2)
        ok = all(torch.allclose(out[k], want[k], atol=1e-5) for k in want) and list(out) == list(want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_owned_ranges_tile_the_track():
    for n_seg, world in ((79, 8), (79, 2), (3, 4), (1, 2), (10, 3)):
        own = segments.owned_ranges(n_seg, 64, world)
        total = (n_seg - 1) * 32 + 64
        spans = [(a, b) for a, b in own if b > a]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))


@pytest.mark.parametrize("n,seg,world", [(1000, 64, 2), (40, 64, 2), (997, 128, 2), (1000, 64, 3), (100, 64, 4)])
def test_multi_rank_gloo_matches_single_process_oracle(n, seg, world):
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, seg, ret), nprocs=world, join=True)
        assert dict(ret) == {r: True for r in range(world)}


def test_single_process_path():
    track = torch.randn(500)
    out = segments.separate_track(_fake_model, track, seg_len=64, batch=4, ola=_cpu_ola)
    want = oracle.separate_track(_fake_model, track, 64)
    for k in want:
        assert torch.allclose(out[k], want[k], atol=1e-5)
