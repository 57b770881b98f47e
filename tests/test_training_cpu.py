"""Host logic of the training step without a GPU: the flat gradient layout against the model's own parameters, and the
data-parallel gradient average on gloo (world_size 2)."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mss_tf_locoformer_b200 as pkg
from mss_tf_locoformer_b200 import _lib, training
from mss_tf_locoformer_b200.engine import Engine, weight_keys

CFG = dict(n_fft=256, hop_length=128, n_sources=4, n_layers=2, emb_dim=32, norm_type="rmsgroupnorm", num_groups=4,
           tf_order="ft", n_heads=4, flash_attention=False, attention_dim=32, pos_enc="rope",
           ffn_type=["swiglu_conv1d", "swiglu_conv1d"], ffn_hidden_dim=[48, 64], conv1d_kernel=4, conv1d_shift=1,
           dropout=0.0, eps=1e-5)


def test_grad_layout_matches_state_dict():
    """One slice per state_dict tensor in tfl_pack_weights order, sized like the tensor, 256-byte aligned, disjoint;
    attn.rope.freqs (requires_grad=False in the reference) has no slice."""
    for cfg in (CFG, dict(CFG, pos_enc="nope", ffn_type="swiglu_conv1d", ffn_hidden_dim=64)):
        model = pkg.TFLocoformerMSS(**cfg)
        eng = Engine(model._engine_cfg)
        sd = model.state_dict()
        keys = weight_keys(eng.cfg)
        n = len(keys)
        offs, sizes = (C.c_int64 * n)(), (C.c_int64 * n)()
        total = _lib.load().tfl_train_grad_layout(eng.plan, offs, sizes, n)
        assert total > 0
        end = 0
        for k, off, size in zip(keys, offs, sizes):
            assert size == sd[k].numel(), k
            if k.endswith("rope.freqs"):
                assert off == -1
                continue
            assert off >= end and off % 64 == 0, k
            end = off + size
        assert end <= total
        trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
        assert sum(s for o, s in zip(offs, sizes) if o >= 0) == trainable
        assert _lib.load().tfl_train_workspace_bytes(eng.plan, 2, 6000, 2048, 1024) > \
            _lib.load().tfl_train_stage_workspace_bytes(eng.plan, 2, 47, 129) > 0


def test_trainer_refuses_dropout_and_cpu():
    import pytest
    model = pkg.TFLocoformerMSS(**dict(CFG, dropout=0.1))
    with pytest.raises(NotImplementedError):
        training.Trainer(model)
    with pytest.raises(RuntimeError, match="no CPU path"):
        training.Trainer(pkg.TFLocoformerMSS(**CFG))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        training.allreduce_mean_(flat)
        want = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ret[rank] = bool(torch.allclose(flat, want))
    finally:
        dist.destroy_process_group()


def test_gradient_average_two_ranks_gloo():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
    flat = torch.ones(8)
    assert training.allreduce_mean_(flat) is flat       # no process group: untouched
