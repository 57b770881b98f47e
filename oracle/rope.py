"""RoPE restatement -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED at this boundary: the reference takes its rotary arithmetic from the
third-party package ``rotary-embedding-torch==0.6.1`` (pinned in
/root/reference/requirements.txt:23), which is neither vendored under /root/reference nor
installable in the build container (no network).  No reference test holds a numeric
fixture for it (reference tests only check shapes: tests/test_tflocoformer.py:77).

What is restated here is the package's published algorithm for the exact call the
reference makes -- ``RotaryEmbedding(dim=head_dim).rotate_queries_or_keys(t)`` with
all-default constructor arguments (call sites: models/mss_tflocoformer.py:153-154 and
:557-558; standalone/tflocoformer_separator.py:99-100,437-438):

  * ``freqs_i = theta ** -(2 i / dim)``, ``i = 0 .. dim/2 - 1``, ``theta = 10000``;
    stored as the non-trainable parameter ``freqs`` (hence the ``...attn.rope.freqs``
    state_dict key).
  * positions ``0 .. L-1`` along the sequence axis (dim -2).
  * angle ``pos * freqs_i`` is shared by the INTERLEAVED pair ``(x[2i], x[2i+1])``:
    ``out[2i]   = x[2i] cos - x[2i+1] sin``
    ``out[2i+1] = x[2i+1] cos + x[2i] sin``
  * the rotation covers all ``dim`` channels of the head; result cast back to t.dtype.
"""
import torch
from torch import nn


def rope_freqs(head_dim: int, theta: float = 10000.0) -> torch.Tensor:
    """``freqs`` buffer of RotaryEmbedding(dim=head_dim) (lang frequencies)."""
    idx = torch.arange(0, head_dim, 2)[: head_dim // 2].float()
    return 1.0 / (theta ** (idx / head_dim))


def rope_rotate(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Rotate ``t[..., L, head_dim]`` by position along dim -2 (interleaved pairs)."""
    length = t.shape[-2]
    pos = torch.arange(length, device=t.device, dtype=torch.float32)
    ang = pos[:, None] * freqs.to(torch.float32)[None, :]          # [L, hd/2]
    cos = ang.cos().to(t.dtype if t.dtype == torch.float64 else torch.float32)
    sin = ang.sin().to(cos.dtype)
    x = t.to(cos.dtype)
    even, odd = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = even * cos - odd * sin
    out[..., 1::2] = odd * cos + even * sin
    return out.to(t.dtype)


class RotaryEmbedding(nn.Module):
    """Stand-in with the third-party module's surface the reference touches.

    Used ONLY by tests/golden/make_golden.py to let the reference import in the build
    container; it owns a ``freqs`` parameter so reference state_dicts keep their keys.
    """

    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.freqs = nn.Parameter(rope_freqs(dim, theta), requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        assert seq_dim == -2
        return rope_rotate(t, self.freqs)
