"""CPU oracle for the TF-Locoformer separation forward path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker.  The product package
(``mss_tf_locoformer_b200``) never imports it and has no CPU fallback.

It is a restatement, in plain torch CPU tensor algebra (matmul / rfft / elementwise),
of the reference's algorithm; every function cites the reference file:line it follows.
It is pinned against outputs of the reference itself (``tests/golden/*.npz``, generated
by ``tests/golden/make_golden.py`` which imports /root/reference in the build
container).  The one boundary that stays "parity unpinned" is RoPE (``oracle/rope.py``):
the reference delegates it to an un-vendored third-party package.
"""
from .locoformer_oracle import (  # noqa: F401
    stft, istft, encoder, rms_group_norm, swiglu_conv_deconv, attention, locoformer_path,
    tf_block, decoder, blocks_forward, mss_forward, separator_forward, bs_forward,
    bs_bands, si_sdr_db,
)
from .rope import rope_freqs, rope_rotate  # noqa: F401
from .stitch import segment_starts, segment_window, stitch_segments, separate_track  # noqa: F401
