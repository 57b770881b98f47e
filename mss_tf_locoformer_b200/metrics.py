"""Evaluation metrics on the device -- evaluation/metrics.py:14-168 without the D2H copy of full tracks.

One kernel pass (tfl_pair_stats) reduces an (estimate, target) pair to five sums in double precision; every metric of
the reference file is a closed form of them.  The formulas below restate the reference's numpy code line for line:
`compute_si_sdr` :14-56 (zero-mean, eps = 1e-8 added to every energy), `compute_sdr` :59-86, `compute_sar` :89-125 and
`compute_sir` :128-168 (which the reference computes with the same projection as SAR).
"""
import math
from typing import Dict

import torch

from . import _lib
from ._lib import check
from .engine import _require_cuda, _stream

STATS_BLOCKS = 64


def pair_stats(estimate: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """estimate, target [..., n] CUDA fp32 -> [rows, 5] float64 (sum e, sum t, sum e^2, sum t^2, sum e*t) on the device."""
    _require_cuda(estimate, "estimate")
    _require_cuda(target, "target")
    if estimate.shape != target.shape:
        raise ValueError(f"shape mismatch: {tuple(estimate.shape)} vs {tuple(target.shape)}")
    e = estimate.detach().to(torch.float32).reshape(-1, estimate.shape[-1]).contiguous()
    t = target.detach().to(torch.float32).reshape(-1, target.shape[-1]).contiguous()
    rows, n = e.shape
    out = torch.empty((rows, 5), dtype=torch.float64, device=e.device)
    scratch = torch.empty((rows, STATS_BLOCKS, 5), dtype=torch.float64, device=e.device)
    lib = _lib.load()
    with torch.cuda.device(e.device):
        check(lib.tfl_pair_stats(e.data_ptr(), t.data_ptr(), rows, n, out.data_ptr(), scratch.data_ptr(),
                                 scratch.numel() * 8, _stream()))
    return out


def _flat_stats(estimate, target):
    """The reference flattens both signals before scoring: one row."""
    s = pair_stats(estimate.reshape(1, -1), target.reshape(1, -1))[0].tolist()   # the only D2H: five doubles
    return s, estimate.numel()


def compute_si_sdr(estimate: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> float:
    (se, st, see, stt, set_), n = _flat_stats(estimate, target)
    dot = set_ - se * st / n                       # <e - mean(e), t - mean(t)>
    t_energy0 = stt - st * st / n                  # |t - mean(t)|^2
    e_energy0 = see - se * se / n
    scale = dot / (t_energy0 + eps)
    signal = scale * scale * t_energy0 + eps
    noise = e_energy0 - 2.0 * scale * dot + scale * scale * t_energy0 + eps
    return float(10.0 * math.log10(signal / noise))


def compute_sdr(estimate: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> float:
    (se, st, see, stt, set_), n = _flat_stats(estimate, target)
    return float(10.0 * math.log10((stt + eps) / (see - 2.0 * set_ + stt + eps)))


def compute_sar(estimate: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> float:
    (se, st, see, stt, set_), n = _flat_stats(estimate, target)
    scale = set_ / (stt + eps)
    signal = scale * scale * stt + eps
    artifact = see - 2.0 * scale * set_ + scale * scale * stt + eps
    return float(10.0 * math.log10(signal / artifact))


def compute_sir(estimate: torch.Tensor, target: torch.Tensor, references=None, eps: float = 1e-8) -> float:
    return compute_sar(estimate, target, eps)      # identical arithmetic in the reference (:128-168)


def compute_all_metrics(estimate: torch.Tensor, target: torch.Tensor) -> Dict[str, float]:
    return {"si_sdr": compute_si_sdr(estimate, target), "sdr": compute_sdr(estimate, target),
            "sar": compute_sar(estimate, target), "sir": compute_sir(estimate, target)}


def evaluate_source_separation(estimates: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor],
                               metrics=("si_sdr", "sdr", "sar", "sir")) -> Dict[str, Dict[str, float]]:
    """evaluation/metrics.py:170-218: per-source metrics plus their average under ``"average"``."""
    fns = {"si_sdr": compute_si_sdr, "sdr": compute_sdr, "sar": compute_sar, "sir": compute_sir}
    results: Dict[str, Dict[str, float]] = {}
    for name, est in estimates.items():
        if name not in targets:
            continue
        results[name] = {m: fns[m](est, targets[name]) for m in metrics if m in fns}
    avg = {}
    for m in metrics:
        values = [results[s][m] for s in results if m in results[s]]
        if values:
            avg[f"avg_{m}"] = float(sum(values) / len(values))
    results["average"] = avg
    return results
