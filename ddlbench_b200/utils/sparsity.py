"""Offline analysis of logged tensors: sparsity and achievable compression ratio, and the
all-reduce timing extractor (reference: ``profiler/image_classification/utils/sparsity/
{measure_sparsity,measure_compression_ratio,plot}.py`` and ``utils/all_reduce/
extract_reduce_times.py``)."""
from __future__ import annotations

import pickle
import re
import sys
from typing import Dict, List, Tuple

import torch


def sparsity(t: torch.Tensor, threshold: float = 0.0) -> float:
    """Fraction of elements with |x| <= threshold."""
    t = t.float()
    return float((t.abs() <= threshold).float().mean())


def measure_sparsity(path: str, thresholds=(0.0, 1e-3)) -> Dict[str, Dict[float, float]]:
    with open(path, "rb") as f:
        tensors = pickle.load(f)
    return {k: {th: sparsity(v, th) for th in thresholds} for k, v in tensors.items()}


def csc_compression_ratio(t: torch.Tensor, threshold: float = 0.0) -> float:
    """Dense bytes / compressed-sparse-column bytes (values + row indices + column pointers),
    the estimate of ``measure_compression_ratio.py:12-22``."""
    m = t.float().reshape(t.shape[0], -1) if t.dim() > 1 else t.float().reshape(1, -1)
    nnz = int((m.abs() > threshold).sum())
    dense = m.numel() * 4
    csc = nnz * 4 + nnz * 4 + (m.shape[1] + 1) * 4
    return dense / max(csc, 1)


def measure_compression_ratio(path: str, threshold: float = 0.0) -> Dict[str, float]:
    with open(path, "rb") as f:
        tensors = pickle.load(f)
    return {k: csc_compression_ratio(v, threshold) for k, v in tensors.items()}


def plot(series: Dict[str, List[float]], out_path: str, ylabel: str = "sparsity") -> None:
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:  # noqa: BLE001 - plotting is optional
        return
    fig, ax = plt.subplots(figsize=(8, 3))
    for k, v in series.items():
        ax.plot(v, label=k)
    ax.set_ylabel(ylabel)
    ax.set_xlabel("layer")
    if len(series) <= 8:
        ax.legend()
    fig.savefig(out_path, bbox_inches="tight")


def extract_reduce_times(lines: List[str]) -> List[Tuple[int, float]]:
    """``allreduce[<bucket>] start=<ms> end=<ms>`` lines (written by FusedDataParallel with
    ``DDLB_LOG_REDUCE_TIMES=1``; the reference parses its patched DDP's
    ``log_reduce_times`` output, ``extract_reduce_times.py:7-40``) -> [(bucket, duration ms)]."""
    out = []
    for ln in lines:
        m = re.search(r"allreduce\[(\d+)\] start=([\d.]+) end=([\d.]+)", ln)
        if m:
            out.append((int(m.group(1)), float(m.group(3)) - float(m.group(2))))
    return out


if __name__ == "__main__":
    for k, v in measure_sparsity(sys.argv[1]).items():
        print(k, v)
