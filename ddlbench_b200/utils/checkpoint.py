"""Per-stage checkpoints (``checkpoint.<stage>.pth.tar``) and resume.

Reference: ``runtime/image_classification/main_with_runtime.py:240-262,392-403,580-584`` —
``{epoch, arch, state_dict, best_prec1, optimizer}`` written each epoch by the first replica of
every stage; ``--resume <prefix>`` loads ``<prefix>.<stage>.pth.tar``.  The reference does not
save the stashed weight versions (resume restarts with all versions equal); neither do we, but
the fp32 master weights and momentum ARE saved so the optimizer state is exact."""
from __future__ import annotations

import os
from typing import Optional

import torch


def checkpoint_path(directory: str, stage: int) -> str:
    return os.path.join(directory, "checkpoint.%d.pth.tar" % stage)


def save_stage_checkpoint(directory: str, stage: int, epoch: int, arch: str, runtime,
                          best_prec1: float = 0.0, write: bool = True) -> str:
    """Every replica of the stage calls this (``runtime.state_dict()`` gathers sharded optimizer
    state); only the one with ``write=True`` (replica 0) touches the file."""
    payload = runtime.state_dict()
    path = checkpoint_path(directory, stage)
    if write:
        os.makedirs(directory, exist_ok=True)
        payload.update({"epoch": epoch, "arch": arch, "best_prec1": best_prec1})
        tmp = path + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, path)          # a crash mid-write never corrupts the previous checkpoint
    return path


def load_stage_checkpoint(directory: str, stage: int, runtime) -> Optional[dict]:
    path = checkpoint_path(directory, stage)
    if not os.path.isfile(path):
        return None
    payload = torch.load(path, map_location="cpu")
    runtime.load_state_dict(payload)
    return payload


def merge_stage_checkpoints(directory: str, num_stages: int) -> dict:
    """Stitch per-stage checkpoints back into one model state dict (the role of the
    reference's ``compute_bleu_scores.py:168-248`` for GNMT)."""
    merged = {}
    for s in range(num_stages):
        payload = torch.load(checkpoint_path(directory, s), map_location="cpu")
        merged.update(payload["state_dict"])
    return merged
