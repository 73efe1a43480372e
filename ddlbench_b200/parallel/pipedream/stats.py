"""Per-stage runtime counters (``pipedream-fork/runtime/runtime_utilities.py:4-27``): compute
seconds and send/receive seconds + bytes, one instance for the forward and one for the
backward direction.  The reference only ever fills the byte counters
(``runtime.py:423-425,444-446``); here compute time is filled too, from CUDA events recorded on
the stage's stream and resolved lazily at ``print_stats`` (no per-minibatch synchronisation).
"""
from __future__ import annotations

from typing import List, Tuple

import torch


class RuntimeStats:
    KEYS = ("compute_time", "send_tensors", "send_tensors_size", "receive_tensors",
            "receive_tensors_size")

    def __init__(self, forward: bool):
        self.forward = forward
        self.stats = {k: 0.0 for k in self.KEYS}
        self._events: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        # (start, end) of every pass for the timeline export: CUDA events or host seconds
        self.spans: List[Tuple[object, object]] = []
        self.keep_spans = False

    # -- compute time ---------------------------------------------------------------
    def begin_compute(self, device):
        if torch.device(device).type != "cuda":
            import time
            return time.perf_counter()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end_compute(self, token) -> None:
        if isinstance(token, float):
            import time
            now = time.perf_counter()
            self.stats["compute_time"] += now - token
            if self.keep_spans:
                self.spans.append((token, now))
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._events.append((token, e))
        if self.keep_spans:
            self.spans.append((token, e))
        if len(self._events) > 4096 and not self.keep_spans:
            self._resolve()

    def _resolve(self) -> None:
        if self._events:
            self._events[-1][1].synchronize()
            self.stats["compute_time"] += sum(a.elapsed_time(b) for a, b in self._events) / 1e3
            self._events.clear()

    # -- traffic --------------------------------------------------------------------
    def sent(self, tensors) -> None:
        self.stats["send_tensors_size"] += sum(t.numel() * t.element_size() for t in tensors)

    def received(self, tensors) -> None:
        self.stats["receive_tensors_size"] += sum(t.numel() * t.element_size() for t in tensors)

    # -- reference API --------------------------------------------------------------
    def print_stats(self, out=print) -> None:
        self._resolve()
        out("Forward Stats:" if self.forward else "Backward Stats:")
        for k in sorted(self.stats):
            out("\t %s %.3f %s" % (k, self.stats[k], "bytes" if k.endswith("_size") else "seconds"))

    def reset_stats(self) -> None:
        self._events.clear()
        for k in self.stats:
            self.stats[k] = 0.0


def pipeline_trace(forward: RuntimeStats, backward: RuntimeStats, rank: int, stage: int) -> dict:
    """Chrome-trace events of one rank's forward / backward passes (``keep_spans`` must have been
    set before the epoch): load the per-rank files together in Perfetto to see the 1F1B
    interleaving and the bubbles.  Device spans are relative to the rank's first forward."""
    events = []
    spans = [("F", i, a, b) for i, (a, b) in enumerate(forward.spans)] + \
            [("B", i, a, b) for i, (a, b) in enumerate(backward.spans)]
    if not spans:
        return {"traceEvents": []}
    first = forward.spans[0][0] if forward.spans else backward.spans[0][0]
    if not isinstance(first, float):
        spans[-1][3].synchronize()
    for kind, i, a, b in spans:
        if isinstance(a, float):
            ts, dur = (a - first) * 1e6, (b - a) * 1e6
        else:
            ts, dur = first.elapsed_time(a) * 1e3, a.elapsed_time(b) * 1e3
        events.append({"name": "%s%d" % (kind, i), "ph": "X", "pid": stage, "tid": rank, "ts": ts,
                       "dur": dur, "cat": "forward" if kind == "F" else "backward"})
    return {"traceEvents": events, "displayTimeUnit": "ms"}
