"""Step timeline: per-op forward/backward spans of one (eager) training step as a Chrome-trace
JSON (``chrome://tracing`` / Perfetto), plus optional NVTX ranges for Nsight.

The reference has no device-side tracing at all (host ``time.time()`` around
synchronised layers, no NVTX, no trace export).  Here the executor's per-op ``begin/end``
hooks record CUDA events on the compute stream (resolved once, after the step), so the trace
shows where a pipeline stage or a data-parallel step actually spends its device time and how
large the gaps between ops are.

    tl = StepTimeline(executor)              # executor: engine.StageExecutor
    with tl:                                 # installs itself as executor.profiler
        outs, ctx = executor.forward(...); ...; executor.backward(ctx, [dloss])
    tl.save("step.trace.json")               # also: tl.table() -> rows sorted by time

``DDLB_NVTX=1`` makes every span an NVTX range as well (``torch.cuda.nvtx``).
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict, List, Optional, Tuple

import torch


class StepTimeline:
    def __init__(self, executor, nvtx: Optional[bool] = None, pid: int = 0):
        self.ex = executor
        self.cuda = torch.device(executor.store.device).type == "cuda"
        self.nvtx = (os.environ.get("DDLB_NVTX", "0") == "1") if nvtx is None else nvtx
        self.nvtx = self.nvtx and self.cuda
        self.pid = pid
        self._open: List[Tuple[int, object]] = []
        self._spans: List[Tuple[int, object, object]] = []
        self._t0 = None
        self._prev = None

    # ---- executor.profiler interface ------------------------------------------------
    def begin(self, ident: int):
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            if self._t0 is None:
                self._t0 = ev
        else:
            ev = time.perf_counter()
            if self._t0 is None:
                self._t0 = ev
        if self.nvtx:
            torch.cuda.nvtx.range_push(self.label(ident))
        self._open.append((ident, ev))
        return len(self._open) - 1

    def end(self, h) -> None:
        ident, ev0 = self._open[h]
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
        if self.cuda:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
        else:
            ev1 = time.perf_counter()
        self._spans.append((ident, ev0, ev1))

    def resolve(self):
        return [(i, d) for i, _, d in self.spans()]

    # ---- context manager ------------------------------------------------------------
    def __enter__(self):
        self._prev = self.ex.profiler
        self.ex.profiler = self
        return self

    def __exit__(self, *exc):
        self.ex.profiler = self._prev
        return False

    # ---- results --------------------------------------------------------------------
    def label(self, ident: int) -> str:
        op = self.ex.plan.ops[ident // 2]
        return "%s %s:%s" % ("bwd" if ident % 2 else "fwd", op.kind, op.name)

    def spans(self) -> List[Tuple[int, float, float]]:
        """[(id, start ms, duration ms)] relative to the first span."""
        if not self._spans:
            return []
        if self.cuda:
            self._spans[-1][2].synchronize()
            return [(i, self._t0.elapsed_time(a), a.elapsed_time(b)) for i, a, b in self._spans]
        return [(i, (a - self._t0) * 1e3, (b - a) * 1e3) for i, a, b in self._spans]

    def chrome_trace(self) -> Dict[str, object]:
        events = []
        for ident, start, dur in self.spans():
            events.append({"name": self.label(ident), "ph": "X", "pid": self.pid,
                           "tid": 1 if ident % 2 else 0, "ts": start * 1e3, "dur": dur * 1e3,
                           "cat": "backward" if ident % 2 else "forward"})
        meta = [{"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 0, "args": {"name": "forward"}},
                {"name": "thread_name", "ph": "M", "pid": self.pid, "tid": 1, "args": {"name": "backward"}}]
        return {"traceEvents": meta + events, "displayTimeUnit": "ms"}

    def save(self, path: str) -> str:
        with open(path, "w") as f:
            json.dump(self.chrome_trace(), f)
        return path

    def table(self, top: int = 20) -> str:
        rows = sorted(self.spans(), key=lambda r: -r[2])[:top]
        total = sum(d for _, _, d in self.spans())
        out = ["%-60s %10s %6s" % ("op", "ms", "%")]
        for ident, _, dur in rows:
            out.append("%-60s %10.4f %6.1f" % (self.label(ident)[:60], dur, 100.0 * dur / max(total, 1e-12)))
        return "\n".join(out)
