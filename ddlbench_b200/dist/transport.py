"""Stage-boundary point-to-point transport (activations forward, gradients backward).

Reference (K5-K9): one Python helper thread per (tensor, peer, direction),
each message = an int32 shape message + payload, ``tensor.detach().clone()`` on send and
``torch.zeros(shape).cuda()`` on every receive, 2-rank ``dist.broadcast`` sub-groups on Gloo
(host-staged even inside a node) — ``pipedream-fork/runtime/communication.py:249-355,635-712``.

Here shapes are static (known from a dry run), so a message is one transfer into a
PRE-ALLOCATED ring of receive buffers; receives are pre-posted one step ahead on a
communication stream so the transfer overlaps the stage's compute; there are no helper threads.

``DistP2P`` — ``torch.distributed`` isend/irecv: NCCL p2p over NVLink on GPUs, Gloo on CPU (the
fake transport that lets the schedules run under pytest on a GPU-less box).  Every
(pair, direction) gets its own 2-rank process group so an outstanding receive can never block a
send in the other direction (the reference needs the same property and creates a group per pair
*per tensor*).  NCCL's rendezvous gives the back-pressure the reference implements with "ack"
tokens.

The building blocks of a transport that bypasses NCCL exist as kernels — ``p2p_put`` (store the
tensor straight into the consumer's symmetric-memory slot over NVLink, then raise a sequence
flag) and ``p2p_wait`` (the consumer's stream spins on the flag), ``ops/csrc/allreduce.cu`` — and
are measured by ``bench/p2p_bandwidth.py``; wiring them in as a second backend (it needs a
consumer->producer acknowledgement for ring reuse) is an open item.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass
class TensorSpec:
    name: str
    shape: Tuple[int, ...]
    dtype: torch.dtype
    channels_last: bool = False

    def empty(self, device) -> torch.Tensor:
        t = torch.empty(self.shape, dtype=self.dtype, device=device)
        if self.channels_last and len(self.shape) == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        return t


class PairGroups:
    """2-rank process groups for every (lower rank, higher rank, direction).  Creation is
    collective over the whole world and happens in the same order on every rank."""

    def __init__(self, pairs: Sequence[Tuple[int, int]]):
        self.groups: Dict[Tuple[int, int, str], object] = {}
        uniq = sorted({(min(a, b), max(a, b)) for a, b in pairs if a != b})
        for a, b in uniq:
            for direction in ("up", "down"):
                self.groups[(a, b, direction)] = dist.new_group([a, b])

    def get(self, src: int, dst: int):
        a, b = min(src, dst), max(src, dst)
        return self.groups[(a, b, "up" if src < dst else "down")]


class _Pending:
    __slots__ = ("works", "tensors")

    def __init__(self, works, tensors):
        self.works, self.tensors = works, tensors

    def wait(self) -> List[torch.Tensor]:
        for w in self.works:
            w.wait()
        return self.tensors


class DistP2P:
    """Send / receive lists of fixed-shape tensors to / from peer ranks."""

    def __init__(self, device, groups: Optional[PairGroups] = None, ring: int = 3):
        self.device = torch.device(device)
        self.groups = groups
        self.rank = dist.get_rank()
        self.ring = ring
        self.gpu = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.gpu else None
        self._bufs: Dict[Tuple[int, str], List[List[torch.Tensor]]] = {}
        self._slot: Dict[Tuple[int, str], int] = {}
        self._inflight_sends: List[Tuple[object, torch.Tensor]] = []
        self.bytes_sent = 0
        self.bytes_received = 0
        self.poison = os.environ.get("DDLB_DEBUG_POISON", "0") == "1"

    def _group(self, src: int, dst: int):
        return self.groups.get(src, dst) if self.groups is not None else None

    # ---------------------------------------------------------------- receive
    def post_recv(self, src: int, specs: Sequence[TensorSpec], channel: str,
                  ring: Optional[int] = None) -> _Pending:
        """Pre-post the receive of one message (all tensors of ``specs``) from ``src``.
        ``ring``: number of receive slots of this channel — must exceed the number of messages
        whose buffers are still referenced (activations saved for a later backward)."""
        key = (src, channel)
        if key not in self._bufs:
            n = ring or self.ring
            self._bufs[key] = [[s.empty(self.device) for s in specs] for _ in range(n)]
            self._slot[key] = 0
        slot = self._slot[key]
        self._slot[key] = (slot + 1) % len(self._bufs[key])
        bufs = self._bufs[key][slot]
        if self.poison:
            # debug mode: a consumer that reads a ring slot before its transfer completed (or a
            # slot that was recycled too early) sees NaN / a sentinel instead of stale, plausible data
            for b in bufs:
                b.fill_(float("nan") if b.is_floating_point() else -1)
        works = []
        g = self._group(src, self.rank)
        if self.gpu:
            with torch.cuda.stream(self.comm_stream):
                for b in bufs:
                    works.append(dist.irecv(b, src=src, group=g))
        else:
            for b in bufs:
                works.append(dist.irecv(b, src=src, group=g))
        self.bytes_received += sum(b.numel() * b.element_size() for b in bufs)
        return _Pending(works, bufs)

    # ---------------------------------------------------------------- send
    def send(self, dst: int, tensors: Sequence[torch.Tensor], channel: str):
        """Asynchronous send; the tensors are kept alive until the transfer completes.  On GPUs
        returns an event (communication stream) that fires once the transfers have been issued
        and read their sources — callers that overwrite ``tensors`` in place wait on it."""
        g = self._group(self.rank, dst)
        self._reap()
        if self.gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                for t in tensors:
                    t = t if t.is_contiguous() or t.is_contiguous(
                        memory_format=torch.channels_last) else t.contiguous()
                    self._inflight_sends.append((dist.isend(t, dst=dst, group=g), t))
                done = torch.cuda.Event()
                done.record(self.comm_stream)
        else:
            done = None
            for t in tensors:
                t = t.contiguous()
                self._inflight_sends.append((dist.isend(t, dst=dst, group=g), t))
        self.bytes_sent += sum(t.numel() * t.element_size() for t in tensors)
        return done

    def _reap(self) -> None:
        self._inflight_sends = [(w, t) for (w, t) in self._inflight_sends if not w.is_completed()]

    def flush(self) -> None:
        for w, _ in self._inflight_sends:
            w.wait()
        self._inflight_sends = []
