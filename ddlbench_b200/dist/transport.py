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

``SymmP2P`` (below, opt-in with ``DDLB_P2P=symm``) is the transport that bypasses NCCL on the data
path: ``p2p_put`` stores the tensor straight into the consumer's symmetric-memory ring slot over
NVLink and raises a sequence flag, ``p2p_wait`` makes the consumer's stream spin on it
(``ops/csrc/allreduce.cu``), credits flow back the same way.  Protocol-tested off-line
(``tests/test_symm_p2p_cpu.py``); HUNG in the 4-GPU hardware run of round 2 (not validated, keep DDLB_P2P=nccl) — ``bench/p2p_bandwidth.py`` measures the
two kernels, ``tools/validate_pipe_graphs.sh`` A/Bs the pipeline with it.
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


# ======================================================================================
# Peer-store transport (opt-in: DDLB_P2P=symm)
# ======================================================================================
_CHANNELS = {"fwd": 0, "bwd": 1}
_ALIGN = 256


def _align(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class _SymmPending:
    """Result of ``SymmP2P.post_recv``: ``wait()`` makes the CURRENT stream wait for the flag
    (one sequence number per tensor) and returns views of the ring slot — no copy."""
    __slots__ = ("owner", "chan", "n", "tensors")

    def __init__(self, owner, chan, n, tensors):
        self.owner, self.chan, self.n, self.tensors = owner, chan, n, tensors

    def wait(self) -> List[torch.Tensor]:
        for _ in range(self.n):
            self.owner._wait_flag(self.chan, "data")
        return self.tensors


class SymmP2P:
    """Stage-boundary transport with NO NCCL on the data path (same interface as ``DistP2P``).

    The sender's ``p2p_put`` kernel stores the tensors straight into the RECEIVER's ring slot
    over NVLink (peer stores into symmetric memory) and then raises a sequence flag there; the
    receiver's compute stream spins on that flag (``p2p_wait``) and uses the slot in place — one
    kernel on each side, no rendezvous, no staging copy, everything replayable from a CUDA graph
    (the sequence counters live in device memory).  Reference: Gloo/NCCL ``send``/``recv`` (or
    ``broadcast`` in 2-rank groups) issued by helper threads, ``runtime/communication.py:635-712``.

    Flow control is by credits: ``post_recv`` — the point where ``DistP2P`` would post its
    ``irecv`` into a ring slot — sends a zero-byte put (= a flag increment) to the sender, and
    ``send`` waits for one credit per message before it writes.  A slot is therefore never
    overwritten before the receiver has declared it free, exactly the contract of
    ``DistP2P.post_recv(ring=...)``.

    Memory: per neighbour pair one symmetric inbox of 2 channels x ``ring`` slots x ``slot_bytes``
    and 8 flag words, allocated collectively (in sorted pair order) at construction.
    """

    def __init__(self, device, rank: int, pairs: Sequence[Tuple[int, int]], slot_bytes: int, ring: int,
                 alloc, ext, blocks: int = 16, use_streams: Optional[bool] = None):
        """``alloc(pair, numel, dtype) -> SymmetricBuffer`` (collective over the two ranks of
        ``pair``, ``ptrs`` indexed by position in the sorted pair); ``ext``: the extension module
        (``p2p_put`` / ``p2p_wait``)."""
        self.device = torch.device(device)
        self.rank, self.ring, self.blocks, self.ext = rank, ring, blocks, ext
        self.slot_bytes = _align(slot_bytes)
        self.gpu = self.device.type == "cuda" if use_streams is None else use_streams
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.gpu else None
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._zero = torch.empty(0, dtype=torch.uint8, device=self.device)
        self.peers: Dict[int, dict] = {}
        for a, b in sorted({(min(x, y), max(x, y)) for x, y in pairs if x != y}):
            if rank not in (a, b):
                continue
            me, other = (0, 1) if rank == a else (1, 0)
            inbox = alloc((a, b), 2 * ring * self.slot_bytes, torch.uint8)
            flags = alloc((a, b), 8, torch.int32)
            self.peers[b if rank == a else a] = {
                "inbox": inbox, "flags": flags, "me": me, "other": other,
                # device-side sequence counters / tickets, one per (channel, kind)
                "seq": {(c, k): torch.zeros(1, dtype=torch.int32, device=self.device)
                        for c in _CHANNELS for k in ("data_rx", "data_tx", "credit_rx", "credit_tx")},
                "ticket": {c: torch.zeros(1, dtype=torch.int32, device=self.device) for c in _CHANNELS},
                "ticket_credit": {c: torch.zeros(1, dtype=torch.int32, device=self.device) for c in _CHANNELS},
                "rx_slot": {c: 0 for c in _CHANNELS}, "tx_slot": {c: 0 for c in _CHANNELS},
            }
        self.bytes_sent = 0
        self.bytes_received = 0

    # ---------------------------------------------------------------- construction on real groups
    @classmethod
    def from_groups(cls, device, groups: PairGroups, slot_bytes: int, ring: int, blocks: int = 16) -> "SymmP2P":
        from .. import ops
        from .symm import alloc_symmetric
        pairs = sorted({(a, b) for (a, b, _d) in groups.groups})

        def alloc(pair, numel, dtype):
            return alloc_symmetric(numel, dtype, device, groups.groups[(pair[0], pair[1], "up")])

        return cls(device, dist.get_rank(), pairs, slot_bytes, ring, alloc, ops.ext(), blocks)

    # ---------------------------------------------------------------- helpers
    @staticmethod
    def _flag_index(channel: str, kind: str) -> int:
        return _CHANNELS[channel] * 2 + (0 if kind == "data" else 1)

    def _wait_flag(self, chan: Tuple[int, str], kind: str) -> None:
        """Enqueue (on the current stream) a wait for the next sequence number of a LOCAL flag."""
        peer, channel = chan
        p = self.peers[peer]
        addr = p["flags"].ptrs[p["me"]] + 4 * self._flag_index(channel, kind)
        self.ext.p2p_wait(addr, p["seq"][(channel, kind + "_rx")], self.status)

    def _raise_peer_flag(self, peer: int, channel: str, kind: str, src: torch.Tensor, dst_addr: int,
                         blocks: int) -> None:
        """Store ``src`` at ``dst_addr`` in the peer's memory (may be empty), then bump its flag."""
        p = self.peers[peer]
        flag = p["flags"].ptrs[p["other"]] + 4 * self._flag_index(channel, kind)
        ticket = p["ticket"][channel] if kind == "data" else p["ticket_credit"][channel]
        self.ext.p2p_put(src, dst_addr, flag, p["seq"][(channel, kind + "_tx")], ticket, blocks)

    def _slot_offset(self, channel: str, slot: int) -> int:
        return (_CHANNELS[channel] * self.ring + slot) * self.slot_bytes

    def _on_comm_stream(self):
        import contextlib
        if not self.gpu:
            return contextlib.nullcontext()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        ctx = torch.cuda.stream(self.comm_stream)
        self.comm_stream.wait_event(ev)
        return ctx

    # ---------------------------------------------------------------- receive
    def post_recv(self, src: int, specs: Sequence[TensorSpec], channel: str,
                  ring: Optional[int] = None) -> _SymmPending:
        """Declare the next ring slot of (src, channel) free and hand the sender a credit for it.
        ``ring`` is accepted for interface parity; the depth is fixed at construction (>= it)."""
        assert ring is None or ring <= self.ring, "SymmP2P was built with a smaller ring"
        p = self.peers[src]
        slot = p["rx_slot"][channel]
        p["rx_slot"][channel] = (slot + 1) % self.ring
        base = self._slot_offset(channel, slot)
        inbox = p["inbox"].tensor
        views, off = [], 0
        for s in specs:
            numel = 1
            for d in s.shape:
                numel *= d
            nbytes = numel * torch.empty((), dtype=s.dtype).element_size()
            flat = inbox[base + off: base + off + nbytes].view(s.dtype)
            if s.channels_last and len(s.shape) == 4:
                n, c, h, w = s.shape
                views.append(flat.view(n, h, w, c).permute(0, 3, 1, 2))
            else:
                views.append(flat.view(s.shape))
            off += _align(nbytes)
        assert off <= self.slot_bytes, "message larger than the slot size SymmP2P was built with"
        with self._on_comm_stream():            # after everything that still reads the slot
            self._raise_peer_flag(src, channel, "credit", self._zero, 0, 1)
        self.bytes_received += off
        return _SymmPending(self, (src, channel), len(specs), views)

    # ---------------------------------------------------------------- send
    def send(self, dst: int, tensors: Sequence[torch.Tensor], channel: str):
        p = self.peers[dst]
        slot = p["tx_slot"][channel]
        p["tx_slot"][channel] = (slot + 1) % self.ring
        base = p["inbox"].ptrs[p["other"]] + self._slot_offset(channel, slot)
        done = None
        with self._on_comm_stream():
            self._wait_flag((dst, channel), "credit")          # the receiver freed this slot
            off = 0
            for t in tensors:
                t = t if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(
                    memory_format=torch.channels_last)) else t.contiguous()
                nbytes = t.numel() * t.element_size()
                assert nbytes % 16 == 0, "p2p_put moves 16-byte vectors"
                if self.gpu:
                    t.record_stream(self.comm_stream)      # keep the source alive for the kernel
                self._raise_peer_flag(dst, channel, "data", t, base + off, self.blocks)
                off += _align(nbytes)
            assert off <= self.slot_bytes, "message larger than the slot size SymmP2P was built with"
            if self.gpu:
                done = torch.cuda.Event()
                done.record(self.comm_stream)
        self.bytes_sent += off
        return done

    def flush(self) -> None:
        """End of an epoch: drain the communication stream and fail loudly if any flag wait gave
        up (``p2p_wait`` stops spinning after ~10 s and records it instead of hanging the GPU)."""
        if self.gpu:
            self.comm_stream.synchronize()
            torch.cuda.current_stream(self.device).synchronize()
        self.check()

    def check(self) -> None:
        s = int(self.status.item())
        if s != 0:
            raise RuntimeError(f"peer-store transport: a flag wait timed out (status={s})")
