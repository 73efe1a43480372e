"""Static, verifiable schedules: the 1F1B order of forward/backward passes on a stage replica
and the replica-to-replica routing of every minibatch.

The reference realises the routing dynamically: per-(tensor, peer) queues fed by helper
threads, a ``messaging_schedule`` walked with mutable row/column cursors, and a deadlock that
was only found in production (``pipedream-fork/runtime/communication.py:455-521,583-608``; the
``>=`` fix at ``:483-487``).  Here the same routing is a pure function of
(config, rank, minibatch index), so a validator can prove — without any GPU or process —
that every send has exactly one matching receive in the same per-channel order.

Routing rules reproduced (replica index arithmetic, not code):
  * forward send of the f-th forward on replica r  -> next-stage replica (f + r) mod n_next
  * backward recv of the b-th backward on replica r <- next-stage replica (b + r) mod n_next
  * forward recv / backward send follow the "messaging schedule": rows[i] = {i, i+n, i+2n, ...}
    (< n_prev) for i < n; a replica starts at row min(r, len(rows)-1), walks a row left to
    right, then moves to the previous row (wrapping).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from .config import PipelineConfig, RankTopology

F, B = "F", "B"


def one_f_one_b(num_iterations: int, warmup: int) -> List[Tuple[str, int]]:
    """[(op, minibatch index)]: ``warmup`` forwards, then strict forward/backward
    alternation, then the remaining backwards
    (``runtime/image_classification/main_with_runtime.py:432-494``)."""
    warmup = min(warmup, num_iterations)
    seq: List[Tuple[str, int]] = [(F, i) for i in range(warmup)]
    b = 0
    for f in range(warmup, num_iterations):
        seq.append((F, f))
        seq.append((B, b))
        b += 1
    while b < num_iterations:
        seq.append((B, b))
        b += 1
    return seq


def fill_drain(num_microbatches: int) -> List[Tuple[str, int]]:
    """GPipe: all forwards, then all backwards in reverse order."""
    return [(F, i) for i in range(num_microbatches)] + \
           [(B, i) for i in reversed(range(num_microbatches))]


class MessagingCursor:
    """Order in which replica ``r`` of a stage with ``n`` replicas talks to the ``n_prev``
    replicas of the previous stage (forward receives; backward sends mirror it)."""

    def __init__(self, rank_in_stage: int, n: int, n_prev: int):
        self.rows: List[List[int]] = []
        for i in range(n):
            row = list(range(i, n_prev, n))
            if row:
                self.rows.append(row)
        self.row = rank_in_stage
        # previous stage has fewer replicas than this one (the reference's `>=` fix)
        while self.rows and self.row >= len(self.rows):
            self.row -= 1
        self.col = 0

    def current(self) -> int:
        return self.rows[self.row][self.col]

    def advance(self) -> None:
        self.col += 1
        if self.col == len(self.rows[self.row]):
            self.col = 0
            self.row -= 1
            if self.row == -1:
                self.row = len(self.rows) - 1

    def take(self) -> int:
        v = self.current()
        self.advance()
        return v


@dataclass
class Routing:
    """Peer RANK (global) of every forward / backward transfer of one replica."""
    topo: RankTopology

    def __post_init__(self):
        t = self.topo
        self._fwd_recv = MessagingCursor(t.rank_in_stage, t.num_ranks_in_stage, len(t.prev_ranks)) \
            if t.prev_ranks else None
        self._bwd_send = MessagingCursor(t.rank_in_stage, t.num_ranks_in_stage, len(t.prev_ranks)) \
            if t.prev_ranks else None

    def fwd_send_peer(self, f: int) -> Optional[int]:
        t = self.topo
        return t.next_ranks[(f + t.rank_in_stage) % len(t.next_ranks)] if t.next_ranks else None

    def bwd_recv_peer(self, b: int) -> Optional[int]:
        t = self.topo
        return t.next_ranks[(b + t.rank_in_stage) % len(t.next_ranks)] if t.next_ranks else None

    def next_fwd_recv_peer(self) -> Optional[int]:
        return self.topo.prev_ranks[self._fwd_recv.take()] if self._fwd_recv else None

    def next_bwd_send_peer(self) -> Optional[int]:
        return self.topo.prev_ranks[self._bwd_send.take()] if self._bwd_send else None


def validate(cfg: PipelineConfig, loader_size: int) -> Dict[str, int]:
    """Dry-run every rank's schedule and check that, on every directed channel
    (src rank, dst rank, direction), the sender's sequence of minibatch transfers has the
    same length as the receiver's expectation, and that replaying all schedules with blocking
    receives never deadlocks.  Returns transfer counts."""
    topos = {r: RankTopology.of(cfg, r) for r in range(cfg.num_ranks)}
    seqs = {r: one_f_one_b(cfg.num_iterations(t.stage, loader_size), t.warmup)
            for r, t in topos.items()}
    routes = {r: Routing(t) for r, t in topos.items()}
    pc = {r: 0 for r in topos}
    chan: Dict[Tuple[int, int, str], int] = {}      # messages in flight per channel
    sent: Dict[Tuple[int, int, str], int] = {}
    pending_peer: Dict[int, Optional[int]] = {r: None for r in topos}
    progressed = True
    while progressed:
        progressed = False
        for r, t in topos.items():
            while pc[r] < len(seqs[r]):
                op, i = seqs[r][pc[r]]
                if op == F:
                    if not t.is_first:
                        if pending_peer[r] is None:
                            pending_peer[r] = routes[r].next_fwd_recv_peer()
                        key = (pending_peer[r], r, "fwd")
                        if chan.get(key, 0) == 0:
                            break
                        chan[key] -= 1
                        pending_peer[r] = None
                    if not t.is_last:
                        key = (r, routes[r].fwd_send_peer(i), "fwd")
                        chan[key] = chan.get(key, 0) + 1
                        sent[key] = sent.get(key, 0) + 1
                else:
                    if not t.is_last:
                        key = (routes[r].bwd_recv_peer(i), r, "bwd")
                        if chan.get(key, 0) == 0:
                            break
                        chan[key] -= 1
                    if not t.is_first:
                        key = (r, routes[r].next_bwd_send_peer(), "bwd")
                        chan[key] = chan.get(key, 0) + 1
                        sent[key] = sent.get(key, 0) + 1
                pc[r] += 1
                progressed = True
    stuck = {r: seqs[r][pc[r]] for r in topos if pc[r] < len(seqs[r])}
    if stuck:
        raise RuntimeError(f"schedule deadlocks: {stuck}")
    left = {k: v for k, v in chan.items() if v}
    if left:
        raise RuntimeError(f"unconsumed messages: {left}")
    return {"transfers": sum(sent.values()), "channels": len(sent)}
