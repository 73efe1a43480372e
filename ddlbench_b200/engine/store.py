"""Flat parameter / gradient / optimizer-state storage.

All parameters of a model (or of one pipeline stage) live in ONE flat weight buffer, all
gradients in ONE flat gradient buffer, with fp32 master weights and momentum kept by the
optimizer.  Consequences:

* the optimizer is one kernel launch (``csrc/sgd.cu``) instead of the reference's
  per-tensor ``torch.optim.SGD`` loop + per-tensor ``zero_grad`` memsets
  (``benchmark/imagenet/imagenet_pytorch.py:155-157``);
* gradient buckets are contiguous ranges of the flat buffer, ordered in REVERSE execution
  order so backward completes them front to back and the fused all-reduce+SGD kernel
  (``csrc/allreduce.cu``) can be launched bucket by bucket while backward is still running
  (Horovod gets this with a background negotiation thread and a 64 MB fusion buffer,
  SURVEY.md §2.7);
* PipeDream weight stashing is ``V`` weight buffers and an index — the reference deep-copies
  ``state_dict``s (``pipedream-fork/runtime/optimizer.py:58-116,157-159``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

ALIGN = 8  # elements; 16 bytes for bf16 — keeps every view TMA / vector-load aligned


@dataclass
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    init: torch.Tensor            # fp32 initial value, logical (PyTorch) layout
    channels_last: bool = False   # store conv filters as [Cout, kh, kw, Cin] (KRSC)
    order: int = 0                # execution order of the owning op (forward)


@dataclass
class FlatEntry:
    name: str
    offset: int
    numel: int
    shape: Tuple[int, ...]
    channels_last: bool


@dataclass
class Bucket:
    offset: int
    length: int                   # multiple of world * ALIGN
    params: List[str] = field(default_factory=list)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FlatLayout:
    """Offsets of every parameter in the flat buffers + the bucket partition."""

    def __init__(self, specs: Sequence[ParamSpec], bucket_elems: int = 8 << 20, world: int = 1,
                 reverse: bool = True):
        ordered = sorted(specs, key=lambda s: s.order, reverse=reverse)
        self.entries: Dict[str, FlatEntry] = {}
        self.buckets: List[Bucket] = []
        self.world = world
        quantum = world * ALIGN
        off = 0
        cur = Bucket(offset=0, length=0)
        for s in ordered:
            n = 1
            for d in s.shape:
                n *= d
            self.entries[s.name] = FlatEntry(s.name, off, n, tuple(s.shape), s.channels_last)
            cur.params.append(s.name)
            off += _round_up(n, ALIGN)
            if off - cur.offset >= bucket_elems:
                off = _round_up(off, quantum)
                cur.length = off - cur.offset
                self.buckets.append(cur)
                cur = Bucket(offset=off, length=0)
        if cur.params:
            off = _round_up(off, quantum)
            cur.length = off - cur.offset
            self.buckets.append(cur)
        self.total = off
        self.order = [s.name for s in ordered]
        # bucket index of each parameter, for "bucket complete" triggers during backward
        self.bucket_of: Dict[str, int] = {}
        for bi, b in enumerate(self.buckets):
            for p in b.params:
                self.bucket_of[p] = bi

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        e = self.entries[name]
        t = flat[e.offset:e.offset + e.numel]
        if e.channels_last and len(e.shape) == 4:
            co, ci, kh, kw = e.shape
            return t.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return t.view(e.shape)


Allocator = Callable[[int, torch.dtype], torch.Tensor]


class ParamStore:
    """Weights (``versions`` slots), gradients, fp32 master + momentum, hyper-parameters."""

    def __init__(self, specs: Sequence[ParamSpec], device, dtype: torch.dtype,
                 versions: int = 1, bucket_elems: int = 8 << 20, world: int = 1, rank: int = 0,
                 shard_optimizer: bool = False, allocator: Optional[Allocator] = None,
                 lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 0.0,
                 grad_scale: float = 1.0):
        self.device = torch.device(device)
        self.dtype = dtype
        self.world, self.rank = world, rank
        self.layout = FlatLayout(specs, bucket_elems=bucket_elems, world=world)
        n = self.layout.total
        alloc = allocator or (lambda numel, dt: torch.zeros(numel, dtype=dt, device=self.device))
        self.versions = versions
        self.w_flat = [alloc(n, dtype) for _ in range(versions)]
        self.g_flat = alloc(n, dtype)
        self.shard_optimizer = shard_optimizer and world > 1
        m = n // world if self.shard_optimizer else n
        self.master = torch.zeros(m, dtype=torch.float32, device=self.device)
        self.mom = torch.zeros(m, dtype=torch.float32, device=self.device)
        self.hyper = torch.tensor([lr, momentum, weight_decay, grad_scale], dtype=torch.float32,
                                  device=self.device)
        self._w_views = [{name: self.layout.view(wf, name) for name in self.layout.entries}
                         for wf in self.w_flat]
        self.g_views = {name: self.layout.view(self.g_flat, name) for name in self.layout.entries}
        self.latest = 0  # index of the newest weight version
        self._init(specs)

    # ---------------------------------------------------------------- init / access
    def _init(self, specs: Sequence[ParamSpec]) -> None:
        full = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
        for s in specs:
            v = self.layout.view(full, s.name)
            v.copy_(s.init.to(self.device, torch.float32))
        self.load_full_master(full)
        # kept until the first broadcast / step so that a sharded store can hand out the
        # full-precision initial master (FusedDataParallel.broadcast_from)
        self._init_full = full if self.shard_optimizer else None

    def load_full_master(self, full: torch.Tensor) -> None:
        """Set every weight version and the (possibly sharded) master from a full fp32 copy."""
        for wf in self.w_flat:
            wf.copy_(full.to(self.dtype))
        if self.shard_optimizer:
            for b, (mo, ln) in zip(self.layout.buckets, self.master_slices()):
                sl = b.length // self.world
                src = full[b.offset + self.rank * sl: b.offset + (self.rank + 1) * sl]
                self.master[mo:mo + ln].copy_(src)
        else:
            self.master.copy_(full)

    def master_slices(self) -> List[Tuple[int, int]]:
        """(offset, length) of each bucket's slice inside the sharded master buffer."""
        out, off = [], 0
        for b in self.layout.buckets:
            sl = b.length // self.world if self.shard_optimizer else b.length
            out.append((off, sl))
            off += sl
        return out

    def gather_full_state(self, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full (unsharded) fp32 master and momentum.  With a sharded optimizer every rank of the
        replica group must call this (it all-gathers the per-bucket slices)."""
        if not self.shard_optimizer:
            return self.master, self.mom
        import torch.distributed as dist
        out = []
        for src in (self.master, self.mom):
            parts = [torch.empty_like(src) for _ in range(self.world)]
            dist.all_gather(parts, src.contiguous(), group=group)
            full = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
            for b, (mo, ln) in zip(self.layout.buckets, self.master_slices()):
                for r in range(self.world):
                    full[b.offset + r * ln: b.offset + (r + 1) * ln].copy_(parts[r][mo:mo + ln])
            out.append(full)
        return out[0], out[1]

    def load_full_state(self, master: torch.Tensor, mom: torch.Tensor) -> None:
        """Inverse of ``gather_full_state``: take this rank's slices of full master / momentum and
        refresh every weight version from the master."""
        master = master.to(self.device, torch.float32)
        mom = mom.to(self.device, torch.float32)
        assert master.numel() == self.layout.total == mom.numel(), "checkpoint / layout size mismatch"
        self.load_full_master(master)
        if self.shard_optimizer:
            for b, (mo, ln) in zip(self.layout.buckets, self.master_slices()):
                self.mom[mo:mo + ln].copy_(mom[b.offset + self.rank * ln: b.offset + (self.rank + 1) * ln])
        else:
            self.mom.copy_(mom)

    def weights(self, version: Optional[int] = None) -> Dict[str, torch.Tensor]:
        return self._w_views[self.latest if version is None else version]

    def set_lr(self, lr: float) -> None:
        self.hyper[0] = lr

    def set_grad_scale(self, s: float) -> None:
        self.hyper[3] = s

    # ---------------------------------------------------------------- optimizer
    def sgd_step(self, out_version: Optional[int] = None, bucket: Optional[int] = None) -> None:
        """Fused SGD on the whole buffer or one bucket (non-sharded path)."""
        from .. import ops
        assert not self.shard_optimizer, "sharded state is updated by the fused all-reduce kernel"
        out_version = self.latest if out_version is None else out_version
        if bucket is None:
            lo, hi = 0, self.layout.total
        else:
            b = self.layout.buckets[bucket]
            lo, hi = b.offset, b.offset + b.length
        ops.sgd_flat_(self.master[lo:hi], self.mom[lo:hi], self.g_flat[lo:hi], self.hyper,
                      self.w_flat[out_version][lo:hi])

    def advance_version(self) -> int:
        """Weight-stashing ring: the next step writes into slot (latest + 1) % V."""
        self.latest = (self.latest + 1) % self.versions
        return self.latest

    def next_version(self) -> int:
        return (self.latest + 1) % self.versions

    # ---------------------------------------------------------------- (de)serialisation
    def state_dict(self) -> Dict[str, torch.Tensor]:
        w = self.weights()
        return {k: v.detach().float().contiguous().cpu() for k, v in w.items()}

    def full_master(self) -> torch.Tensor:
        assert not self.shard_optimizer
        return self.master

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        full = torch.zeros(self.layout.total, dtype=torch.float32, device=self.device)
        for name in self.layout.entries:
            self.layout.view(full, name).copy_(sd[name].to(self.device, torch.float32))
        self.load_full_master(full)
