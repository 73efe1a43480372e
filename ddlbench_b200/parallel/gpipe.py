"""GPipe: synchronous micro-batch pipeline parallelism with activation re-computation.

Replaces the external ``torchgpipe`` package the reference harness depends on
(``GPipe(model, balance, chunks)``, ``balance_by_time`` — ``benchmark/imagenet/
imagenet_gpipe.py:76-88``; semantics summarised in SURVEY.md §2.7):

* partition ``i`` holds layers ``[sum(balance[:i]), sum(balance[:i+1]))`` of the flat layer
  sequence; skip connections need no stash/pop "portals" because the layer graph knows which
  tensors are live at every cut and they simply travel with the activation;
* a mini-batch is split into ``chunks`` micro-batches; forward of all micro-batches (fill),
  then backward in reverse order (drain), gradients accumulate, ONE synchronous optimizer
  step follows — statistically identical to large-batch SGD;
* ``checkpoint='except_last'`` (torchgpipe's default): every micro-batch but the last drops
  its saved activations after forward and re-computes them right before its backward
  (BatchNorm running statistics are frozen during the re-computation).

Execution model: one process per GPU (partition = rank), stage-boundary tensors move through
``dist.transport.DistP2P`` (NCCL p2p on a communication stream, pre-posted into rings); with no
process group all partitions run in one process (``devices`` may then name several GPUs, the
reference's single-process layout, or a single device for tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from ..dist.transport import DistP2P, PairGroups
from ..engine import StageExecutor
from ..engine import ops as E
from ..models.graph import LayerGraphModel
from .pipedream.runtime import infer_boundary_specs


def _cuts_from_balance(balance: Sequence[int], n_layers: int) -> List[int]:
    assert sum(balance) == n_layers, f"balance {list(balance)} must sum to {n_layers} layers"
    cuts, acc = [], 0
    for b in balance[:-1]:
        acc += b
        cuts.append(acc)
    return cuts


def solve_min_max_partition(costs: Sequence[float], parts: int) -> List[int]:
    """Contiguous partition of ``costs`` into ``parts`` blocks minimising the largest block sum
    (what ``torchgpipe.balance`` solves).  Returns the block sizes."""
    n = len(costs)
    parts = max(1, min(parts, n))
    pre = [0.0]
    for c in costs:
        pre.append(pre[-1] + c)
    INF = float("inf")
    best = [[INF] * (n + 1) for _ in range(parts + 1)]
    arg = [[0] * (n + 1) for _ in range(parts + 1)]
    best[0][0] = 0.0
    for p in range(1, parts + 1):
        for j in range(p, n + 1):
            for k in range(p - 1, j):
                v = max(best[p - 1][k], pre[j] - pre[k])
                if v < best[p][j]:
                    best[p][j], arg[p][j] = v, k
    sizes, j = [], n
    for p in range(parts, 0, -1):
        k = arg[p][j]
        sizes.append(j - k)
        j = k
    return sizes[::-1]


def balance_by_time(partitions: int, model: LayerGraphModel, sample: torch.Tensor,
                    device=None, dtype: Optional[torch.dtype] = None, trials: int = 3) -> List[int]:
    """Profile forward+backward time of every layer on one device with ``sample`` and split the
    layer sequence into ``partitions`` contiguous blocks of (near) equal time.  Device-timed
    with CUDA events on GPUs (reference: ``torchgpipe.balance.balance_by_time`` times each
    child with host clocks + synchronize for ~1 s)."""
    from ..utils.layer_profiler import profile_layers
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    prof = profile_layers(model, sample, device=dev, dtype=dtype, trials=trials)
    costs = [f + b for f, b in zip(prof.forward_ms, prof.backward_ms)]
    return solve_min_max_partition(costs, partitions)


def balance_by_size(partitions: int, model: LayerGraphModel) -> List[int]:
    """Split by parameter bytes + a per-layer constant (cheap fallback, no device needed)."""
    costs = [1.0 + sum(p.numel() for p in model.node_module(i).parameters()) / 1e5
             for i in range(len(model.nodes))]
    return solve_min_max_partition(costs, partitions)


class GPipe:
    def __init__(self, model: LayerGraphModel, balance: Sequence[int], chunks: int = 1,
                 checkpoint: str = "except_last", devices: Optional[Sequence] = None,
                 dtype: Optional[torch.dtype] = None, lr: float = 0.1, momentum: float = 0.9,
                 weight_decay: float = 0.0, batch_size: Optional[int] = None,
                 tcgen05: bool = False, groups: Optional[PairGroups] = None,
                 sample_input: Optional[torch.Tensor] = None):
        assert checkpoint in ("always", "except_last", "never")
        if isinstance(model, torch.nn.Sequential) and not isinstance(model, LayerGraphModel):
            # torchgpipe-style model: an nn.Sequential, possibly with @skippable layers
            # (stash / pop skip connections).  Needs one sample batch to resolve the skips.
            from .skip import sequential_to_graph
            if sample_input is None:
                raise ValueError("GPipe(nn.Sequential, ...) needs sample_input=<one input batch> to "
                                 "convert the model (and its stash/pop skips) into a layer graph")
            model = sequential_to_graph(model, sample_input)
        self.model = model
        self.balance = list(balance)
        self.chunks = chunks
        self.checkpoint = checkpoint
        self.cuts = _cuts_from_balance(self.balance, len(model.nodes))
        self.num_partitions = len(self.balance)
        self.distributed = dist.is_initialized() and dist.get_world_size() > 1
        bounds = [0] + self.cuts + [len(model.nodes)]
        if self.distributed:
            assert dist.get_world_size() == self.num_partitions, \
                "one process per partition: world size must equal len(balance)"
            self.rank = dist.get_rank()
            dev = devices[self.rank] if devices else (
                torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                else torch.device("cpu"))
            self._devices = [torch.device(dev)]
            self.local_parts = [self.rank]
        else:
            self.rank = 0
            if devices is None:
                devices = ["cuda" if torch.cuda.is_available() else "cpu"] * self.num_partitions
            self._devices = [torch.device(d) for d in devices]
            self.local_parts = list(range(self.num_partitions))
        self.execs: Dict[int, StageExecutor] = {}
        for k, p in enumerate(self.local_parts):
            dev = self._devices[k if not self.distributed else 0]
            dt = dtype or (torch.bfloat16 if dev.type == "cuda" else torch.float32)
            self.execs[p] = StageExecutor.from_model(
                model, bounds[p], bounds[p + 1], device=dev, dtype=dt, lr=lr, momentum=momentum,
                weight_decay=weight_decay, tcgen05=tcgen05, first_stage=(p == 0))
        self.dtype = next(iter(self.execs.values())).dtype
        self.p2p: Optional[DistP2P] = None
        self._specs = None
        self._micro = None
        if self.distributed:
            if groups is None:
                groups = PairGroups([(r, r + 1) for r in range(self.num_partitions - 1)])
            self.p2p = DistP2P(self._devices[0], groups)
        self.last_stats: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ torchgpipe-like surface
    @property
    def devices(self) -> List[torch.device]:
        return list(self._devices)

    def parameters(self):
        for ex in self.execs.values():
            for v in ex.store.weights().values():
                yield v

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd: Dict[str, torch.Tensor] = {}
        for ex in self.execs.values():
            sd.update(ex.state_dict())
        return sd

    def set_lr(self, lr: float) -> None:
        for ex in self.execs.values():
            ex.store.set_lr(lr)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _rng_get(device):
        """RNG state a forward pass on ``device`` draws from (dropout masks)."""
        device = torch.device(device)
        if device.type == "cuda":
            return (torch.get_rng_state(), torch.cuda.get_rng_state(device))
        return (torch.get_rng_state(), None)

    @staticmethod
    def _rng_set(device, state) -> None:
        device = torch.device(device)
        torch.set_rng_state(state[0])
        if state[1] is not None:
            torch.cuda.set_rng_state(state[1], device)

    def _recompute(self, ex, inp, rng):
        """Re-run a checkpointed forward EXACTLY: same dropout masks (RNG state of the original
        pass restored, as torchgpipe's checkpointing does), BatchNorm running statistics untouched;
        the generator continues afterwards from where it was."""
        now = self._rng_get(ex.device)
        self._rng_set(ex.device, rng)
        E.BN_MOMENTUM_SCALE[0] = 0.0
        try:
            _, ctx = ex.forward(inp, training=True)
        finally:
            E.BN_MOMENTUM_SCALE[0] = 1.0
            self._rng_set(ex.device, now)
        return ctx

    def _keep_ctx(self, mb: int) -> bool:
        if self.checkpoint == "never":
            return True
        if self.checkpoint == "always":
            return False
        return mb == self.chunks - 1

    def _ensure_specs(self, micro: int) -> None:
        if self._micro == micro:
            return
        ex = next(iter(self.execs.values()))
        self._specs = infer_boundary_specs(self.model, self.cuts, micro, self.dtype,
                                           ex.channels_last)
        self._micro = micro

    # ------------------------------------------------------------------ one training step
    def train_step(self, x: Optional[torch.Tensor], target: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Forward all micro-batches, backward all micro-batches, one optimizer step.
        Returns stats[3] (mean loss over the mini-batch, #top1, #top5) on the rank that owns
        the last partition (None elsewhere).  ``x`` is needed on the first partition's rank and
        ``target`` on the last partition's rank (reference: input -> first device, target ->
        last device, ``imagenet_gpipe.py:169-170``)."""
        from .. import ops
        for ex in self.execs.values():      # one tick per step: recomputed micro-batches see the same masks
            ops.rng_tick(ex.device)
        if self.distributed:
            return self._step_distributed(x, target)
        return self._step_local(x, target)

    def _step_local(self, x, target):
        M, P = self.chunks, self.num_partitions
        xs = x.chunk(M)
        ts = target.chunk(M)
        saved: list = []      # [mb][p] -> (ctx or None, inputs, rng state of the checkpointed pass)
        dls, stats_acc = [], None
        for mb in range(M):
            row = []
            cur = [self.execs[0].prepare_input(xs[mb])]
            for p in range(P):
                ex = self.execs[p]
                cur = [c.to(ex.device, non_blocking=True) for c in cur]
                keep = self._keep_ctx(mb)
                rng = None if keep else self._rng_get(ex.device)
                outs, ctx = ex.forward(cur, training=True)
                row.append((ctx if keep else None, cur, rng))
                cur = outs
            st, dl = self.execs[P - 1].loss(cur[0], ts[mb].to(self.execs[P - 1].device),
                                            grad_scale=1.0 / M)
            stats_acc = st.clone() if stats_acc is None else stats_acc + st
            dls.append(dl)
            saved.append(row)
        first_bwd = True
        for mb in reversed(range(M)):
            grads = [dls[mb]]
            for p in reversed(range(P)):
                ex = self.execs[p]
                ctx, inp, rng = saved[mb][p]
                if ctx is None:
                    ctx = self._recompute(ex, inp, rng)
                grads = [g.to(ex.device) if g is not None else None for g in grads]
                grads = ex.backward(ctx, grads, accumulate=not first_bwd, final=(mb == 0))
            first_bwd = False
            saved[mb] = None
        for ex in self.execs.values():
            ex.store.sgd_step()
        stats_acc[0] /= M
        self.last_stats = stats_acc
        return stats_acc

    def _step_distributed(self, x, target):
        M, P, r = self.chunks, self.num_partitions, self.rank
        ex = self.execs[r]
        first, last = r == 0, r == P - 1
        micro = (x.shape[0] // M) if first else None
        if micro is None:
            micro = self._micro_hint
        self._ensure_specs(micro)
        in_specs = self._specs[r - 1] if not first else None
        g_specs = self._specs[r] if not last else None
        ring = M + 2
        if first:
            xs = x.chunk(M)
        if last:
            ts = target.chunk(M)
        saved, dls, stats_acc = [], [], None
        pending = None if first else self.p2p.post_recv(r - 1, in_specs, "fwd", ring=ring)
        for mb in range(M):
            if first:
                cur = [ex.prepare_input(xs[mb])]
            else:
                cur = list(pending.wait())
                pending = self.p2p.post_recv(r - 1, in_specs, "fwd", ring=ring) if mb + 1 < M else None
            keep = self._keep_ctx(mb)
            rng = None if keep else self._rng_get(ex.device)
            outs, ctx = ex.forward(cur, training=True)
            saved.append((ctx if keep else None, cur, rng))
            if last:
                st, dl = ex.loss(outs[0], ts[mb].to(ex.device), grad_scale=1.0 / M)
                stats_acc = st.clone() if stats_acc is None else stats_acc + st
                dls.append(dl)
            else:
                self.p2p.send(r + 1, outs, "fwd")
        pending = None if last else self.p2p.post_recv(r + 1, g_specs, "bwd", ring=3)
        first_bwd = True
        for mb in reversed(range(M)):
            if last:
                grads = [dls[mb]]
            else:
                grads = list(pending.wait())
                pending = self.p2p.post_recv(r + 1, g_specs, "bwd", ring=3) if mb > 0 else None
            ctx, inp, rng = saved[mb]
            if ctx is None:
                ctx = self._recompute(ex, inp, rng)
            gin = ex.backward(ctx, grads, accumulate=not first_bwd, final=(mb == 0))
            first_bwd = False
            saved[mb] = None
            if not first:
                self.p2p.send(r - 1, gin, "bwd")
        ex.store.sgd_step()
        self.p2p.flush()
        if last:
            stats_acc[0] /= M
            self.last_stats = stats_acc
            return stats_acc
        return None

    _micro_hint: int = 0

    def set_micro_batch(self, micro: int) -> None:
        """Ranks that do not see the input need the micro-batch size for shape inference."""
        self._micro_hint = micro

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def eval_step(self, x: Optional[torch.Tensor], target: Optional[torch.Tensor]):
        if not self.distributed:
            cur = [self.execs[0].prepare_input(x)]
            for p in range(self.num_partitions):
                ex = self.execs[p]
                cur, _ = ex.forward([c.to(ex.device) for c in cur], training=False)
            st, _ = self.execs[self.num_partitions - 1].loss(cur[0], target.to(cur[0].device),
                                                             want_grad=False)
            return st
        r, P = self.rank, self.num_partitions
        ex = self.execs[r]
        micro = x.shape[0] if r == 0 else self._micro_hint
        self._ensure_specs(micro)
        if r == 0:
            cur = [ex.prepare_input(x)]
        else:
            cur = list(self.p2p.post_recv(r - 1, self._specs[r - 1], "efwd", ring=2).wait())
        outs, _ = ex.forward(cur, training=False)
        st = None
        if r == P - 1:
            st, _ = ex.loss(outs[0], target.to(ex.device), want_grad=False)
        else:
            self.p2p.send(r + 1, outs, "efwd")
            self.p2p.flush()
        return st
