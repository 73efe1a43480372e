"""Horovod-style synchronous data parallelism.

Two layers:

1. A drop-in subset of the ``horovod.torch`` API the reference harness uses
   (``benchmark/imagenet/imagenet_horovod.py:61-143,278-281``): ``init / rank / size /
   local_rank / DistributedOptimizer / broadcast_parameters / broadcast_optimizer_state /
   allreduce`` with ``Average`` / ``Sum``.  It works on any ``nn.Module`` + ``torch.optim``
   optimizer, on NCCL (GPU) or Gloo (CPU tests): per-parameter post-accumulate hooks fill a
   fusion buffer in reverse registration order and launch one asynchronous all-reduce per
   full buffer; ``step()`` waits for the handles and runs the wrapped optimizer.  This is
   the *baseline* path (what Horovod does).

2. ``FusedDataParallel`` — the product path: the model runs on the static engine with its
   flat weight/gradient buffers in symmetric (peer-mapped) memory; as backward completes a
   gradient bucket the engine fires a hook and ONE hand-written kernel
   (``csrc/allreduce.cu``) reduces the bucket across GPUs over NVLink (peer loads, or
   ``multimem.ld_reduce`` in the NVSwitch), applies SGD-momentum to this rank's shard of the
   fp32 master weights and broadcasts the new bf16 weights to every rank (peer stores /
   ``multimem.st``), on a communication stream overlapped with the rest of backward.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

Average = "average"
Sum = "sum"

_state = {"initialized": False, "local_rank": 0, "group": None}


# ------------------------------------------------------------------------------- hvd API
def init(backend: Optional[str] = None, group=None) -> None:
    """Initialise from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT); a no-op single-process world if they are absent."""
    if _state["initialized"]:
        return
    if not dist.is_initialized() and "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        lr = int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
        # watchdog: a rank that dies leaves its peers blocked in a collective; they give up after
        # DDLB_DIST_TIMEOUT_S seconds (default 30 min; the reference sets 120 min,
        # runtime/communication.py:41-43) and the launcher then stops the rest of the job
        import datetime
        kw = {"timeout": datetime.timedelta(seconds=float(os.environ.get("DDLB_DIST_TIMEOUT_S", "1800")))}
        if backend == "nccl":
            torch.cuda.set_device(lr)
            kw["device_id"] = torch.device("cuda", lr)
        dist.init_process_group(backend=backend, **kw)
    _state["local_rank"] = int(os.environ.get("LOCAL_RANK", 0))
    _state["group"] = group
    _state["initialized"] = True


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
    _state["initialized"] = False


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(_state["group"]) > 1


def rank() -> int:
    return dist.get_rank(_state["group"]) if dist.is_initialized() else 0


def size() -> int:
    return dist.get_world_size(_state["group"]) if dist.is_initialized() else 1


def local_rank() -> int:
    return _state["local_rank"]


def allreduce(tensor: torch.Tensor, name: Optional[str] = None, op: str = Average) -> torch.Tensor:
    """Out-of-place all-reduce (averaging by default), as ``hvd.allreduce``."""
    if not is_distributed():
        return tensor.clone()
    t = tensor.clone()
    if dist.get_backend(_state["group"]) == "nccl" and not t.is_cuda:
        t = t.cuda()
    dist.all_reduce(t, group=_state["group"])
    if op == Average:
        t = t / size()
    return t.to(tensor.device)


def broadcast_parameters(params, root_rank: int = 0) -> None:
    """Broadcast a ``state_dict`` / ``named_parameters`` / list of tensors from ``root_rank``."""
    if not is_distributed():
        return
    if isinstance(params, dict):
        tensors = [v for _, v in sorted(params.items()) if torch.is_tensor(v)]
    else:
        tensors = [p[1] if isinstance(p, tuple) else p for p in params]
    for t in tensors:
        dist.broadcast(t.data if hasattr(t, "data") else t, src=root_rank, group=_state["group"])


def broadcast_optimizer_state(optimizer, root_rank: int = 0) -> None:
    if not is_distributed():
        return
    opt = getattr(optimizer, "_opt", optimizer)
    for grp in opt.param_groups:
        for p in grp["params"]:
            st = opt.state.get(p, {})
            for k in sorted(st):
                if torch.is_tensor(st[k]):
                    dist.broadcast(st[k], src=root_rank, group=_state["group"])
    # scalar hyper-parameters
    objs = [[{k: v for k, v in g.items() if k != "params"} for g in opt.param_groups]]
    dist.broadcast_object_list(objs, src=root_rank, group=_state["group"])
    for g, new in zip(opt.param_groups, objs[0]):
        g.update(new)


class DistributedOptimizer:
    """``hvd.DistributedOptimizer`` work-alike (baseline path)."""

    def __init__(self, optimizer: torch.optim.Optimizer, named_parameters=None,
                 backward_passes_per_step: int = 1, op: str = Average,
                 fusion_bytes: int = 64 << 20):
        self._opt = optimizer
        self._op = op
        self._bpps = backward_passes_per_step
        self._params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        self._fusion_bytes = fusion_bytes
        self._pending: List[Tuple[object, torch.Tensor, List[torch.Tensor]]] = []
        self._ready: List[torch.Tensor] = []
        self._ready_bytes = 0
        self._count: Dict[int, int] = {}
        self._hooks = []
        self._stream = None
        if is_distributed():
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))

    # torch.optim passthroughs
    @property
    def param_groups(self):
        return self._opt.param_groups

    @property
    def state(self):
        return self._opt.state

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)

    def zero_grad(self, set_to_none: bool = False):
        return self._opt.zero_grad(set_to_none=set_to_none)

    def _hook(self, p: torch.Tensor) -> None:
        c = self._count.get(id(p), 0) + 1
        self._count[id(p)] = c
        if c % self._bpps:
            return
        self._ready.append(p.grad)
        self._ready_bytes += p.grad.numel() * p.grad.element_size()
        if self._ready_bytes >= self._fusion_bytes:
            self._flush()

    def _flush(self) -> None:
        if not self._ready:
            return
        grads, self._ready, self._ready_bytes = self._ready, [], 0
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, group=_state["group"], async_op=True)
        self._pending.append((work, flat, grads))

    def synchronize(self) -> None:
        self._flush()
        for work, flat, grads in self._pending:
            work.wait()
            if self._op == Average:
                flat.div_(size())
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        self._pending = []

    def step(self, closure=None):
        if is_distributed():
            self.synchronize()
        return self._opt.step(closure)


# ------------------------------------------------------------------------------- fused path
class FusedDataParallel:
    """Couples a ``StageExecutor`` whose store lives in symmetric memory with the fused
    all-reduce+SGD kernel.  ``group`` may be the world (Horovod mode) or a stage's replica
    group (PipeDream hybrid, K4 in SURVEY.md §2.9)."""

    def __init__(self, executor, ctx, blocks: int = 32, overlap: bool = True):
        from .. import ops
        self.ex = executor
        self.ctx = ctx
        self.ops = ops
        # one peer (2-replica stages): twice the CTAs keep enough bytes in flight on the single link
        self.blocks = min(blocks * 2 if ctx.world <= 2 else blocks, 64)
        st = executor.store
        assert st.shard_optimizer and st.world == ctx.world
        self.w_sym = executor._sym_w        # SymmetricBuffer list (per version)
        self.g_sym = executor._sym_g
        self.comm_stream = torch.cuda.Stream(device=st.device) if overlap else None
        self._ev_pool: List[torch.cuda.Event] = []
        self._launched = 0
        self.log_reduce_times = os.environ.get("DDLB_LOG_REDUCE_TIMES", "0") == "1"
        self._reduce_events = []
        executor.on_bucket_ready = self._on_bucket
        # average over ranks folded into grad_scale
        st.set_grad_scale(1.0 / ctx.world)

    @staticmethod
    def make_allocator(ctx, keep: Dict[str, list]):
        """Allocator for ParamStore that places weight / gradient buffers in symmetric memory.
        ParamStore allocates the weight version slots first and the gradient buffer last."""
        def alloc(numel: int, dtype: torch.dtype) -> torch.Tensor:
            sb = ctx.alloc(numel, dtype)
            keep.setdefault("bufs", []).append(sb)
            return sb.tensor
        return alloc

    def _on_bucket(self, b: int) -> None:
        st = self.ex.store
        bucket = st.layout.buckets[b]
        mo, ml = st.master_slices()[b]
        # weight stashing (PipeDream stages with V > 1 versions): the update reads the fp32 master
        # (always the newest weights) and writes the NEXT slot of the ring, (latest + 1) % V — the
        # slot of the minibatch whose backward is running, which no longer reads this bucket's
        # weights (all of the bucket's gradients are final) — on every replica, after the kernel's
        # barrier-in has established that all replicas are past this bucket.  finish_step() then
        # advances the ring.
        out_v = st.next_version() if st.versions > 1 else st.latest
        wsym = self.w_sym[out_v]
        # NVLS from 3 ranks on; with a single peer plain P2P loads/stores are faster (profiles/allreduce_sweep_n2.json)
        use_mm = self.ctx.use_multimem and self.ctx.world > 2 and wsym.mc_ptr != 0 and self.g_sym.mc_ptr != 0
        main = torch.cuda.current_stream(st.device)
        stream = self.comm_stream or main
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            self.comm_stream.wait_event(ev)
        log = self.log_reduce_times and not torch.cuda.is_current_stream_capturing()
        if log:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        with torch.cuda.stream(stream):
            self.ops.ext().ar_sgd_twoshot(
                self.g_sym.ptrs, wsym.ptrs, self.ctx.flags.ptrs, self.g_sym.mc_ptr, wsym.mc_ptr,
                st.master[mo:mo + ml], st.mom[mo:mo + ml], st.hyper, self.ctx.epoch,
                self.ctx.status, bucket.offset, bucket.length // st.world, st.rank, self.blocks,
                use_mm)
            self.ops._count()
        if log:
            e1.record(stream)
            self._reduce_events.append((b, bucket.length, e0, e1))
        self._launched += 1

    def finish_step(self) -> None:
        """Join the communication stream back into the compute stream (end of backward)."""
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            torch.cuda.current_stream(self.ex.store.device).wait_event(ev)
        if self.ex.store.versions > 1:
            self.ex.store.advance_version()
        if self._reduce_events:
            self.flush_reduce_log()

    def flush_reduce_log(self, out=print) -> None:
        """Per-bucket all-reduce timing lines (``DDLB_LOG_REDUCE_TIMES=1``, eager mode only) —
        the information the reference gets from its patched DDP's ``log_reduce_times=True``
        (``profiler/image_classification/main.py:248-249``), parsed by
        ``utils.sparsity.extract_reduce_times``.  Times are device times relative to the first
        bucket of the step, in seconds."""
        evs, self._reduce_events = self._reduce_events, []
        if not evs:
            return
        evs[-1][3].synchronize()
        base = evs[0][2]
        for b, n, e0, e1 in evs:
            out("allreduce[%d] start=%.6f end=%.6f elements=%d" % (
                b, base.elapsed_time(e0) / 1e3, base.elapsed_time(e1) / 1e3, n))

    def broadcast_from(self, root: int = 0) -> None:
        """Make every rank start from rank ``root``'s weights (hvd.broadcast_parameters) and
        rebuild the sharded fp32 master from them."""
        st = self.ex.store
        # broadcast the fp32 MASTER (all-gather of the root's view is unnecessary: before the first
        # step every rank still holds the full-precision initial values it was constructed with, so
        # the root's full master is re-created from its init and sent as fp32 — the low bits survive)
        full = getattr(st, "_init_full", None)
        if full is None:
            full = st.w_flat[st.latest].float()
        full = full.clone()
        src = dist.get_global_rank(self.ctx.group, root) if self.ctx.group is not None else root
        dist.broadcast(full, src=src, group=self.ctx.group)
        st.load_full_master(full)
        st._init_full = None
        for k in sorted(self.ex.buffers):
            dist.broadcast(self.ex.buffers[k], src=src, group=self.ctx.group)
        torch.cuda.synchronize(st.device)
        dist.barrier(self.ctx.group)
