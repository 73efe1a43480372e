"""Training-step driver for the single-GPU (`pytorch`) and data-parallel (`horovod`) modes.

The reference step is five eager lines — H2D copy, forward, loss, ``zero_grad``+backward,
``optimizer.step`` — issued as ~600 separate library kernels every iteration
(``benchmark/imagenet/imagenet_pytorch.py:149-157``).  At the benchmark's per-GPU batch of 32
a B200 finishes each of those kernels faster than the host can launch the next one, so here
the WHOLE step (forward, fused loss, backward, bucketed fused all-reduce+SGD on the
communication stream) is captured once into a CUDA graph and replayed; input staging
(pinned host -> device copy of the raw uint8 batch on a copy stream, double-buffered) and the
loss read-back overlap the previous / next replay.
"""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..engine import StageExecutor
from ..models import build_model
from ..models.graph import LayerGraphModel


class Trainer:
    def __init__(self, arch: str = "resnet50", dataset: str = "imagenet", batch_size: int = 32,
                 device: Optional[str] = None, dtype: Optional[torch.dtype] = None,
                 lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 1e-4,
                 cuda_graph: bool = True, distributed: bool = False, allreduce: str = "fused",
                 tcgen05: bool = False, seed: int = 1, model: Optional[LayerGraphModel] = None,
                 bucket_elems: int = 8 << 20, comm_blocks: int = 32, group=None):
        torch.manual_seed(seed)
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        on_gpu = self.device.type == "cuda"
        if on_gpu:
            torch.cuda.manual_seed(seed)
            torch.backends.cudnn.benchmark = True
        self.dtype = dtype or (torch.bfloat16 if on_gpu else torch.float32)
        self.batch_size = batch_size
        self.model = model if model is not None else build_model(arch, dataset)
        self.arch, self.dataset = arch, dataset
        self.distributed = distributed and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.allreduce = allreduce if self.distributed else "none"
        self.cuda_graph = cuda_graph and on_gpu
        self.fdp = None
        self._sym: Dict[str, list] = {}
        kw = dict(device=self.device, dtype=self.dtype, lr=lr, momentum=momentum,
                  weight_decay=weight_decay, bucket_elems=bucket_elems, tcgen05=tcgen05)
        if self.distributed and self.allreduce == "fused" and on_gpu:
            from ..dist.symm import SymmContext
            from ..parallel.dp import FusedDataParallel
            self.symm = SymmContext(self.device, group)
            alloc = FusedDataParallel.make_allocator(self.symm, self._sym)
            self.ex = StageExecutor.from_model(self.model, world=self.world, rank=self.rank,
                                               shard_optimizer=True, allocator=alloc, **kw)
            bufs = self._sym["bufs"]
            self.ex._sym_w, self.ex._sym_g = bufs[:-1], bufs[-1]
            self.fdp = FusedDataParallel(self.ex, self.symm, blocks=comm_blocks)
            self.fdp.broadcast_from(0)
        else:
            self.ex = StageExecutor.from_model(self.model, **kw)
            if self.distributed:
                self.ex.store.set_grad_scale(1.0 / self.world)
                full = self.ex.store.master
                dist.broadcast(full, src=0, group=group)
                self.ex.store.load_full_master(full.clone())
        self.group = group
        # opt-in: per-bucket SGD on a side stream while backward is still running (single GPU).
        # The optimizer is 85 us (1.8 %) at the end of the ResNet-50 step; overlapped, only the
        # last bucket's update stays on the critical path.  Off until measured on hardware.
        self._sgd_overlap = (os.environ.get("DDLB_SGD_OVERLAP", "0") == "1" and self.fdp is None
                             and not self.distributed)
        self._sgd_stream = torch.cuda.Stream(device=self.device) if (self._sgd_overlap and on_gpu) else None
        self._sgd_buckets_done = 0
        if self._sgd_overlap:
            self.ex.on_bucket_ready = self._sgd_bucket
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_x: Optional[torch.Tensor] = None
        self.static_t: Optional[torch.Tensor] = None
        self.static_stats: Optional[torch.Tensor] = None
        self.launches_per_step = 0
        self.steps_done = 0

    # ------------------------------------------------------------------ one eager step
    def _step_eager(self, x: torch.Tensor, target: torch.Tensor, update: bool = True) -> torch.Tensor:
        ex = self.ex
        ops.rng_tick(self.device)          # fresh dropout masks every step (also inside the graph)
        outs, ctx = ex.forward([x], training=True)
        stats, dlogits = ex.loss(outs[0], target)
        ex.backward(ctx, [dlogits])
        if self.fdp is not None:
            self.fdp.finish_step()
        elif update:
            self._update_baseline()
        return stats

    def _sgd_bucket(self, b: int) -> None:
        """Bucket ``b`` has its final gradients (and its weights are no longer read by this
        backward): update it now, off the critical path."""
        st = self.ex.store
        if self._sgd_stream is None:
            st.sgd_step(bucket=b)
        else:
            main = torch.cuda.current_stream(self.device)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self._sgd_stream):
                self._sgd_stream.wait_event(ev)
                st.sgd_step(bucket=b)
        self._sgd_buckets_done += 1

    def _update_baseline(self) -> None:
        """Non-fused update: (NCCL all-reduce of the flat gradient buffer, then) fused SGD."""
        if self.distributed:
            dist.all_reduce(self.ex.store.g_flat, group=self.group)
        if self._sgd_overlap:
            # every bucket was updated as it became ready; just join the side stream
            assert self._sgd_buckets_done == len(self.ex.store.layout.buckets), "a bucket was not updated"
            self._sgd_buckets_done = 0
            if self._sgd_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self._sgd_stream)
            return
        self.ex.store.sgd_step()

    @property
    def _nccl_outside_graph(self) -> bool:
        # the NCCL baseline keeps its collective OUT of the captured graph (forward+backward
        # are replayed, the all-reduce and the optimizer kernel are launched eagerly after it)
        return self.distributed and self.fdp is None

    # ------------------------------------------------------------------ public step
    def step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One training step on device tensors.  Returns stats[3] (mean loss, #top1, #top5) —
        a device tensor; reading it is the caller's (asynchronous) business."""
        if not self.cuda_graph:
            self.steps_done += 1
            return self._step_eager(self.ex.prepare_input(x), target)
        if self.graph is None:
            self._capture(x, target)
            if not self.cuda_graph:          # capture failed: eager from now on
                self.steps_done += 1
                return self._step_eager(self.ex.prepare_input(x), target)
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if target.data_ptr() != self.static_t.data_ptr():
            self.static_t.copy_(target, non_blocking=True)
        self._replay()
        self.steps_done += 1
        return self.static_stats

    def step_u8(self, x_u8: torch.Tensor, target: torch.Tensor, mean, std) -> torch.Tensor:
        """Step from a raw uint8 [B, H, W, C] device batch: the uint8 -> normalised
        channels-last conversion is one fused kernel writing the graph's static input."""
        if self.cuda_graph:
            if self.graph is None:
                self._capture(ops.normalize_u8(x_u8, mean, std, dtype=self.dtype), target)
                if not self.cuda_graph:
                    return self.step(ops.normalize_u8(x_u8, mean, std, dtype=self.dtype), target)
            ops.normalize_u8(x_u8, mean, std, dtype=self.dtype, out=self.static_x)
            self.static_t.copy_(target, non_blocking=True)
            self._replay()
            self.steps_done += 1
            return self.static_stats
        return self.step(ops.normalize_u8(x_u8, mean, std, dtype=self.dtype), target)

    def _replay(self) -> None:
        self.graph.replay()
        if self._nccl_outside_graph:
            self._update_baseline()

    def _capture(self, x: torch.Tensor, target: torch.Tensor) -> None:
        ex = self.ex
        self.static_x = ex.prepare_input(x).clone()
        self.static_t = target.to(self.device).clone()
        # warm-up on a side stream: workspace allocation, lazy inits.  The three warm-up steps are
        # real steps (optimizer, momentum, BatchNorm running statistics, fused all-reduce): snapshot
        # the training state and put it back afterwards, so the first counted step starts from the
        # initial weights and the run is step-for-step identical to eager mode.
        st = ex.store
        snap = [t.clone() for t in [st.master, st.mom] + list(st.w_flat) + list(ex.buffers.values())]
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._step_eager(self.static_x, self.static_t)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.distributed:
            dist.barrier(self.group)          # nobody restores while a peer still reads its weights
        for dst, src in zip([st.master, st.mom] + list(st.w_flat) + list(ex.buffers.values()), snap):
            dst.copy_(src)
        del snap
        torch.cuda.synchronize(self.device)
        if self.distributed:
            dist.barrier(self.group)
        # Capture.  If it fails (all ranks decide together), retry once without programmatic
        # dependent launch, and as a last resort stay in eager mode — a benchmark run must
        # degrade, not die.
        ok = self._try_capture()
        if not ok and ops.ext().set_pdl(False):
            print("[trainer] graph capture failed; retrying without programmatic dependent launch",
                  flush=True)
            ok = self._try_capture()
        if not ok:
            print("[trainer] graph capture failed; continuing in eager mode", flush=True)
            self.cuda_graph = False
            self.graph = None
        torch.cuda.synchronize(self.device)
        if self.distributed:
            dist.barrier(self.group)

    def _try_capture(self) -> bool:
        before = ops.launches
        g = torch.cuda.CUDAGraph()
        failed = 0
        try:
            # DDLB_MAIN_PRIO=1: capture on a high-priority stream (the main-stream kernel nodes inherit
            # it, the weight-gradient side stream keeps the default).  Measured SLOWER (5.87 vs 5.31
            # ms/step, profiles/r2/README.md): the starved weight gradients end up on the critical
            # path at the optimizer.  Off.
            cap = None
            if os.environ.get("DDLB_MAIN_PRIO", "0") == "1":
                if getattr(self, "_cap_stream", None) is None:
                    self._cap_stream = torch.cuda.Stream(device=self.device, priority=-1)
                cap = self._cap_stream
            with torch.cuda.graph(g, stream=cap):
                self.static_stats = self._step_eager(self.static_x, self.static_t,
                                                     update=not self._nccl_outside_graph)
        except Exception as e:  # noqa: BLE001
            print("[trainer] capture error: %s" % str(e).splitlines()[0][:200], flush=True)
            failed = 1
            try:
                torch.cuda.synchronize(self.device)
            except Exception:  # noqa: BLE001
                pass
        if self.distributed:
            flag = torch.tensor([failed], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            failed = int(flag.item())
        if failed:
            self._sgd_buckets_done = 0
            return False
        self.launches_per_step = ops.launches - before
        self.graph = g
        return True

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def eval_step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        outs, _ = self.ex.forward([self.ex.prepare_input(x)], training=False)
        stats, _ = self.ex.loss(outs[0], target.to(self.device), want_grad=False)
        return stats

    def set_lr(self, lr: float) -> None:
        self.ex.store.set_lr(lr)

    def check(self) -> None:
        if self.fdp is not None:
            self.symm.check()
