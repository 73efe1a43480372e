"""PipeDream stage runtime: asynchronous 1F1B pipelining with weight stashing and intra-stage
replication, one process per GPU.

Reference: ``StageRuntime`` + ``CommunicationHandler`` + ``OptimizerWithWeightStashing``
(``pipedream-fork/runtime/{runtime,communication,optimizer}.py``) driven by the loop in
``runtime/image_classification/main_with_runtime.py:406-500``.  What is kept: the stage /
rank configuration files, the warm-up formula, the 1F1B order, the replica routing, per-stage
LR scaling, the iteration reconciliation, macrobatching, ``--no_input_pipelining``, per-stage
checkpoints.  What is different by design:

* each stage is a compiled plan on the static engine — no autograd tape, explicit backward;
* weight stashing = V slots of the flat bf16 weight buffer + an index; "load_old_params" is
  "run backward against slot v", "step" writes slot (latest+1) % V from the fp32 master in the
  fused SGD kernel — zero copies (reference: 2 ``load_state_dict`` + 1 full copy per step);
* stage-boundary tensors have static shapes (meta-device dry run), receives are pre-posted
  into rings on a communication stream, no helper threads, no per-message allocation;
* replica gradient reduction uses the fused all-reduce+SGD kernel over the stage's replica
  group on GPUs (Gloo/NCCL ``all_reduce`` + fused SGD otherwise);
* evaluation needs no ack tokens: back-pressure comes from the bounded receive rings;
* (opt-in, ``DDLB_PIPE_GRAPHS=1``) the forward and the backward of a stage replay as CUDA graphs,
  one pair per weight-version slot: an eager pipeline is bound by launch cost (~2.1 ms per
  minibatch and stage for ResNet-50), not by the GPU.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ...dist.transport import DistP2P, PairGroups, SymmP2P, TensorSpec
from ...engine import FwdContext, StageExecutor
from ...models.graph import INPUT, LayerGraphModel
from .config import PipelineConfig, RankTopology
from .schedule import F, Routing, one_f_one_b


def infer_boundary_specs(model: LayerGraphModel, cuts: Sequence[int], batch: int,
                         dtype: torch.dtype, channels_last: bool) -> List[List[TensorSpec]]:
    """Shapes of the tensors crossing each cut (``cuts[k]`` = index of the first node of stage
    k+1), from a meta-device dry run (reference: CPU dry run with ``torch.zeros``,
    ``main_with_runtime.py:168-192``)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    with torch.no_grad():
        x = torch.empty((batch,) + tuple(model.input_shape), device="meta")
        live = {INPUT: x}
        shapes[INPUT] = tuple(x.shape)
        for i, n in enumerate(model.nodes):
            mod = model.node_module(i)
            ins = [live[k] for k in n.inputs]
            # the real modules stay where they are; only meta ALIASES of their state are used
            state = {k: torch.empty_like(v, device="meta")
                     for k, v in list(mod.named_parameters()) + list(mod.named_buffers())}
            was_training = mod.training
            mod.eval()
            out = torch.func.functional_call(mod, state, tuple(ins))
            mod.train(was_training)
            live[n.name] = out
            shapes[n.name] = tuple(out.shape)
    specs = []
    for c in cuts:
        names = model.live_after(c - 1)
        specs.append([TensorSpec(nm, shapes[nm], dtype, channels_last and len(shapes[nm]) == 4)
                      for nm in names])
    return specs


def reorder_by_stage(model: LayerGraphModel, stage_of_node: Sequence[int]) -> Tuple[LayerGraphModel, List[int]]:
    """Stable-sort the nodes by stage id so every stage is a contiguous slice (the partitioner
    assigns stages along a topological order of *its* graph, which need not be the definition
    order for branchy blocks).  Returns the re-ordered model view and the cut indices."""
    order = sorted(range(len(model.nodes)), key=lambda i: (stage_of_node[i], i))
    nodes = [model.nodes[i] for i in order]
    m = LayerGraphModel(nodes, name=model.arch, input_shape=model.input_shape,
                        num_classes=model.num_classes)
    stages = [stage_of_node[i] for i in order]
    cuts = [k for k in range(1, len(stages)) if stages[k] != stages[k - 1]]
    return m, cuts


class _EagerGraph:
    """Stand-in for ``torch.cuda.CUDAGraph`` where graphs do not exist (CPU tests of the slot
    logic): "capture" runs the body once, "replay" runs it again."""

    def __init__(self, body):
        self.body = body

    def replay(self) -> None:
        self.body()


class _GraphSlot:
    """Static tensors and the two captured graphs (forward, backward) of ONE weight-version
    slot.  A minibatch whose forward ran on weight slot v keeps the slot ``busy`` until its
    backward has replayed: the saved activations live in the graph's private memory pool, so
    a second in-flight minibatch on the same weight slot (only during the warm-up forwards,
    which all see version 0) takes the eager path instead."""
    __slots__ = ("version", "x", "target", "outs", "stats", "dl", "ctx", "g", "gin", "fwd", "bwd",
                 "pool", "busy", "sent")

    def __init__(self, version: int):
        self.version = version
        self.x = self.target = self.outs = self.stats = self.dl = self.ctx = None
        self.g = self.gin = self.fwd = self.bwd = self.pool = None
        self.busy = False
        self.sent = []          # events of the transfers that still read this slot's tensors


class StageRuntime:
    def __init__(self, model: LayerGraphModel, cuts: Sequence[int], cfg: PipelineConfig,
                 batch_size: int, device="cpu", dtype: torch.dtype = torch.float32,
                 lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 1e-4,
                 rank: Optional[int] = None, num_versions: Optional[int] = None,
                 no_input_pipelining: bool = False, macrobatch: bool = False,
                 tcgen05: bool = False, fused_replica_allreduce: bool = True,
                 groups: Optional[PairGroups] = None, verbose: bool = False,
                 graphs: Optional[bool] = None, loss_scale: float = 1.0):
        self.cfg = cfg
        self.rank = dist.get_rank() if rank is None else rank
        self.topo = RankTopology.of(cfg, self.rank)
        self.device = torch.device(device)
        self.dtype = dtype
        self.batch_size = batch_size
        t = self.topo
        assert cfg.num_stages == len(cuts) + 1, "number of cuts must be num_stages - 1"
        bounds = [0] + list(cuts) + [len(model.nodes)]
        start, stop = bounds[t.stage], bounds[t.stage + 1]
        self.warmup = 0 if no_input_pipelining else t.warmup
        versions = num_versions if num_versions is not None else self.warmup + 1
        self.macrobatch = macrobatch
        self.update_interval = 1
        if macrobatch:
            # PipeDream-2BW-like: <= 2 versions, one step per `versions` minibatches with the
            # gradients averaged (reference optimizer.py:36-52,118-149)
            self.update_interval = max(versions, 1)
            versions = min(2, versions)
        on_gpu = self.device.type == "cuda"
        self.channels_last = on_gpu
        # replica group of this stage (gradient averaging among the stage's replicas)
        self.replica_group = None
        self._replica_groups = {}
        for s in sorted(cfg.stage_to_rank_map):
            ranks = cfg.stage_to_rank_map[s]
            if len(ranks) > 1:
                g = dist.new_group(ranks)          # collective: same order on every rank
                self._replica_groups[s] = g
                if s == t.stage:
                    self.replica_group = g
        self.fdp = None
        kw = dict(device=self.device, dtype=dtype, versions=versions,
                  lr=lr * cfg.lr_scale(t.stage), momentum=momentum, weight_decay=weight_decay,
                  tcgen05=tcgen05, first_stage=t.is_first)
        if self.replica_group is not None and on_gpu and fused_replica_allreduce:
            from ...dist.symm import SymmContext
            from ..dp import FusedDataParallel
            self.symm = SymmContext(self.device, self.replica_group)
            keep: Dict[str, list] = {}
            alloc = FusedDataParallel.make_allocator(self.symm, keep)
            self.ex = StageExecutor.from_model(model, start, stop, world=t.num_ranks_in_stage,
                                               rank=t.rank_in_stage, shard_optimizer=True,
                                               allocator=alloc, **kw)
            self.ex._sym_w, self.ex._sym_g = keep["bufs"][:-1], keep["bufs"][-1]
            self.fdp = FusedDataParallel(self.ex, self.symm)
            self.fdp.broadcast_from(0)
        else:
            self.ex = StageExecutor.from_model(model, start, stop, **kw)
            if self.replica_group is not None:
                full = self.ex.store.master.clone()
                dist.broadcast(full, src=t.stage_ranks[0], group=self.replica_group)
                self.ex.store.load_full_master(full)
                self.ex.store.set_grad_scale(1.0 / t.num_ranks_in_stage)
        if self.update_interval != 1:
            # macrobatch: every stage averages the update_interval minibatch gradients it accumulates
            # (its OWN interval: warm-up depth + 1), like the reference optimizer's
            # ``p.grad /= update_interval`` (pipedream-fork/runtime/optimizer.py:140-142); the loss
            # is not pre-scaled (the last stage has interval 1); the 1/replicas of the replica average
            # rides in the same scale on both the fused and the NCCL reduction path.
            reps = t.num_ranks_in_stage if self.replica_group is not None else 1
            self.ex.store.set_grad_scale(1.0 / (reps * self.update_interval))
        # static loss scaling (reference: --fp16 --loss_scale S, runtime.py:225-230,269-277 and
        # optimizer.py:137-155): the loss gradient is multiplied by S on the last stage, so every
        # gradient that flows back through the pipeline carries S, and every stage's optimizer divides
        # it out again (folded into the gradient scale of the fused SGD / all-reduce kernels)
        self.loss_scale = float(loss_scale)
        if self.loss_scale != 1.0:
            self.ex.store.set_grad_scale(float(self.ex.store.hyper[3]) / self.loss_scale)
        if verbose and self.replica_group is not None:
            print("Replicating stage: ranks=%d, module_size=%d" % (
                t.num_ranks_in_stage, self.ex.num_params() * 4), flush=True)
        # boundary tensor specs (+ the labels, forwarded hop by hop as int64 [B])
        specs = infer_boundary_specs(model, cuts, batch_size, dtype, self.channels_last)
        lab = TensorSpec("target", (batch_size,), torch.int64)
        self.in_specs = (specs[t.stage - 1] + [lab]) if not t.is_first else []
        self.out_specs = (specs[t.stage] + [lab]) if not t.is_last else []
        self.grad_in_specs = specs[t.stage] if not t.is_last else []     # grads of my outputs
        self.grad_out_specs = specs[t.stage - 1] if not t.is_first else []  # grads I send back
        if groups is None and dist.is_initialized() and dist.get_world_size() > 1:
            pairs = []
            for s in range(cfg.num_stages - 1):
                for a in cfg.stage_to_rank_map[s]:
                    for b in cfg.stage_to_rank_map[s + 1]:
                        pairs.append((a, b))
            groups = PairGroups(pairs)
        self.ring = self.warmup + 3
        if os.environ.get("DDLB_P2P", "nccl") == "symm" and on_gpu and groups is not None:
            # peer-store transport: every rank derives the same slot size / ring depth
            def msg_bytes(sp):
                return sum((math.prod(s.shape) * torch.empty((), dtype=s.dtype).element_size() + 255) // 256 * 256
                           for s in sp + [lab])
            slot = max(msg_bytes(sp) for sp in specs) if specs else 256
            depth = 3 + max(RankTopology.of(cfg, r).warmup for r in range(cfg.num_ranks))
            self.p2p = SymmP2P.from_groups(self.device, groups, slot, depth)
        else:
            self.p2p = DistP2P(self.device, groups)
        self.routing = Routing(self.topo)
        self.fwd_id = self.bwd_id = 0
        self._ctxs: List[Tuple[FwdContext, Optional[torch.Tensor], Optional[int]]] = []
        self._next_fwd = None
        self._next_bwd = None
        self.loss_stats: Optional[torch.Tensor] = None
        from .stats import RuntimeStats
        self.forward_stats, self.backward_stats = RuntimeStats(True), RuntimeStats(False)
        self.training = True
        self.verbose = verbose
        self._accum = 0
        # graph replay of the stage's forward / backward (see _GraphSlot)
        self.graphs = (os.environ.get("DDLB_PIPE_GRAPHS", "0") == "1") if graphs is None else bool(graphs)
        if self.graphs and (self.fdp is not None or self.update_interval != 1):
            if verbose:
                print("stage graphs off: fused replica all-reduce / macrobatching keep host-side "
                      "state per minibatch", flush=True)
            self.graphs = False
        self._slots: Dict[int, _GraphSlot] = {}
        self._seen_fwd = 0
        # lazy initialisation (cuDNN autotune, workspaces, handles) must be over before a capture
        self._graph_after = int(os.environ.get("DDLB_PIPE_GRAPH_AFTER", 2 * versions + 2))
        self.graph_replays = 0

    # ------------------------------------------------------------------ graphed stage passes
    def _capture(self, slot: _GraphSlot, body):
        """Run ``body`` under stream capture into the slot's private pool (GPU) or once, eagerly
        (CPU stand-in); returns an object with ``replay()``, or None if the capture failed (the
        caller then runs the pass eagerly and graphs are switched off for the rest of the run).
        The raw capture_begin/end API is used instead of ``torch.cuda.graph``: its device-wide
        synchronize would wait for the pre-posted receives of the NEXT messages."""
        if self.device.type != "cuda":
            body()
            return _EagerGraph(body)
        if slot.pool is None:
            slot.pool = torch.cuda.graph_pool_handle()
        if getattr(self, "_capture_stream", None) is None:
            self._capture_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        side = self._capture_stream
        side.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        ok = True
        with torch.cuda.stream(side):
            # thread_local: the NCCL watchdog and the loader threads keep making CUDA calls
            g.capture_begin(pool=slot.pool, capture_error_mode="thread_local")
            try:
                body()
            except Exception as e:  # noqa: BLE001
                ok = False
                print("[pipeline] stage graph capture failed (%s); continuing eagerly"
                      % str(e).splitlines()[0][:160], flush=True)
            try:
                g.capture_end()
            except Exception as e:  # noqa: BLE001
                if ok:
                    print("[pipeline] stage graph capture failed (%s); continuing eagerly"
                          % str(e).splitlines()[0][:160], flush=True)
                ok = False
        cur.wait_stream(side)
        if not ok:
            self.graphs = False
            return None
        return g

    def _slot_for(self, version: int) -> Optional[_GraphSlot]:
        if not (self.graphs and self.training) or self._seen_fwd < self._graph_after:
            return None
        slot = self._slots.get(version)
        if slot is None:
            slot = self._slots[version] = _GraphSlot(version)
        return None if slot.busy else slot

    def _wait_sent(self, slot: _GraphSlot) -> None:
        """The transfers that read the slot's static outputs must be over before a replay
        overwrites them (always true w+1 minibatches later; this makes it explicit)."""
        if slot.sent and self.device.type == "cuda":
            cur = torch.cuda.current_stream(self.device)
            for ev in slot.sent:
                cur.wait_event(ev)
        slot.sent = []

    def _graphed_forward(self, slot: _GraphSlot, xs: List[torch.Tensor], target: torch.Tensor):
        ex, last = self.ex, self.topo.is_last
        if slot.x is None:
            slot.x = [torch.empty_like(x) for x in xs]
            slot.target = torch.empty_like(target)
        for d, s_ in zip(slot.x, xs):
            d.copy_(s_)
        slot.target.copy_(target)
        self._wait_sent(slot)
        if slot.fwd is None:
            def body():
                slot.outs, slot.ctx = ex.forward(slot.x, training=True, version=slot.version)
                if last:
                    slot.stats, slot.dl = ex.loss(slot.outs[0], slot.target, want_grad=True,
                                                  grad_scale=self.loss_scale)
            slot.fwd = self._capture(slot, body)
            if slot.fwd is None:
                return None
            if self.device.type != "cuda":          # the stand-in's "capture" already computed
                slot.busy = True
                return slot.outs, (slot.stats.clone() if last else None)
        slot.fwd.replay()
        self.graph_replays += 1
        slot.busy = True
        return slot.outs, (slot.stats.clone() if last else None)

    def _graphed_backward(self, slot: _GraphSlot, grads: Optional[List[torch.Tensor]]):
        ex = self.ex
        if grads is not None:                        # last stage: slot.dl is already in place
            if slot.g is None:
                slot.g = [torch.empty_like(g) for g in grads]
            for d, s_ in zip(slot.g, grads):
                d.copy_(s_)
        self._wait_sent(slot)
        if slot.bwd is None:
            def body():
                # backward() empties the context as it goes: hand it a copy of the list so a
                # failed capture leaves slot.ctx usable for the eager fallback below
                view = FwdContext(0, slot.ctx.version)
                view.saved = [dict(d) for d in slot.ctx.saved]
                view.outputs, view.need_input_grad = slot.ctx.outputs, slot.ctx.need_input_grad
                slot.gin = ex.backward(view, slot.g if grads is not None else [slot.dl],
                                       accumulate=False, final=True)
            slot.bwd = self._capture(slot, body)
            if slot.bwd is None:
                # the forward replay filled slot.ctx's tensors, so the plain backward works on them
                slot.busy = False
                return ex.backward(slot.ctx, slot.g if grads is not None else [slot.dl],
                                   accumulate=False, final=True)
            if self.device.type != "cuda":
                slot.busy = False
                return slot.gin
        slot.bwd.replay()
        self.graph_replays += 1
        slot.busy = False
        return slot.gin

    # ------------------------------------------------------------------ epoch control
    def train(self, num_iterations: int) -> None:
        self.training = True
        self.n_iter = num_iterations
        self.fwd_id = self.bwd_id = 0
        self._ctxs = []
        self.routing = Routing(self.topo)
        self._next_fwd = self._next_bwd = None
        self._posted_fwd = self._posted_bwd = 0
        self._accum = 0
        for slot in self._slots.values():
            slot.busy = False
        self._prepost_fwd()

    def eval(self, num_iterations: int) -> None:
        self.train(num_iterations)
        self.training = False

    def _prepost_fwd(self) -> None:
        """Post the receive of the NEXT forward input so its transfer overlaps compute."""
        if self.topo.is_first or self._next_fwd is not None or self._posted_fwd >= self.n_iter:
            return
        src = self.routing.next_fwd_recv_peer()
        self._next_fwd = (src, self.p2p.post_recv(src, self.in_specs, "fwd", ring=self.ring))
        self._posted_fwd += 1

    def _prepost_bwd(self) -> None:
        if self.topo.is_last or self._next_bwd is not None or self._posted_bwd >= self.n_iter:
            return
        src = self.routing.bwd_recv_peer(self._posted_bwd)
        self._next_bwd = (src, self.p2p.post_recv(src, self.grad_in_specs, "bwd", ring=3))
        self._posted_bwd += 1

    # ------------------------------------------------------------------ forward
    def run_forward(self, inputs: Optional[torch.Tensor] = None,
                    target: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """One forward pass.  First stage: ``inputs`` / ``target`` come from the loader; other
        stages receive them.  Last stage returns stats[3] (loss, #top1, #top5)."""
        t = self.topo
        ex = self.ex
        if self.training:
            from ... import ops as _ops
            _ops.rng_tick(self.device)         # dropout masks differ from minibatch to minibatch
        if t.is_first:
            xs = [ex.prepare_input(inputs)]
            target = target.to(self.device)
        else:
            if self._next_fwd is None:
                self._prepost_fwd()
            _, pending = self._next_fwd
            self._next_fwd = None
            bufs = pending.wait()
            xs, target = list(bufs[:-1]), bufs[-1]
            self.forward_stats.received(bufs)
            self._prepost_fwd()               # overlap the next transfer with this compute
        version = ex.store.latest
        slot = self._slot_for(version)
        tok = self.forward_stats.begin_compute(self.device)
        stats = None
        dl = None
        res = self._graphed_forward(slot, xs, target) if slot is not None else None
        if res is not None:
            outs, stats = res
            ctx = slot
        else:
            slot = None
            outs, ctx = ex.forward(xs, training=self.training, version=version)
            if t.is_last:
                stats, dl = ex.loss(outs[0], target, want_grad=self.training,
                                    grad_scale=self.loss_scale)
        self.forward_stats.end_compute(tok)
        if t.is_last:
            self.loss_stats = stats
        else:
            ev = self.p2p.send(self.routing.fwd_send_peer(self.fwd_id), list(outs) + [target], "fwd")
            if slot is not None and ev is not None:
                slot.sent.append(ev)
            self.forward_stats.sent(list(outs) + [target])
        if self.training:
            self._ctxs.append((ctx, dl, None))
            self._seen_fwd += 1
        self.fwd_id += 1
        return stats

    # ------------------------------------------------------------------ backward
    def run_backward(self) -> None:
        t = self.topo
        ex = self.ex
        ctx, dl, _ = self._ctxs.pop(0)
        if t.is_last:
            grads = [dl]
        else:
            if self._next_bwd is None:
                self._prepost_bwd()
            _, pending = self._next_bwd
            self._next_bwd = None
            grads = list(pending.wait())
            self.backward_stats.received(grads)
            self._prepost_bwd()
        accumulate = self._accum > 0
        final = (self._accum + 1) == self.update_interval
        tok = self.backward_stats.begin_compute(self.device)
        slot = ctx if isinstance(ctx, _GraphSlot) else None
        if slot is not None:
            gin = self._graphed_backward(slot, None if t.is_last else grads)
        else:
            gin = ex.backward(ctx, grads, accumulate=accumulate, final=final)
        self.backward_stats.end_compute(tok)
        if not t.is_first:
            ev = self.p2p.send(self.routing.next_bwd_send_peer(), [g for g in gin], "bwd")
            if slot is not None and ev is not None:
                slot.sent.append(ev)
            self.backward_stats.sent(gin)
        self._accum += 1
        if final:
            self._optimizer_step()
            self._accum = 0
        self.bwd_id += 1

    def _optimizer_step(self) -> None:
        ex = self.ex
        st = ex.store
        if self.fdp is not None:
            # fused all-reduce+SGD already ran per bucket (into the next weight slot when the stage
            # stashes versions); this joins the communication stream and advances the ring
            self.fdp.finish_step()
            return
        if self.replica_group is not None:
            dist.all_reduce(st.g_flat, group=self.replica_group)   # averaged via grad_scale
        nxt = st.next_version()
        st.sgd_step(out_version=nxt)
        st.advance_version()

    # ------------------------------------------------------------------ full 1F1B epoch
    def run_epoch(self, loader: Optional[Iterator] = None, num_iterations: Optional[int] = None,
                  on_forward=None) -> None:
        """``warmup`` forwards, 1F1B steady state, drain
        (``main_with_runtime.py:432-494``)."""
        n = num_iterations
        self.train(n)
        it = iter(loader) if loader is not None else None
        for op, i in one_f_one_b(n, self.warmup):
            if op == F:
                if self.topo.is_first:
                    x, y = next(it)
                    stats = self.run_forward(x, y)
                else:
                    stats = self.run_forward()
                if on_forward is not None:
                    on_forward(i, stats)
            else:
                self.run_backward()
        self.p2p.flush()

    def run_eval(self, loader: Optional[Iterator], num_iterations: int, on_forward=None) -> None:
        self.eval(num_iterations)
        it = iter(loader) if loader is not None else None
        for i in range(num_iterations):
            if self.topo.is_first:
                x, y = next(it)
                stats = self.run_forward(x, y)
            else:
                stats = self.run_forward()
            if on_forward is not None:
                on_forward(i, stats)
        self.p2p.flush()

    # ------------------------------------------------------------------ checkpoints
    def state_dict(self) -> Dict[str, object]:
        """Per-stage checkpoint payload (reference ``checkpoint.<stage>.pth.tar``,
        ``main_with_runtime.py:392-403,580-584``).  Unlike the reference, the optimizer's
        fp32 master + momentum are saved, so a resumed run continues bit-identically.  COLLECTIVE
        over the stage's replicas when the optimizer state is sharded (fused replica reduction):
        the full master / momentum are gathered so that the file written by replica 0 is complete."""
        st = self.ex.store
        master, mom = st.gather_full_state(self.replica_group)
        return {"stage": self.topo.stage, "state_dict": self.ex.state_dict(),
                "master": master.detach().cpu().clone(), "momentum": mom.detach().cpu().clone(),
                "hyper": st.hyper.detach().cpu().clone()}

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        self.ex.load_state_dict(sd["state_dict"])
        st = self.ex.store
        if "master" in sd and sd["master"].numel() == st.layout.total:
            st.load_full_state(sd["master"], sd["momentum"])     # re-sliced by rank when sharded

    def set_lr(self, lr: float) -> None:
        self.ex.store.set_lr(lr * self.cfg.lr_scale(self.topo.stage))
