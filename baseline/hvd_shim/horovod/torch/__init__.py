"""Stand-in for ``horovod.torch`` used ONLY to run the unmodified reference script
``benchmark/imagenet/imagenet_horovod.py`` on a box where Horovod cannot be installed (no
network, no MPI).  It implements the handful of calls that script makes with plain
``torch.distributed`` (NCCL): what Horovod itself is built to do (``HOROVOD_GPU_ALLREDUCE=NCCL``,
reference ``run/run/run_template.sh:172-173``).  It deliberately imports nothing from
ddlbench_b200: the comparator must not run any of our models, kernels or engine.

Semantics kept: per-parameter hooks enqueue gradients as they are produced, a fusion buffer
(64 MB) is all-reduced asynchronously when full, ``step()`` waits for all handles, divides by
the world size (``op=Average``) and runs the wrapped optimizer.
"""
import os

import torch
import torch.distributed as dist

Average = "average"
Sum = "sum"
_FUSION_BYTES = 64 << 20


def init():
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


def rank():
    return dist.get_rank()


def size():
    return dist.get_world_size()


def local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def allreduce(tensor, name=None, op=Average):
    t = tensor.detach().clone().cuda()
    dist.all_reduce(t)
    if op == Average:
        t /= size()
    return t.to(tensor.device)


def broadcast_parameters(params, root_rank=0):
    items = params.items() if isinstance(params, dict) else params
    for _, p in sorted(items, key=lambda kv: kv[0]):
        if torch.is_tensor(p):
            dist.broadcast(p.data if p.is_cuda else p.data.cuda(), src=root_rank)


def broadcast_optimizer_state(optimizer, root_rank=0):
    opt = getattr(optimizer, "_opt", optimizer)
    for g in opt.param_groups:
        for p in g["params"]:
            for k in sorted(opt.state.get(p, {})):
                v = opt.state[p][k]
                if torch.is_tensor(v) and v.is_cuda:
                    dist.broadcast(v, src=root_rank)


class _DistributedOptimizer:
    def __init__(self, optimizer, named_parameters=None, backward_passes_per_step=1, op=Average):
        self._opt = optimizer
        self._op = op
        self._bpps = backward_passes_per_step
        self._ready, self._ready_bytes, self._pending, self._count = [], 0, [], {}
        for g in optimizer.param_groups:
            for p in g["params"]:
                if p.requires_grad:
                    p.register_post_accumulate_grad_hook(self._hook)

    def __getattr__(self, name):
        return getattr(self._opt, name)

    @property
    def param_groups(self):
        return self._opt.param_groups

    def zero_grad(self, *a, **k):
        return self._opt.zero_grad(*a, **k)

    def _hook(self, p):
        c = self._count.get(id(p), 0) + 1
        self._count[id(p)] = c
        if c % self._bpps:
            return
        self._ready.append(p.grad)
        self._ready_bytes += p.grad.numel() * p.grad.element_size()
        if self._ready_bytes >= _FUSION_BYTES:
            self._flush()

    def _flush(self):
        if not self._ready:
            return
        grads, self._ready, self._ready_bytes = self._ready, [], 0
        flat = torch.cat([g.reshape(-1) for g in grads])
        self._pending.append((dist.all_reduce(flat, async_op=True), flat, grads))

    def synchronize(self):
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
        self.synchronize()
        return self._opt.step(closure)


def DistributedOptimizer(optimizer, named_parameters=None, backward_passes_per_step=1, op=Average,
                         **kwargs):
    return _DistributedOptimizer(optimizer, named_parameters, backward_passes_per_step, op)
