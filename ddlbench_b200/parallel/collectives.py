"""Generic hand-written collectives over symmetric memory (csrc/allreduce.cu): an in-place
all-reduce for arbitrary CUDA tensors (one-shot for small messages, two-shot / NVLS for
large ones) and a device-side barrier.  Used by ``hvd.allreduce``-style metric averaging on
GPUs, by the PipeDream replica groups, and by the micro-benchmarks in ``bench/``
(reference: ``pipedream-fork/runtime/tests/communication/all_to_all.py``)."""
from __future__ import annotations

import torch

from .. import ops
from ..dist.symm import SymmContext, SymmetricBuffer

ONE_SHOT_MAX_BYTES = 512 << 10
ONE_SHOT_MAX_BYTES_PAIR = 2 << 20     # two ranks: one phase, same bytes per direction up to here


class SymmAllreduce:
    def __init__(self, ctx: SymmContext, max_bytes: int = 64 << 20, blocks: int = 32):
        self.ctx = ctx
        # a single peer needs more CTAs to keep the link full (see csrc/allreduce.cu)
        self.blocks = min(blocks * 2, 64) if ctx.world <= 2 else blocks
        self.max_bytes = max_bytes
        # staging buffers per dtype share one allocation (viewed as bytes)
        self.stage: SymmetricBuffer = ctx.alloc(max_bytes // 4, torch.float32)

    def _quantum(self, itemsize: int) -> int:
        return self.ctx.world * (16 // itemsize)

    def allreduce_(self, t: torch.Tensor, average: bool = True, algo: str = "auto") -> torch.Tensor:
        """In-place all-reduce of a float32 / bfloat16 CUDA tensor of any size."""
        assert t.is_cuda and t.dtype in (torch.float32, torch.bfloat16)
        ctx = self.ctx
        flat = t.reshape(-1)
        item = flat.element_size()
        cap = self.max_bytes // item
        q = self._quantum(item)
        cap -= cap % q
        scale = 1.0 / ctx.world if average else 1.0
        stage = self.stage.tensor.view(torch.uint8)[: cap * item].view(t.dtype)
        for off in range(0, flat.numel(), cap):
            n = min(cap, flat.numel() - off)
            npad = (n + q - 1) // q * q
            stage[:n].copy_(flat[off:off + n])
            if npad > n:
                stage[n:npad].zero_()
            one_max = ONE_SHOT_MAX_BYTES_PAIR if ctx.world <= 2 else ONE_SHOT_MAX_BYTES
            use_one = algo == "oneshot" or (algo == "auto" and npad * item <= one_max)
            if use_one:
                out = torch.empty(npad, dtype=t.dtype, device=t.device)
                ops.ext().ar_oneshot(self.stage.ptrs, ctx.flags.ptrs, out, ctx.epoch, ctx.status, 0,
                                     scale, ctx.rank, min(self.blocks, 8))
                flat[off:off + n].copy_(out[:n])
            else:
                ops.ext().ar_twoshot(self.stage.ptrs, ctx.flags.ptrs, self.stage.mc_ptr,
                                     t.dtype == torch.bfloat16, ctx.epoch, ctx.status, 0,
                                     npad // ctx.world, scale, ctx.rank, self.blocks,
                                     # NVLS pays off from 3 ranks on: with one peer the switch
                                     # reduction only adds a hop (measured 285 vs 372 GB/s at 64 MiB)
                                     ctx.use_multimem and algo != "p2p" and (ctx.world > 2 or algo == "nvls"))
                flat[off:off + n].copy_(stage[:n])
            ops._count()
        return t

    def barrier(self) -> None:
        ops.ext().peer_barrier(self.ctx.flags.ptrs, self.ctx.epoch, self.ctx.status, self.ctx.rank)
        ops._count()
