"""Native-op front end.

``ops.<name>(...)`` runs the hand-written sm_100a kernel when its tensors live on a GPU and
the plain-PyTorch fp32 reference (``ops.reference``) on CPU, so every layer above this module
is device-agnostic and testable without a GPU.  On a machine that HAS a GPU the extension
must load: a missing ``.so`` raises instead of silently falling back (the driver records which
shared objects a GPU run actually loaded).

``ops.launches`` counts kernel launches issued through this module; ``bench.py`` reports it
as ``gpu_launches``.
"""
from __future__ import annotations

import importlib.util
import os
import threading
from typing import List, Optional, Sequence, Tuple

import torch

from . import reference as ref
from .reference import ACT_NONE, ACT_RELU, ACT_RELU6

ACT_BITMASK = 0x100   # flag or-ed into `act`: the `y` argument of the BN backward kernels is a bitmask

_ext = None
_ext_err: Optional[BaseException] = None
_lock = threading.Lock()
launches = 0  # number of native kernel launches issued through this module


def _count(n: int = 1) -> None:
    global launches
    launches += n


def load_ext(build_if_missing: bool = True):
    """Import the in-tree extension (building it first if the .so is absent)."""
    global _ext, _ext_err
    if _ext is not None:
        return _ext
    with _lock:
        if _ext is not None:
            return _ext
        from .build import build, so_path
        try:
            p = so_path()
            if not p.exists():
                if not build_if_missing:
                    raise FileNotFoundError(str(p))
                build()
            spec = importlib.util.spec_from_file_location("_ddlb_ext", str(p))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            assert mod.bn_replicas() == BN_REPLICAS, "BN_REPLICAS mismatch with the extension"
            if os.environ.get("DDLB_PDL", "1") == "0":     # profiling aid: true per-kernel durations
                mod.set_pdl(False)
            _ext = mod
        except BaseException as e:  # noqa: BLE001 - remembered and re-raised on GPU use
            _ext_err = e
            raise
    return _ext


def ext():
    """The extension module; raises loudly on a GPU box where it cannot be loaded."""
    if _ext is not None:
        return _ext
    try:
        return load_ext()
    except BaseException as e:  # noqa: BLE001
        raise RuntimeError(
            "ddlbench_b200 native extension (_ddlb_ext) is not available but a CUDA tensor "
            "was passed to a native op; run `python -m ddlbench_b200.ops.build`") from e


def available() -> bool:
    try:
        load_ext(build_if_missing=False)
        return True
    except BaseException:  # noqa: BLE001
        return False


def _on_gpu(t: torch.Tensor) -> bool:
    return t.is_cuda


# ------------------------------------------------------------------------------- optimizer
def sgd_flat_(master: torch.Tensor, mom: torch.Tensor, grad: torch.Tensor, hyper: torch.Tensor,
              w_out: torch.Tensor, w_out2: Optional[torch.Tensor] = None) -> None:
    """One fused SGD(momentum, wd) step on flat buffers; see csrc/sgd.cu."""
    if _on_gpu(master):
        ext().sgd_flat(master, mom, grad, w_out, w_out2, hyper)
        _count()
    else:
        ref.sgd_flat(master, mom, grad, hyper, [w_out, w_out2])


def l2norm(x: torch.Tensor) -> torch.Tensor:
    if _on_gpu(x):
        _count()
        return ext().sumsq(x.reshape(-1)).sqrt()
    return x.float().norm().reshape(1)


# ------------------------------------------------------------------------------- loss
class XentWorkspace:
    def __init__(self, rows: int, device):
        self.row_stats = torch.zeros(rows * 3, dtype=torch.float32, device=device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)


def xent(logits: torch.Tensor, target: torch.Tensor, smoothing: float = 0.0,
         grad_scale: float = 1.0, want_grad: bool = True,
         ws: Optional[XentWorkspace] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Fused softmax cross-entropy.  Returns (stats[3]: mean loss, #top-1 hits, #top-5 hits;
    dlogits = grad_scale * d(mean loss)/d(logits) or None)."""
    if _on_gpu(logits):
        if ws is None or ws.row_stats.numel() < logits.shape[0] * 3:
            ws = XentWorkspace(logits.shape[0], logits.device)
        logits = logits.contiguous()
        d = torch.empty_like(logits) if want_grad else None
        stats = ext().xent(logits, target, d, ws.row_stats, ws.ticket, float(smoothing),
                           float(grad_scale) / logits.shape[0])
        _count()
        return stats, d
    stats, d = ref.xent(logits, target, smoothing, grad_scale / logits.shape[0])
    return stats, (d if want_grad else None)


# ------------------------------------------------------------------------------- batch norm
def _mc(x: torch.Tensor) -> Tuple[int, int]:
    c = x.shape[1]
    return x.numel() // c, c


def _nhwc_ok(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and
            (x.dim() == 2 or x.is_contiguous(memory_format=torch.channels_last)))


# Single-pass cluster kernels (csrc/bn_cluster.cu).  Per kernel they beat the two-kernel path on
# the 7x7/14x14 layers (ncu: 10-14 us vs 15-24 us) but inside the training graph the step time is
# unchanged: 8/16-CTA clusters need whole-GPC slots and stop overlapping with the weight-gradient
# kernels on the side stream.  (Round 1: no gain inside the graph with the library's small wgrad
# kernels on the side stream; round 2, with every convolution on the tcgen05 kernels: 5.52 -> 5.39
# ms/step, profiles/r2/README.md.)  On by default; DDLB_BN_CLUSTER=0 disables them.
BN_CLUSTER_FUSED = os.environ.get("DDLB_BN_CLUSTER", "1") == "1"
# fused BN-fold GEMM kernel (csrc/gemm_bnfold_tcgen05.cu); without it a deferred BatchNorm is
# applied by the ordinary kernel right before its consumer (same result, no saving)
BNFOLD_KERNEL = os.environ.get("DDLB_BN_FOLD_KERNEL", "1") == "1"
BN_REPLICAS = 8  # must match DDL_BN_REPLICAS (csrc/common.cuh); checked at first use


# BatchNorm after a convolution whose epilogue accumulated the statistics: 1 = the apply kernel turns
# the sums into (mean, invstd) itself (no finalize in the convolution's tail, no finalize launch)
BN_APPLY_FINALIZE = os.environ.get("DDLB_BN_APPLY_FIN", "1") == "1"


class BnWorkspace:
    """Per-layer fp32 scratch: R replicated partial-sum buffers (self-zeroing), the
    last-CTA ticket counters, and the backward coefficients."""

    def __init__(self, channels: int, device):
        self.sums = torch.zeros(BN_REPLICAS * 2 * channels, dtype=torch.float32, device=device)
        # [0:32) block counters of the BN kernels, [32:64) legacy GEMM ticket, [64:) one finalize
        # ticket per 64-channel tile for the implicit-GEMM kernels
        self.counter = torch.zeros(64 + (channels + 63) // 64 + 1, dtype=torch.int32, device=device)
        self.coef = torch.empty(3 * channels, dtype=torch.float32, device=device)


def bn_forward(x, res, gamma, beta, running_mean, running_var, eps: float, momentum: float,
               act: int, ws: Optional[BnWorkspace] = None, stats_ready: bool = False,
               training: bool = True, stat_ready: Optional[torch.Tensor] = None,
               mask_out: Optional[torch.Tensor] = None):
    """y = act(BN(x) [+ res]).  Returns (y, stat[2, C]).  ``mask_out`` (uint8, numel/8): receives
    the activation derivative as one bit per element for ``bn_backward(mask=...)``.  ``stats_ready``: the producer
    (tcgen05 GEMM epilogue) already accumulated sum / sum-of-squares into ``ws.sums``."""
    M, C = _mc(x)
    if not training:
        stat = torch.stack([running_mean.float(), torch.rsqrt(running_var.float() + eps)])
        if _nhwc_ok(x):
            y = torch.empty_like(x)
            ext().bn_apply(x, res, y, stat.contiguous(), gamma, beta, None, M, C, act)
            _count()
            return y, stat
        return ref.bn_apply(x, res, stat, gamma, beta, act), stat
    if _nhwc_ok(x):
        e = ext()
        if ws is None:
            ws = BnWorkspace(C, x.device)
        stat = stat_ready if stat_ready is not None else \
            torch.empty(2 * C, dtype=torch.float32, device=x.device)
        if stat_ready is None and not stats_ready and BN_CLUSTER_FUSED:
            # single pass on a thread-block cluster (statistics + apply, x read once)
            y = torch.empty_like(x)
            if e.bn_fwd_fused(x, res, y, mask_out, stat, running_mean, running_var, gamma, beta,
                              M, C, eps, momentum, act):
                _count()
                return y, stat.view(2, C)
        if stat_ready is not None:
            _count(-1)      # (mean, invstd) were finalized inside the producing GEMM
        elif stats_ready and BN_APPLY_FINALIZE:
            # the apply kernel finalizes the statistics itself from the producer's sums
            y = torch.empty_like(x)
            if e.bn_apply_sums(x, res, y, ws.sums, ws.counter, stat, running_mean, running_var, gamma,
                               beta, mask_out, M, C, eps, momentum, act):
                _count()
                return y, stat.view(2, C)
            e.bn_finalize(ws.sums, stat, running_mean, running_var, M, C, eps, momentum)
        elif stats_ready:
            e.bn_finalize(ws.sums, stat, running_mean, running_var, M, C, eps, momentum)
        else:
            e.bn_stats(x, ws.sums, ws.counter, stat, running_mean, running_var, M, C, eps, momentum)
        y = torch.empty_like(x)
        e.bn_apply(x, res, y, stat, gamma, beta, mask_out, M, C, act)
        _count(2)
        return y, stat.view(2, C)
    stat = ref.bn_stat(x, eps)
    if running_mean is not None:
        ref.bn_update_running(stat, M, eps, momentum, running_mean, running_var)
    return ref.bn_apply(x, res, stat, gamma, beta, act), stat


def bn_statistics(x, running_mean, running_var, eps: float, momentum: float,
                  ws: Optional[BnWorkspace] = None, stats_ready: bool = False,
                  training: bool = True) -> torch.Tensor:
    """(mean, invstd) [2, C] of a BatchNorm WITHOUT applying it (the apply is folded into the
    consumer, see ``gemm_tn_bnfold``); updates the running statistics like ``bn_forward``."""
    M, C = _mc(x)
    if not training:
        return torch.stack([running_mean.float(), torch.rsqrt(running_var.float() + eps)])
    if _nhwc_ok(x):
        e = ext()
        if ws is None:
            ws = BnWorkspace(C, x.device)
        stat = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        if stats_ready:
            e.bn_finalize(ws.sums, stat, running_mean, running_var, M, C, eps, momentum)
        else:
            e.bn_stats(x, ws.sums, ws.counter, stat, running_mean, running_var, M, C, eps, momentum)
        _count()
        return stat.view(2, C)
    stat = ref.bn_stat(x, eps)
    if running_mean is not None:
        ref.bn_update_running(stat, M, eps, momentum, running_mean, running_var)
    return stat


class BnFold:
    """A BatchNorm(+residual)+activation whose application is deferred into its consumer:
    ``y`` (and ``mask``) are allocated but NOT yet written."""
    __slots__ = ("c", "res", "stat", "gamma", "beta", "act", "y", "mask")

    def __init__(self, c, res, stat, gamma, beta, act, y, mask):
        self.c, self.res, self.stat, self.gamma, self.beta = c, res, stat, gamma, beta
        self.act, self.y, self.mask = act, y, mask

    def materialize(self) -> torch.Tensor:
        """Plain (unfused) application: used when the consumer cannot fold."""
        if _nhwc_ok(self.c):
            M, C = _mc(self.c)
            ext().bn_apply(self.c, self.res, self.y, self.stat.reshape(-1).contiguous(), self.gamma,
                           self.beta, self.mask, M, C, self.act)
            _count()
        else:
            self.y.copy_(ref.bn_apply(self.c, self.res, self.stat, self.gamma, self.beta, self.act))
        return self.y


def gemm_tn_bnfold(fold: BnFold, b: torch.Tensor, col_sums: Optional[torch.Tensor] = None,
                   bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """D[M, N] = y @ b[N, K]^T with y = act(BN(c) [+ res]) computed ON THE WAY INTO the GEMM: the
    kernel loads the raw ``c`` tile, rewrites it in shared memory, stores it to ``fold.y`` (which
    the backward and any later consumer need) with its 1-bit mask, and feeds it to the tensor
    cores (csrc/gemm_bnfold_tcgen05.cu) — the element-wise BatchNorm pass never runs."""
    c = fold.c
    n, k, h, w = c.shape
    if BNFOLD_KERNEL and _nhwc_ok(c) and hasattr(ext(), "gemm_bf16_tn_bnfold"):
        a2d = c.permute(0, 2, 3, 1).reshape(n * h * w, k)
        y2d = fold.y.permute(0, 2, 3, 1).reshape(n * h * w, k)
        r2d = fold.res.permute(0, 2, 3, 1).reshape(n * h * w, k) if fold.res is not None else None
        out = torch.empty((n * h * w, b.shape[0]), dtype=c.dtype, device=c.device)
        if ext().gemm_bf16_tn_bnfold(a2d, r2d, b, out, y2d, fold.mask, fold.stat.reshape(-1), fold.gamma,
                                     fold.beta, fold.act, col_sums, bias, relu):
            _count()
            return out
    y = fold.materialize()
    a2d = y.permute(0, 2, 3, 1).reshape(n * h * w, k)
    return gemm_tn(a2d, b, col_sums=col_sums, bias=bias, relu=relu)


def bn_backward(dy, dy2, y, x, stat, gamma, act: int, want_dz: bool, dgamma_out, dbeta_out,
                accumulate: bool = False, ws: Optional[BnWorkspace] = None,
                mask: Optional[torch.Tensor] = None):
    """Backward of ``bn_forward``.  ``mask``: the bitmask ``bn_forward`` wrote (read instead of y).  Writes dgamma/dbeta into the given (flat-buffer) views.
    Returns (dx, dz) where dz (if requested) is the gradient flowing to the residual input."""
    M, C = _mc(x)
    if _nhwc_ok(x):
        e = ext()
        if ws is None:
            ws = BnWorkspace(C, x.device)
        stat_f = stat.reshape(-1)
        dz = torch.empty_like(x) if want_dz else None
        yy = y if act != ACT_NONE else None
        if mask is not None and act != ACT_NONE:
            yy, act = mask, act | ACT_BITMASK
        if BN_CLUSTER_FUSED:
            # single pass on a thread-block cluster when a channel slab fits in its shared memory
            dx = torch.empty_like(x)
            if e.bn_bwd_fused(dy, dy2, yy, x, dz, dx, stat_f, gamma, dgamma_out, dbeta_out,
                              accumulate, M, C, act):
                _count()
                return dx, dz
        e.bn_bwd_reduce(dy, dy2, yy, x, dz, stat_f, ws.sums, ws.counter, gamma, ws.coef,
                        dgamma_out, dbeta_out, accumulate, M, C, act)
        dx = torch.empty_like(x)
        if want_dz:
            e.bn_bwd_apply(dz, None, None, None, x, dx, stat_f, ws.coef, M, C, ACT_NONE)
        else:
            e.bn_bwd_apply(None, dy, dy2, yy, x, dx, stat_f, ws.coef, M, C, act)
        _count(2)
        return dx, dz
    dx, dz, dgamma, dbeta = ref.bn_backward(dy, dy2, y, x, stat, gamma, act, want_dz)
    if accumulate:
        dgamma_out.add_(dgamma.to(dgamma_out.dtype))
        dbeta_out.add_(dbeta.to(dbeta_out.dtype))
    else:
        dgamma_out.copy_(dgamma.to(dgamma_out.dtype))
        dbeta_out.copy_(dbeta.to(dbeta_out.dtype))
    return dx, dz


def bn_bwd_fold_supported(c, dy, dy2, act: int, mask, wt) -> bool:
    """Can the BatchNorm backward of ``c`` be folded into the 1x1 data-gradient GEMM
    (csrc/gemm_bnbwd_tcgen05.cu)?  Mirrors the kernel's own shape checks so that it never
    declines after ``bn_bwd_reduce`` has already written dgamma / dbeta."""
    if not (_nhwc_ok(c) and c.dim() == 4 and not BN_CLUSTER_FUSED):
        return False
    M, K = _mc(c)
    cl = torch.channels_last
    return (K % 64 == 0 and K <= 2048 and wt.dim() == 2 and wt.shape[1] == K and wt.shape[0] % 8 == 0 and
            wt.is_contiguous() and wt.dtype == torch.bfloat16 and
            dy.shape == c.shape and dy.dtype == c.dtype and dy.is_contiguous(memory_format=cl) and
            (dy2 is None or (dy2.shape == c.shape and dy2.dtype == c.dtype and
                             dy2.is_contiguous(memory_format=cl))) and
            (act == ACT_NONE or mask is not None) and hasattr(ext(), "gemm_bf16_tn_bnbwd"))


def bn_backward_dgrad_fold(dy, dy2, c, stat, gamma, act: int, want_dz: bool, dgamma_out, dbeta_out,
                           accumulate: bool, ws: Optional[BnWorkspace], mask, wt):
    """BatchNorm backward + the data gradient of the 1x1 convolution that produced ``c``, with the
    element-wise half of the BatchNorm backward folded into the GEMM's operand load:
    ``bn_bwd_reduce`` (sums, dgamma, dbeta, coefficients), then ONE kernel that rebuilds
    dz = (dy [+ dy2]) * act' and dc = BN'(dz, c) tile by tile in shared memory, multiplies by
    ``wt`` = W^T [Cin, Cout] and stores dc (for the weight gradient) and dz (residual gradient,
    if ``want_dz``) on the way.  Call only when ``bn_bwd_fold_supported``.
    Returns (dc [N, C, H, W], dz or None, dx2d [N*H*W, Cin])."""
    M, C = _mc(c)
    e = ext()
    if ws is None:
        ws = BnWorkspace(C, c.device)
    stat_f = stat.reshape(-1)
    yy, flags = (mask, act | ACT_BITMASK) if act != ACT_NONE else (None, act)
    e.bn_bwd_reduce(dy, dy2, yy, c, None, stat_f, ws.sums, ws.counter, gamma, ws.coef,
                    dgamma_out, dbeta_out, accumulate, M, C, flags)
    dc = torch.empty_like(c)
    dz = torch.empty_like(c) if want_dz else None
    dx2d = torch.empty((M, wt.shape[0]), dtype=c.dtype, device=c.device)

    def rows(t):
        if t is None:
            return None
        v = t.permute(0, 2, 3, 1).reshape(M, C)
        assert v.data_ptr() == t.data_ptr(), "NHWC operand expected (reshape must not copy)"
        return v

    ok = e.gemm_bf16_tn_bnbwd(rows(dy), rows(dy2), rows(c), yy, wt, dx2d, rows(dc), rows(dz), stat_f,
                              ws.coef, act)
    assert ok, "gemm_bf16_tn_bnbwd declined a shape bn_bwd_fold_supported accepted"
    _count(2)
    return dc, dz, dx2d


def bias_act(x, bias, act: int):
    if _nhwc_ok(x):
        M, C = _mc(x)
        y = torch.empty_like(x)
        ext().bias_act(x, bias, y, M, C, act)
        _count()
        return y
    return ref.bias_act(x, bias, act)


def act_bwd(dy, dy2, y, act: int):
    """dz = (dy [+ dy2]) * act'(y)."""
    if dy.is_cuda and dy.dtype == torch.bfloat16 and dy.numel() % 8 == 0 and \
            (dy2 is None or dy2.stride() == dy.stride()) and (y is None or y.stride() == dy.stride()):
        dz = torch.empty_like(dy)
        ext().act_bwd(dy, dy2, y if act != ACT_NONE else None, dz, act)
        _count()
        return dz
    return ref.act_bwd(dy, dy2, y, act)


def bias_grad_(dy, out, accumulate: bool = False, ws: Optional[BnWorkspace] = None):
    """out[c] (+)= sum over all non-channel dims of dy."""
    if _nhwc_ok(dy):
        M, C = _mc(dy)
        if ws is None:
            ws = BnWorkspace(C, dy.device)
        ext().colsum(dy, ws.sums, ws.counter, out, accumulate, M, C)
        _count()
        return
    dims = [0] + list(range(2, dy.dim()))
    g = dy.float().sum(dim=dims)
    if accumulate:
        out.add_(g.to(out.dtype))
    else:
        out.copy_(g.to(out.dtype))


# ------------------------------------------------------------------------------- misc
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_u8(x_u8: torch.Tensor, mean: Sequence[float], std: Sequence[float],
                 dtype=torch.bfloat16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [N, H, W, C] -> normalised logical [N, C, H, W] channels-last tensor."""
    n, h, w, c = x_u8.shape
    if x_u8.is_cuda and dtype == torch.bfloat16:
        if out is None:
            out = torch.empty((n, c, h, w), dtype=dtype, device=x_u8.device,
                              memory_format=torch.channels_last)
        ext().u8_to_bf16_norm(x_u8.contiguous(), out, c, list(mean), list(std))
        _count()
        return out
    y = ref.u8_to_norm(x_u8, list(mean), list(std), dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def revert_varlen(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        _count()
        return ext().revert_varlen(x.contiguous(), lengths.to(x.device, torch.int64))
    return ref.revert_varlen(x, lengths)


def set_mask(lengths: torch.Tensor, T: int) -> torch.Tensor:
    if lengths.is_cuda:
        _count()
        return ext().set_mask(lengths.to(torch.int64), T)
    return ref.set_mask(lengths, T)


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """[N, C, H, W] -> [N, C]."""
    if _nhwc_ok(x) and x.dim() == 4:
        n, c, h, w = x.shape
        y = torch.empty((n, c), dtype=x.dtype, device=x.device)
        ext().gap_fwd(x, y, n, h * w, c)
        _count()
        return y
    return ref.global_avg_pool(x)


def global_avg_pool_bwd(dy: torch.Tensor, shape) -> torch.Tensor:
    n, c, h, w = shape
    if dy.is_cuda and dy.dtype == torch.bfloat16 and c % 8 == 0:
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device,
                         memory_format=torch.channels_last)
        ext().gap_bwd(dy.contiguous(), dx, n, h * w, c)
        _count()
        return dx
    return (dy.float() / (h * w)).to(dy.dtype).view(n, c, 1, 1).expand(n, c, h, w).contiguous()


def maxpool_supported(x: torch.Tensor, k, s, p, dilation, ceil_mode) -> bool:
    return (_nhwc_ok(x) and x.dim() == 4 and k[0] == k[1] and s[0] == s[1] and p[0] == p[1] and
            list(dilation) == [1, 1] and not ceil_mode and 1 <= k[0] <= 15)


def maxpool_fwd(x: torch.Tensor, k: int, s: int, p: int, want_idx: bool = True):
    """NHWC bf16 max-pool; returns (y, idx) with idx = arg-max position inside the window."""
    n, c, h, w = x.shape
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    y = torch.empty((n, c, ho, wo), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last)
    idx = torch.empty((n, c, ho, wo), dtype=torch.uint8, device=x.device,
                      memory_format=torch.channels_last) if want_idx else None
    ext().maxpool_fwd(x, y, idx, k, s, p)
    _count()
    return y, idx


def maxpool_bwd(dy: torch.Tensor, idx: torch.Tensor, in_shape, k: int, s: int, p: int):
    dx = torch.empty(tuple(in_shape), dtype=dy.dtype, device=dy.device,
                     memory_format=torch.channels_last)
    ext().maxpool_bwd(dy.contiguous(memory_format=torch.channels_last), idx, dx, k, s, p)
    _count()
    return dx



# ------------------------------------------------------------------------------- avg pool / dropout
def avgpool_supported(x: torch.Tensor) -> bool:
    return _nhwc_ok(x) and x.dim() == 4


def avgpool_fwd(x: torch.Tensor, out_hw, k: int, s: int, p: int, adaptive: bool,
                count_include_pad: bool = True) -> Optional[torch.Tensor]:
    """NHWC bf16 average pool (fixed window or adaptive); None if the kernel declines."""
    n, c = x.shape[:2]
    y = torch.empty((n, c, out_hw[0], out_hw[1]), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last)
    if ext().avgpool(False, x, y, k, s, p, adaptive, count_include_pad):
        _count()
        return y
    return None


def avgpool_bwd(dy: torch.Tensor, in_shape, k: int, s: int, p: int, adaptive: bool,
                count_include_pad: bool = True) -> Optional[torch.Tensor]:
    dx = torch.empty(tuple(in_shape), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
    if ext().avgpool(True, dy.contiguous(memory_format=torch.channels_last), dx, k, s, p, adaptive,
                     count_include_pad):
        _count()
        return dx
    return None


_RNG_STEP: dict = {}


def rng_step(device) -> torch.Tensor:
    """Device-resident step counter mixed into the dropout Philox counter (one int64 per device)."""
    device = torch.device(device)
    t = _RNG_STEP.get(device)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _RNG_STEP[device] = t
    return t


def rng_tick(device) -> None:
    """Advance the step counter — once per training step, INSIDE a captured graph (a plain in-place
    add), so every replay draws fresh dropout masks while a recomputed forward of the same step
    (GPipe checkpointing) reproduces them."""
    if torch.device(device).type == "cuda":
        rng_step(device).add_(1)


def dropout_fwd(x: torch.Tensor, p: float):
    """-> (y, mask bits uint8[numel/8]) on the native kernel, or None when not applicable."""
    dense = x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.numel() % 8 == 0 and dense):
        return None
    y = torch.empty_like(x)
    mask = torch.empty(x.numel() // 8, dtype=torch.uint8, device=x.device)
    # seed: torch's CUDA seed; salt: 63 fresh bits from the HOST generator (a restored RNG state —
    # GPipe recompute — gives the same salt; every other call a different one)
    seed = torch.cuda.initial_seed() & 0x7FFFFFFFFFFFFFFF
    salt = int(torch.randint(0, 2 ** 62, (1,)).item())
    ext().dropout_fwd(x, y, mask, float(p), seed, salt, rng_step(x.device))
    _count()
    return y, mask


def dropout_bwd(dy: torch.Tensor, mask: torch.Tensor, p: float) -> torch.Tensor:
    dy = dy if (dy.is_contiguous() or (dy.dim() == 4 and dy.is_contiguous(memory_format=torch.channels_last))) \
        else dy.contiguous()
    dx = torch.empty_like(dy)
    ext().dropout_bwd(dy, mask, dx, 1.0 / (1.0 - p))
    _count()
    return dx

# ------------------------------------------------------------------------------- GEMM
def gemm_tn_supported(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and
            a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and
            a.shape[1] % 8 == 0 and b.shape[0] % 8 == 0 and a.stride(0) % 8 == 0 and
            b.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)


def gemm_tn(a: torch.Tensor, b: torch.Tensor, col_sums: Optional[torch.Tensor] = None,
            bias: Optional[torch.Tensor] = None, relu: bool = False,
            out: Optional[torch.Tensor] = None, bn_finalize=None) -> torch.Tensor:
    """D[M, N] = A[M, K] @ B[N, K]^T on the tcgen05 tensor cores (csrc/gemm_tcgen05.cu).
    ``col_sums`` (the [R, 2, N] fp32 replicated BN workspace, zero on entry) receives the
    per-column sum / sum-of-squares of D spread over its R replicas.  ``bn_finalize`` =
    (ticket, stat_out[2N], running_mean, running_var, eps, momentum): the last CTA also turns
    the sums into (mean, invstd), updates the running statistics and re-zeroes the workspace."""
    if gemm_tn_supported(a, b):
        if out is None:
            out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.bfloat16, device=a.device)
        if bn_finalize is not None and col_sums is not None:
            tk, st, rm, rv, eps, mom = bn_finalize
            ext().gemm_bf16_tn(a, b, out, col_sums, bias, relu, tk, st, rm, rv, float(eps), float(mom))
        else:
            ext().gemm_bf16_tn(a, b, out, col_sums, bias, relu, None, None, None, None, 0.0, 0.0)
        _count()
        return out
    up = torch.float32 if a.dtype in (torch.float16, torch.bfloat16) else a.dtype
    d = a.to(up) @ b.to(up).t()
    if bias is not None:
        d = d + bias.to(up)
    if relu:
        d = torch.relu(d)
    d = d.to(a.dtype)
    if col_sums is not None:
        df = d.to(col_sums.dtype)
        n = d.shape[1]
        col_sums[:n].add_(df.sum(0))
        col_sums[n:2 * n].add_((df * df).sum(0))
    if out is not None:
        out.copy_(d)
        return out
    return d


def gemm_nt_mn_supported(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and
            b.dim() == 2 and a.shape[0] == b.shape[0] and a.stride(1) == 1 and b.stride(1) == 1 and
            a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0 and a.stride(0) % 8 == 0 and
            b.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)


def gemm_nt_mn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None,
               splits: int = 0) -> torch.Tensor:
    """out[M, N] (fp32) += a[K, M]^T @ b[K, N] — the weight-gradient GEMM (reduction over the slow
    dimension of both operands, csrc/gemm_wgrad_tcgen05.cu).  ``out`` is zero-initialised when
    omitted; ``splits`` = k-splits per output tile (0: fill the machine)."""
    if out is None:
        out = torch.zeros((a.shape[1], b.shape[1]), dtype=torch.float32, device=a.device)
    if gemm_nt_mn_supported(a, b) and out.dtype == torch.float32 and out.stride(1) == 1:
        ext().gemm_bf16_nt_mn(a, b, out, splits, None)
        _count()
        return out
    out.add_(a.float().t() @ b.float())
    return out



# ------------------------------------------------------------------------------- implicit GEMM
# Persistent tcgen05 implicit-GEMM family (csrc/igemm_tcgen05.cu) + one-launch weight gradients
# (csrc/gemm_wgrad_tcgen05.cu): every dense convolution forward / data gradient / weight gradient
# and the FC layers.  DDLB_IGEMM=0 routes back to the library (A/B and bring-up only).
IGEMM = os.environ.get("DDLB_IGEMM", "1") == "1"
_WG_WS: dict = {}


def _wgrad_ws(device, numel: int = 0):
    """fp32 scratch for the split-K planes of the weight-gradient kernels that run back to back on
    one stream (48 MB: planes of the widest layers; a layer that needs more planes than fit simply
    uses fewer k-splits).  Never needs zeroing."""
    key = (device, torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0)
    ent = _WG_WS.get(key)
    if ent is None:
        ent = torch.empty(12 << 20, dtype=torch.float32, device=device)
        _WG_WS[key] = ent
    return ent


_IG_WS: dict = {}
_IG_SPLIT_TILES = 448        # slots (tile x k-slice) of 128x128 fp32: 28 MB


def _ig_ws(device):
    """fp32 scratch (partial tiles, never needs zeroing) + zero tickets (self-resetting) of the
    split-K implicit-GEMM launches of one stream."""
    key = (device, torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0)
    ent = _IG_WS.get(key)
    if ent is None:
        ent = (torch.empty(_IG_SPLIT_TILES * 16384, dtype=torch.float32, device=device),
               torch.zeros(_IG_SPLIT_TILES, dtype=torch.int32, device=device))
        _IG_WS[key] = ent
    return ent


def igemm_conv_supported(x: torch.Tensor, wshape, stride, padding, dilation, groups,
                         stats_channels: int = 0) -> bool:
    """Shapes the persistent implicit-GEMM kernels take: bf16 NHWC, groups 1, dilation 1, square
    stride 1 or 2, channel counts that are multiples of 8 (TMA 16-byte rows)."""
    if not (IGEMM and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    co, ci, kh, kw = wshape
    st = list(stride)
    return (st[0] == st[1] and st[0] in (1, 2) and list(dilation) == [1, 1] and groups == 1 and
            kh * kw <= 49 and kh <= 7 and kw <= 7 and ci % 8 == 0 and co % 8 == 0 and
            stats_channels <= 2048 and padding[0] < kh + st[0] and padding[1] < kw + st[0] and
            x.is_contiguous(memory_format=torch.channels_last))


def _finalize_after(col_sums, bn_finalize, rows: int, channels: int) -> None:
    """The kernel delivered the sums only (its one-tile variant: ~1500 short CTAs, where a per-CTA
    fence + ticket costs more than this 2 us launch): finalize them here, same outputs."""
    _, st, rm, rv, eps, mom = bn_finalize
    ext().bn_finalize(col_sums, st, rm, rv, rows, channels, float(eps), float(mom))
    _count()


def _fin_args(bn_finalize):
    if bn_finalize is None:
        return (None, None, None, None, 0.0, 0.0)
    tk, st, rm, rv, eps, mom = bn_finalize
    return (tk, st, rm, rv, float(eps), float(mom))


def conv_fprop(x: torch.Tensor, w: torch.Tensor, stride: int, padding: Sequence[int],
               col_sums: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
               relu: bool = False, bn_finalize=None) -> Optional[torch.Tensor]:
    """y = conv2d(x, w, stride, padding) [+bias, ReLU] with the BatchNorm sum / sum-of-squares (and,
    with ``bn_finalize``, mean / invstd / running statistics) produced by the same kernel."""
    co, ci, kh, kw = w.shape
    ph, pw = int(padding[0]), int(padding[1])
    oh = (x.shape[2] + 2 * ph - kh) // stride + 1
    ow = (x.shape[3] + 2 * pw - kw) // stride + 1
    if not w.is_contiguous(memory_format=torch.channels_last):
        w = w.contiguous(memory_format=torch.channels_last)
    y = torch.empty((x.shape[0], co, oh, ow), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last)
    rc = ext().igemm_conv_fprop(x, w, y, ph, pw, stride, col_sums, bias, relu, *_fin_args(bn_finalize),
                                *_ig_ws(x.device))
    if rc < 0:
        return None
    _count()
    if bn_finalize is not None and rc == 0:
        _finalize_after(col_sums, bn_finalize, y.numel() // co, co)
    return y


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, x_shape, stride: int,
               padding: Sequence[int]) -> Optional[torch.Tensor]:
    """Data gradient of conv2d(x, w): reads the FORWARD filter (MN-major B operand), no transposes."""
    if not w.is_contiguous(memory_format=torch.channels_last):
        w = w.contiguous(memory_format=torch.channels_last)
    dx = torch.empty(tuple(x_shape), dtype=dy.dtype, device=dy.device,
                     memory_format=torch.channels_last)
    if ext().igemm_conv_dgrad(dy, w, dx, int(padding[0]), int(padding[1]), stride, *_ig_ws(dy.device)):
        _count()
        return dx
    return None


def conv_wgrad_into(dy: torch.Tensor, x: torch.Tensor, gw: torch.Tensor, stride: int,
                    padding: Sequence[int], accumulate: bool) -> bool:
    """gw (+)= weight gradient of conv2d(x, w); ``gw`` is the [Cout, Cin, KH, KW] KRSC view of the
    flat gradient buffer (split-K partial tiles go to fp32 planes that a small reduce kernel sums;
    no atomics, no zero-fill, no separate conversion pass)."""
    co, ci, kh, kw = gw.shape
    if not (gw.is_cuda and gw.is_contiguous(memory_format=torch.channels_last) and
            gw.dtype in (torch.bfloat16, torch.float32) and gw.data_ptr() % 16 == 0):
        return False
    ws = _wgrad_ws(gw.device)
    if kh == 1 and kw == 1 and stride == 1 and padding[0] == 0 and padding[1] == 0:
        n, _, h, wd = x.shape
        a = dy.permute(0, 2, 3, 1).reshape(n * h * wd, co)
        b = x.permute(0, 2, 3, 1).reshape(n * h * wd, ci)
        ok = ext().wgrad_gemm(a, b, ws, gw.permute(0, 2, 3, 1).reshape(co, ci), accumulate)
    else:
        ok = ext().wgrad_conv(dy, x, ws, gw, accumulate, int(padding[0]), int(padding[1]),
                              stride)
    if ok:
        _count()
    return bool(ok)


def igemm_ok2d(a: torch.Tensor, n: int) -> bool:
    return (IGEMM and a.is_cuda and a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1 and
            a.shape[1] % 8 == 0 and a.stride(0) % 8 == 0 and n % 8 == 0 and a.data_ptr() % 16 == 0)


def gemm_ig(a: torch.Tensor, b: torch.Tensor, b_mn: bool = False,
            col_sums: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
            relu: bool = False, bn_finalize=None) -> Optional[torch.Tensor]:
    """D[M, N] = a[M, K] @ (b[N, K]^T if not b_mn else b[K, N]) on the persistent kernel."""
    n = b.shape[1] if b_mn else b.shape[0]
    if (GEMM_PAIR and not b_mn and col_sums is None and bn_finalize is None and
            a.shape[0] >= GEMM_PAIR_MIN[0] and n >= GEMM_PAIR_MIN[1] and a.shape[1] >= GEMM_PAIR_MIN[2]):
        out = gemm_pair(a, b, bias, relu)
        if out is not None:
            return out
    out = torch.empty((a.shape[0], n), dtype=torch.bfloat16, device=a.device)
    rc = ext().igemm_gemm(a, b, out, b_mn, col_sums, bias, relu, *_fin_args(bn_finalize), *_ig_ws(a.device))
    if rc < 0:
        return None
    _count()
    if bn_finalize is not None and rc == 0:
        _finalize_after(col_sums, bn_finalize, a.shape[0], n)
    return out


GEMM_PAIR = os.environ.get("DDLB_GEMM_PAIR", "1") == "1"
# (M, N, K) from which the CTA-pair kernel takes over from the single-CTA one: a 256-row tile grid
# of at least one wave of 74 clusters and a reduction long enough to be bandwidth- rather than
# epilogue-bound (profiles/gemm_tcgen05_vs_cublas.json)
GEMM_PAIR_MIN = (2048, 256, 512)


def gemm_pair(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
              relu: bool = False, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """D[M, N] = a[M, K] @ b[N, K]^T (+ bias)(ReLU) on the CTA-pair (``cta_group::2``) persistent
    kernel (csrc/gemm2cta_tcgen05.cu); None if the shape is declined."""
    if not (a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and
            b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and a.shape[1] == b.shape[1]):
        return None
    if a.shape[0] < 32 or b.shape[0] < 64:
        return None
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.bfloat16, device=a.device)
    if not ext().gemm2cta(a, b, out, bias, relu):
        return None
    _count()
    return out


def wgrad_gemm_into(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, accumulate: bool) -> bool:
    """out[M, N] (+)= a[K, M]^T @ b[K, N] in one launch (see ``conv_wgrad_into``)."""
    if not (IGEMM and gemm_nt_mn_supported(a, b) and out.is_cuda and out.dim() == 2 and
            out.stride(1) == 1 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0 and
            out.dtype in (torch.bfloat16, torch.float32)):
        return False
    if ext().wgrad_gemm(a, b, _wgrad_ws(out.device), out, accumulate):
        _count()
        return True
    return False

def im2col_rows(x: torch.Tensor, kh: int, kw: int, stride: int, pad: int) -> torch.Tensor:
    """[N, C, H, W] -> patch matrix [N*OH*OW, roundup(kh*kw*C, 8)] with K ordered (kh, kw, c) — the
    A operand that turns a few-channel first convolution into a tensor-core GEMM
    (csrc/im2col.cu)."""
    n, c, h, w = x.shape
    oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    k = kh * kw * c
    kpad = (k + 7) // 8 * 8
    if x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last):
        a = torch.empty((n * oh * ow, kpad), dtype=x.dtype, device=x.device)
        if ext().im2col_rows(x, a, kh, kw, stride, pad):
            _count()
            return a
    return ref.im2col_rows(x, kh, kw, stride, pad)


def conv_tn_supported(x: torch.Tensor, w: torch.Tensor, stride, padding, dilation, groups) -> bool:
    """Shapes the tcgen05 implicit-GEMM convolution (csrc/conv_tcgen05.cu) takes: stride 1 or 2,
    no groups/dilation, Cin % 64 == 0, Cout % 64 == 0, bf16 NHWC (1x1 stride-1 convolutions are
    plain GEMMs and go through ``gemm_tn`` instead)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    co, ci, kh, kw = w.shape
    st = list(stride)
    return (st[0] == st[1] and st[0] in (1, 2) and list(dilation) == [1, 1] and groups == 1 and
            (kh * kw > 1 or st[0] == 2) and kh <= 7 and kw <= 7 and ci % 64 == 0 and co % 64 == 0 and
            x.is_contiguous(memory_format=torch.channels_last) and
            (w.is_contiguous(memory_format=torch.channels_last) or kh * kw == 1))


def conv_tn(x: torch.Tensor, w: torch.Tensor, col_sums: Optional[torch.Tensor] = None,
            bias: Optional[torch.Tensor] = None, relu: bool = False, stride: int = 1,
            padding: Optional[Sequence[int]] = None) -> Optional[torch.Tensor]:
    """y = conv2d(x, w, stride, padding) with optional fused bias/ReLU and BatchNorm partial
    sums, on the tensor cores.  ``padding`` defaults to "same" (k // 2).  Returns None if the
    kernel declines the shape."""
    co, ci, kh, kw = w.shape
    ph, pw = (kh // 2, kw // 2) if padding is None else (int(padding[0]), int(padding[1]))
    oh = (x.shape[2] + 2 * ph - kh) // stride + 1
    ow = (x.shape[3] + 2 * pw - kw) // stride + 1
    y = torch.empty((x.shape[0], co, oh, ow), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last)
    if kh * kw == 1 and not w.is_contiguous(memory_format=torch.channels_last):
        w = w.contiguous(memory_format=torch.channels_last)
    if ext().conv_bf16_nhwc(x, w, y, ph, pw, stride, col_sums, bias, relu):
        _count()
        return y
    return None


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, wshape, stride: int = 1,
               padding: Optional[Sequence[int]] = None) -> Optional[torch.Tensor]:
    """Weight gradient [Cout, Cin, KH, KW] (fp32, KRSC memory) of a convolution from dy and x on
    the tensor cores (csrc/gemm_wgrad_tcgen05.cu).  Returns None if the kernel declines."""
    co, ci, kh, kw = wshape
    ph, pw = (kh // 2, kw // 2) if padding is None else (int(padding[0]), int(padding[1]))
    dw = torch.zeros((co, kh, kw, ci), dtype=torch.float32, device=x.device)
    if ext().conv_wgrad_bf16_nhwc(dy, x, dw, ph, pw, stride):
        _count()
        return dw.permute(0, 3, 1, 2)
    return None


def dwconv_supported(x: torch.Tensor, wshape, stride, padding, dilation, groups) -> bool:
    """Depthwise 3x3 (groups == channels), stride 1 or 2, bf16 NHWC (csrc/dwconv.cu)."""
    co, cig, kh, kw = wshape
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and groups == co == x.shape[1] and
            cig == 1 and kh == 3 and kw == 3 and list(dilation) == [1, 1] and stride[0] == stride[1] and
            stride[0] in (1, 2) and padding[0] == padding[1] and 0 <= padding[0] <= 2 and co % 8 == 0 and
            x.is_contiguous(memory_format=torch.channels_last))


def _dw_taps(w: torch.Tensor) -> torch.Tensor:
    """[C, 1, 3, 3] -> tap-major [9, C] (the layout the kernels read with 16-byte vectors)."""
    return w.reshape(w.shape[0], 9).t().contiguous()


def dwconv_fwd(x: torch.Tensor, w: torch.Tensor, stride: int, pad: int) -> torch.Tensor:
    n, c, h, wd = x.shape
    oh, ow = (h + 2 * pad - 3) // stride + 1, (wd + 2 * pad - 3) // stride + 1
    y = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    assert ext().dwconv3x3(0, x, _dw_taps(w), y, h, wd, stride, pad)
    _count()
    return y


def dwconv_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, stride: int, pad: int,
               need_dx: bool = True):
    """-> (dx or None, dw [C, 1, 3, 3] fp32)."""
    n, c, h, wd = x.shape
    e = ext()
    dx = None
    if need_dx:
        dx = torch.empty_like(x)
        assert e.dwconv3x3(1, dy, _dw_taps(w), dx, h, wd, stride, pad)
        _count()
    dwt = torch.zeros((9, c), dtype=torch.float32, device=x.device)
    assert e.dwconv3x3(2, dy, x, dwt, h, wd, stride, pad)
    _count()
    return dx, dwt.t().reshape(c, 1, 3, 3)


__all__ = [n for n in dir() if not n.startswith("_")]
