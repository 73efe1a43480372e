"""Fused layer ops with explicit forward AND backward (no autograd tape).

Each op is the unit the static executor schedules: conv+BN+activation(+residual) is one op,
conv+bias+activation is one op, and so on.  Library GEMM/conv kernels (cuDNN/cuBLAS through
``torch.ops.aten``) are used where no hand-written kernel exists yet; everything elementwise
or reduction-shaped runs in the hand-written kernels of ``ddlbench_b200.ops`` (on CPU: their
fp32 references), and 1x1 convolutions / FC layers can run on the tcgen05 GEMM
(``csrc/gemm_tcgen05.cu``) with the BatchNorm statistics fused into its epilogue.

Parameter gradients are written straight into views of the flat gradient buffer
(write-not-accumulate on the first micro-batch), so no ``zero_grad`` pass exists.
"""
from __future__ import annotations

import os as _os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, ACT_RELU6

aten = torch.ops.aten
Ctx = Dict[str, object]
# set by the executor: route 1x1 convs / FC through the tcgen05 GEMM kernels
TCGEN05 = {"conv1x1": False, "linear": False}
# 0.0 while a forward is being RE-computed (GPipe checkpointing) so BatchNorm running
# statistics are not updated twice for the same micro-batch
BN_MOMENTUM_SCALE = [1.0]
# side stream for weight-gradient kernels (set by the executor during backward) and the
# operands that must outlive it
WGRAD_OVERLAP = {"stream": None, "keep": []}

# ---- switches; read once at import
# BN finalize inside the GEMM epilogue (measured slower than a separate 4 us launch)
FUSE_BN_FINALIZE = _os.environ.get("DDLB_FUSE_FINALIZE", "0") == "1"
# BN backward reads the 1-bit activation mask written by the forward instead of y
ACT_BITMASK_ON = _os.environ.get("DDLB_ACT_BITMASK", "1") == "1"
# round-1 one-tile implicit GEMM for the 3x3 convolutions (csrc/conv_tcgen05.cu): validated, A/B baseline
# of the persistent family (only consulted with DDLB_IGEMM=0)
CONV3X3_TCGEN05 = _os.environ.get("DDLB_CONV3X3_TCGEN05", "0") == "1"
# first convolution as im2col + round-1 GEMM with DDLB_IGEMM=0 (the default path always uses im2col + igemm)
STEM_IM2COL = _os.environ.get("DDLB_STEM_IM2COL", "0") == "1"
# BatchNorm application deferred into the consuming 1x1-convolution GEMM (opt-in, see
# ops.gemm_tn_bnfold); pending (not yet written) outputs are tracked by object identity
BN_FOLD = _os.environ.get("DDLB_BN_FOLD", "0") == "1"
PENDING_FOLD: Dict[int, "ops.BnFold"] = {}
# hand-written depthwise 3x3 convolution (csrc/dwconv.cu; numerics validated on hardware:
# tests/test_kernels_gpu.py::test_dwconv3x3): default on, DDLB_DWCONV=0 falls back to cuDNN's grouped conv
DWCONV = _os.environ.get("DDLB_DWCONV", "1") == "1"
# weight gradient of 1x1 convolutions / linear layers on tcgen05 (csrc/gemm_wgrad_tcgen05.cu): opt-in
WGRAD_TCGEN05 = _os.environ.get("DDLB_WGRAD_TCGEN05", "0") == "1"
# data gradient of those 3x3 convolutions through the same kernel (flipped, transposed filter)
CONV3X3_DGRAD = CONV3X3_TCGEN05 and _os.environ.get("DDLB_CONV3X3_DGRAD", "0") == "1"
# element-wise half of the BatchNorm backward folded into the 1x1 data-gradient GEMM's operand
# load (csrc/gemm_bnbwd_tcgen05.cu, ops.bn_backward_dgrad_fold): validated, round-1 family only
BN_BWD_FOLD = _os.environ.get("DDLB_BN_BWD_FOLD", "0") == "1"
# data gradient of the 1x1 convolutions on the (validated) tcgen05 TN GEMM: dx = dy @ W needs W^T
# as the K-major B operand, i.e. one small transpose per layer and step; round-1 family only (the default dgrad reads the filter MN-major)
# against the library's dgrad
DGRAD1X1_TCGEN05 = _os.environ.get("DDLB_DGRAD1X1_TCGEN05", "0") == "1"
# weight-gradient kernels on a side stream during backward
OVERLAP_WGRAD_DEFAULT = _os.environ.get("DDLB_WGRAD_OVERLAP", "1") == "1"


def _write(dst: torch.Tensor, src: torch.Tensor, accumulate: bool) -> None:
    if accumulate:
        dst.add_(src.to(dst.dtype) if src.dtype != dst.dtype else src)
    else:
        dst.copy_(src)


def _sum_grads(dys: Sequence[torch.Tensor]) -> torch.Tensor:
    out = dys[0]
    for d in dys[1:]:
        out = out + d
    return out


def _pair(v) -> List[int]:
    return [v, v] if isinstance(v, int) else list(v)


class Op:
    kind = "op"

    def __init__(self, name: str, inputs: Sequence[str], output: str):
        self.name = name
        self.inputs = list(inputs)
        self.output = output
        self.params: List[str] = []       # names inside the ParamStore
        self.order = 0                    # forward execution index

    def param_specs(self):                # -> List[ParamSpec]
        return []

    def buffers(self) -> Dict[str, torch.Tensor]:
        return {}

    def forward(self, P, B, xs: List[torch.Tensor], ctx: Ctx, training: bool) -> torch.Tensor:
        raise NotImplementedError

    def backward(self, P, G, dys: List[torch.Tensor], ctx: Ctx, accumulate: bool,
                 need: List[bool]) -> List[Optional[torch.Tensor]]:
        raise NotImplementedError

    def to_device(self, device) -> None:
        pass

    def __repr__(self):
        return f"{self.kind}({self.name}: {self.inputs} -> {self.output})"


# ------------------------------------------------------------------------------- conv helpers
class _ConvMixin:
    def _conv_cfg(self, conv: nn.Conv2d):
        self.stride = _pair(conv.stride)
        self.padding = _pair(conv.padding)
        self.dilation = _pair(conv.dilation)
        self.groups = conv.groups
        self.wshape = tuple(conv.weight.shape)

    def _is_gemm(self, x: torch.Tensor) -> bool:
        co, ci, kh, kw = self.wshape
        return (TCGEN05["conv1x1"] and kh == 1 and kw == 1 and self.stride == [1, 1] and
                self.padding == [0, 0] and self.groups == 1 and x.is_cuda and
                x.dtype == torch.bfloat16 and ci % 8 == 0 and co % 8 == 0 and
                x.is_contiguous(memory_format=torch.channels_last))

    def _is_ig(self, x: torch.Tensor, stats: bool = False) -> bool:
        """Persistent tcgen05 implicit-GEMM family (default path of every dense convolution)."""
        return TCGEN05["conv1x1"] and ops.igemm_conv_supported(
            x, self.wshape, self.stride, self.padding, self.dilation, self.groups,
            self.wshape[0] if stats else 0)

    def _is_stem_gemm(self, x: torch.Tensor) -> bool:
        """Few-channel first convolution (7x7/s2 on RGB, 3x3 on RGB/grey) as patch matrix + GEMM."""
        co, ci, kh, kw = self.wshape
        return ((STEM_IM2COL or ops.IGEMM) and TCGEN05["conv1x1"] and ci <= 4 and kh * kw > 1 and
                self.groups == 1 and self.dilation == [1, 1] and self.stride[0] == self.stride[1] and
                self.padding[0] == self.padding[1] and kh == kw and co % 8 == 0 and x.is_cuda and
                x.dtype == torch.bfloat16 and x.dim() == 4 and
                x.is_contiguous(memory_format=torch.channels_last) and
                kh * (x.shape[3] + 2 * self.padding[0]) * ci * 2 <= 48 * 1024)

    def _stem_weight_matrix(self, w: torch.Tensor, kpad: int) -> torch.Tensor:
        co, ci, kh, kw = self.wshape
        wm = getattr(self, "_wm", None)
        if wm is None or wm.device != w.device or wm.shape[1] != kpad:
            wm = torch.zeros((co, kpad), dtype=w.dtype, device=w.device)
            if not torch.cuda.is_current_stream_capturing():
                self._wm = wm          # padded columns stay zero; only [:, :k] is rewritten
        wm[:, :kh * kw * ci].copy_(w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci))
        return wm

    def _is_conv_tc(self, x: torch.Tensor, w: torch.Tensor) -> bool:
        return CONV3X3_TCGEN05 and TCGEN05["conv1x1"] and ops.conv_tn_supported(
            x, w, self.stride, self.padding, self.dilation, self.groups)

    def _conv_fwd(self, x, w, bias=None, col_sums=None, relu=False, bn_finalize=None, ctx=None):
        """Returns (y, stats_fused).  ``ctx``: the op's saved-tensor dict (the stem path keeps its
        patch matrix there for the weight gradient)."""
        fold = PENDING_FOLD.pop(id(x), None) if PENDING_FOLD else None
        if fold is not None and self._is_gemm(x):
            n, ci, h, wd = x.shape
            co = self.wshape[0]
            d = ops.gemm_tn_bnfold(fold, w.reshape(co, ci), col_sums=col_sums, bias=bias, relu=relu)
            return d.view(n, h, wd, co).permute(0, 3, 1, 2), col_sums is not None
        if fold is not None:
            fold.materialize()              # consumer cannot fold after all: plain application
        if self._is_ig(x, col_sums is not None):
            y = ops.conv_fprop(x, w, self.stride[0], self.padding, col_sums=col_sums, bias=bias,
                               relu=relu, bn_finalize=bn_finalize)
            if y is not None:
                return y, col_sums is not None
            assert bn_finalize is None, "igemm declined a shape igemm_conv_supported accepted"
        if self._is_gemm(x):
            n, ci, h, wd = x.shape
            co = self.wshape[0]
            a = x.permute(0, 2, 3, 1).reshape(n * h * wd, ci)
            d = ops.gemm_tn(a, w.reshape(co, ci), col_sums=col_sums, bias=bias, relu=relu,
                            bn_finalize=bn_finalize)
            return d.view(n, h, wd, co).permute(0, 3, 1, 2), col_sums is not None
        if self._is_stem_gemm(x):
            n, ci, h, wd = x.shape
            co, _, kh, kw = self.wshape
            a = ops.im2col_rows(x, kh, kw, self.stride[0], self.padding[0])
            if ctx is not None and (WGRAD_TCGEN05 or ops.IGEMM):
                ctx["patch"] = a
            oh = (h + 2 * self.padding[0] - kh) // self.stride[0] + 1
            ow = (wd + 2 * self.padding[1] - kw) // self.stride[1] + 1
            wm = self._stem_weight_matrix(w, a.shape[1])
            d = ops.gemm_ig(a, wm, col_sums=col_sums, bias=bias, relu=relu,
                            bn_finalize=bn_finalize) if ops.igemm_ok2d(a, co) else None
            if d is None:
                d = ops.gemm_tn(a, wm, col_sums=col_sums, bias=bias, relu=relu, bn_finalize=bn_finalize)
            return d.view(n, oh, ow, co).permute(0, 3, 1, 2), col_sums is not None
        if DWCONV and ops.dwconv_supported(x, self.wshape, self.stride, self.padding, self.dilation,
                                           self.groups) and bias is None and not relu:
            return ops.dwconv_fwd(x, w, self.stride[0], self.padding[0]), False
        if self._is_conv_tc(x, w):
            y = ops.conv_tn(x, w, col_sums=col_sums, bias=bias, relu=relu, stride=self.stride[0],
                            padding=self.padding)
            if y is not None:
                return y, col_sums is not None
        y = aten.convolution(x, w, None, self.stride, self.padding, self.dilation, False, [0, 0],
                             self.groups)
        return y, False

    def _wgrad_tc(self, dy, x, w, patch=None):
        """Weight gradient on the tcgen05 kernels (opt-in paths), or None: the caller then uses the
        library.  Returned in [Cout, Cin, KH, KW] logical shape, fp32."""
        if not (dy.is_cuda and dy.is_contiguous(memory_format=torch.channels_last)):
            return None
        if patch is not None:
            # stem: dW[Cout, kh*kw*Cin (padded)] = dY^T @ patch matrix, then drop the padding
            co, ci, kh, kw = self.wshape
            n, _, oh, ow = dy.shape
            dy2 = dy.permute(0, 2, 3, 1).reshape(n * oh * ow, co)
            dwm = torch.empty((co, patch.shape[1]), dtype=torch.float32, device=dy.device)
            if not ops.wgrad_gemm_into(dy2, patch, dwm, False):
                dwm = ops.gemm_nt_mn(dy2, patch)
            return dwm[:, :kh * kw * ci].reshape(co, kh, kw, ci).permute(0, 3, 1, 2)
        if not WGRAD_TCGEN05:
            return None
        if self._is_conv_tc(x, w):
            # 3x3 / strided convolutions: shifted-window MN-major kernel (may decline -> None)
            return ops.conv_wgrad(dy, x, self.wshape, self.stride[0], self.padding)
        if self._is_gemm(x):
            n, ci, h, wd = x.shape
            co = self.wshape[0]
            return ops.gemm_nt_mn(dy.permute(0, 2, 3, 1).reshape(n * h * wd, co),
                                  x.permute(0, 2, 3, 1).reshape(n * h * wd, ci)).view(co, ci, 1, 1)
        return None

    def _conv_bwd_ig(self, dy, x, w, need_dx, gw, accumulate, side):
        """Data + weight gradient on the tcgen05 kernels: dgrad reads the forward filter as an
        MN-major operand (no flip / transpose), the weight gradient lands in the flat gradient
        buffer in one launch (on the side stream when one is configured)."""
        args = (None, self.stride, self.padding, self.dilation, False, [0, 0], self.groups)
        dx = None
        if need_dx:
            dx = ops.conv_dgrad(dy, w, x.shape, self.stride[0], self.padding)
            if dx is None:
                dx = aten.convolution_backward(dy, x, w, *args, [True, False, False])[0]

        def wgrad():
            if not ops.conv_wgrad_into(dy, x, gw, self.stride[0], self.padding, accumulate):
                _write(gw, aten.convolution_backward(dy, x, w, *args, [False, True, False])[1], accumulate)

        if side is None:
            wgrad()
            return dx
        main = torch.cuda.current_stream(dy.device)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            wgrad()
        WGRAD_OVERLAP["keep"].append((dy, x))
        return dx

    def _conv_bwd(self, dy, x, w, need_dx: bool, gw: torch.Tensor, accumulate: bool, patch=None):
        """Data gradient (returned) and weight gradient (written into ``gw``).  With a side
        stream configured (``WGRAD_OVERLAP``) the weight-gradient kernel — which nothing
        needs before the optimizer — runs concurrently with the rest of backward: at the
        benchmark's batch size every kernel fills only a fraction of the 148 SMs."""
        if DWCONV and ops.dwconv_supported(x, self.wshape, self.stride, self.padding, self.dilation,
                                           self.groups) and dy.is_contiguous(memory_format=torch.channels_last):
            dx, dw = ops.dwconv_bwd(dy, x, w, self.stride[0], self.padding[0], need_dx)
            _write(gw, dw, accumulate)
            return dx
        side = WGRAD_OVERLAP["stream"] if dy.is_cuda else None
        if patch is None and self._is_ig(x) and dy.is_cuda and dy.dtype == torch.bfloat16 and \
                dy.is_contiguous(memory_format=torch.channels_last):
            return self._conv_bwd_ig(dy, x, w, need_dx, gw, accumulate, side)
        args = (None, self.stride, self.padding, self.dilation, False, [0, 0], self.groups)
        dx_tc = None
        if need_dx and CONV3X3_DGRAD and self._is_conv_tc(x, w) and self.stride == [1, 1] and \
                self.padding == [self.wshape[2] // 2, self.wshape[3] // 2] and \
                self.wshape[2] % 2 == 1 and self.wshape[3] % 2 == 1 and \
                dy.is_contiguous(memory_format=torch.channels_last):
            # data gradient of a stride-1 "same" convolution = the same convolution of dy with the
            # spatially flipped, channel-transposed filter: reuse the tcgen05 kernel
            wt = w.flip(2, 3).permute(1, 0, 2, 3).contiguous(memory_format=torch.channels_last)
            dx_tc = ops.conv_tn(dy, wt)
        elif need_dx and DGRAD1X1_TCGEN05 and self._is_gemm(x) and \
                dy.is_contiguous(memory_format=torch.channels_last):
            n, ci, h, wd = x.shape
            co = self.wshape[0]
            wt = w.reshape(co, ci).t().contiguous()                    # [Cin, Cout]: N x K, K-major
            a = dy.permute(0, 2, 3, 1).reshape(n * h * wd, co)
            if ops.gemm_tn_supported(a, wt):
                dx_tc = ops.gemm_tn(a, wt).view(n, h, wd, ci).permute(0, 3, 1, 2)
        lib_dx = need_dx and dx_tc is None
        if side is None:
            dw = self._wgrad_tc(dy, x, w, patch)
            if dw is None:
                dx, dw, _ = aten.convolution_backward(dy, x, w, *args, [lib_dx, True, False])
            else:
                dx = aten.convolution_backward(dy, x, w, *args, [True, False, False])[0] if lib_dx else None
            _write(gw, dw, accumulate)
            return dx_tc if dx_tc is not None else dx
        main = torch.cuda.current_stream(dy.device)
        ev = torch.cuda.Event()
        ev.record(main)
        dx = dx_tc if dx_tc is not None else (
            aten.convolution_backward(dy, x, w, *args, [True, False, False])[0] if lib_dx else None)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            dw = self._wgrad_tc(dy, x, w, patch)
            if dw is None:
                dw = aten.convolution_backward(dy, x, w, *args, [False, True, False])[1]
            _write(gw, dw, accumulate)
        # keep the operands alive until the streams are joined (executor.backward)
        WGRAD_OVERLAP["keep"].append((dy, x, dw))
        return dx


class ConvBnAct(Op, _ConvMixin):
    """y = act(BN(conv(x)) [+ res]) — one op forward, one op backward."""
    kind = "conv_bn_act"

    def __init__(self, name, inputs, output, conv: nn.Conv2d, bn: nn.BatchNorm2d, act: int,
                 has_res: bool, conv_name: str, bn_name: str):
        super().__init__(name, inputs, output)
        self._conv_cfg(conv)
        self.act, self.has_res = act, has_res
        self.eps, self.momentum = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
        self.channels = bn.num_features
        self.k_w, self.k_g, self.k_b = conv_name + ".weight", bn_name + ".weight", bn_name + ".bias"
        self.k_rm, self.k_rv = bn_name + ".running_mean", bn_name + ".running_var"
        self.params = [self.k_w, self.k_g, self.k_b]
        self._init = {self.k_w: conv.weight.detach().float(), self.k_g: bn.weight.detach().float(),
                      self.k_b: bn.bias.detach().float()}
        self._bufs = {self.k_rm: bn.running_mean.detach().float().clone(),
                      self.k_rv: bn.running_var.detach().float().clone()}
        self.ws: Optional[ops.BnWorkspace] = None

    def param_specs(self):
        from .store import ParamSpec
        return [ParamSpec(self.k_w, self.wshape, self._init[self.k_w], True, self.order),
                ParamSpec(self.k_g, (self.channels,), self._init[self.k_g], False, self.order),
                ParamSpec(self.k_b, (self.channels,), self._init[self.k_b], False, self.order)]

    def buffers(self):
        return self._bufs

    def to_device(self, device):
        if torch.device(device).type == "cuda":
            self.ws = ops.BnWorkspace(self.channels, device)

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        res = xs[1] if self.has_res else None
        ig = self._is_ig(x, True) or (ops.IGEMM and self._is_stem_gemm(x) and self.channels <= 2048)
        fuse_stats = training and self.ws is not None and \
            (ig or self._is_gemm(x) or self._is_stem_gemm(x) or self._is_conv_tc(x, P[self.k_w]))
        mom = self.momentum * BN_MOMENTUM_SCALE[0]
        stat_ready = None
        if fuse_stats and ig and ops.BN_APPLY_FINALIZE and not getattr(self, "defer_apply", False):
            # sums only: the BatchNorm apply kernel finalizes them (ops.bn_forward, stats_ready)
            c, fused = self._conv_fwd(x, P[self.k_w], col_sums=self.ws.sums, ctx=ctx)
        elif fuse_stats and (FUSE_BN_FINALIZE or ig) and not getattr(self, "defer_apply", False):
            # statistics AND their finalize ride in the GEMM epilogue (last CTA).  Measured
            # slower than a separate 4 us finalize launch (the last CTA gathers R*2*N values
            # with 128 threads and every CTA pays a fence + ticket), so this is off by default.
            stat_ready = torch.empty(2 * self.channels, dtype=torch.float32, device=x.device)
            fin = (self.ws.counter[64:] if ig else self.ws.counter[32:33], stat_ready, B[self.k_rm],
                   B[self.k_rv], self.eps, mom)
            c, fused = self._conv_fwd(x, P[self.k_w], col_sums=self.ws.sums, bn_finalize=fin, ctx=ctx)
        elif fuse_stats:
            c, fused = self._conv_fwd(x, P[self.k_w], col_sums=self.ws.sums, ctx=ctx)
        else:
            c, fused = self._conv_fwd(x, P[self.k_w])
        # 1 bit/element activation mask for the backward (read instead of y: 16x fewer bytes)
        mask = None
        if training and self.act != ACT_NONE and ACT_BITMASK_ON and ops._nhwc_ok(c):
            mask = torch.empty(c.numel() // 8, dtype=torch.uint8, device=c.device)
        if getattr(self, "defer_apply", False) and stat_ready is None:
            # the designated consumer (a 1x1-convolution GEMM) applies this BatchNorm while loading
            # its operand and writes y / mask as a side effect
            stat = ops.bn_statistics(c, B[self.k_rm], B[self.k_rv], self.eps, mom, self.ws,
                                     stats_ready=fused, training=training)
            y = torch.empty_like(c)
            PENDING_FOLD[id(y)] = ops.BnFold(c, res, stat, P[self.k_g], P[self.k_b], self.act, y, mask)
        else:
            y, stat = ops.bn_forward(c, res, P[self.k_g], P[self.k_b], B[self.k_rm], B[self.k_rv],
                                     self.eps, mom, self.act, self.ws, stats_ready=fused,
                                     training=training, stat_ready=stat_ready, mask_out=mask)
        if training:
            ctx["x"], ctx["c"], ctx["stat"], ctx["mask"] = x, c, stat, mask
            ctx["y"] = y if self.act != ACT_NONE else None
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy2 = dys[1] if len(dys) > 1 else None
        if len(dys) > 2:
            dys = [_sum_grads(dys[:-1]), dys[-1]]
            dy2 = dys[1]
        x_in = ctx["x"]
        if BN_BWD_FOLD and need[0] and self._is_gemm(x_in):
            co, ci = self.wshape[:2]
            wt = P[self.k_w].reshape(co, ci).t().contiguous()          # [Cin, Cout]: N x K, K-major
            if ops.bn_bwd_fold_supported(ctx["c"], dys[0], dy2, self.act, ctx.get("mask"), wt):
                dc, dz, dx2d = ops.bn_backward_dgrad_fold(
                    dys[0], dy2, ctx["c"], ctx["stat"], P[self.k_g], self.act,
                    self.has_res and need[1], G[self.k_g], G[self.k_b], accumulate, self.ws,
                    ctx.get("mask"), wt)
                n, _, h, wd = x_in.shape
                dx = dx2d.view(n, h, wd, ci).permute(0, 3, 1, 2)
                self._conv_bwd(dc, x_in, P[self.k_w], False, G[self.k_w], accumulate)   # weight gradient
                return [dx, dz] if self.has_res else [dx]
        dc, dz = ops.bn_backward(dys[0], dy2, ctx["y"], ctx["c"], ctx["stat"], P[self.k_g],
                                 self.act, self.has_res and need[1], G[self.k_g], G[self.k_b],
                                 accumulate, self.ws, mask=ctx.get("mask"))
        dx = self._conv_bwd(dc, x_in, P[self.k_w], need[0], G[self.k_w], accumulate,
                            patch=ctx.get("patch"))
        return [dx, dz] if self.has_res else [dx]


class ConvBiasAct(Op, _ConvMixin):
    """y = act(conv(x) + bias) (VGG)."""
    kind = "conv_bias_act"

    def __init__(self, name, inputs, output, conv: nn.Conv2d, act: int, conv_name: str):
        super().__init__(name, inputs, output)
        self._conv_cfg(conv)
        self.act = act
        self.has_bias = conv.bias is not None
        self.channels = conv.out_channels
        self.k_w, self.k_b = conv_name + ".weight", conv_name + ".bias"
        self.params = [self.k_w] + ([self.k_b] if self.has_bias else [])
        self._init = {self.k_w: conv.weight.detach().float()}
        if self.has_bias:
            self._init[self.k_b] = conv.bias.detach().float()
        self.ws: Optional[ops.BnWorkspace] = None

    def param_specs(self):
        from .store import ParamSpec
        s = [ParamSpec(self.k_w, self.wshape, self._init[self.k_w], True, self.order)]
        if self.has_bias:
            s.append(ParamSpec(self.k_b, (self.channels,), self._init[self.k_b], False, self.order))
        return s

    def to_device(self, device):
        if torch.device(device).type == "cuda":
            self.ws = ops.BnWorkspace(self.channels, device)

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        bias = P[self.k_b] if self.has_bias else None
        if self._is_gemm(x) and self.act in (ACT_NONE, ACT_RELU):
            y, _ = self._conv_fwd(x, P[self.k_w], bias=bias, relu=self.act == ACT_RELU)
        else:
            c, _ = self._conv_fwd(x, P[self.k_w])
            y = ops.bias_act(c, bias, self.act) if (bias is not None or self.act) else c
        if training:
            ctx["x"] = x
            ctx["y"] = y if self.act != ACT_NONE else None
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        if self.act != ACT_NONE or len(dys) > 1:
            if len(dys) > 2:
                dys = [_sum_grads(dys[:-1]), dys[-1]]
            dz = ops.act_bwd(dys[0], dys[1] if len(dys) > 1 else None, ctx["y"], self.act)
        else:
            dz = dys[0]
        if self.has_bias:
            ops.bias_grad_(dz, G[self.k_b], accumulate, self.ws)
        dx = self._conv_bwd(dz, ctx["x"], P[self.k_w], need[0], G[self.k_w], accumulate)
        return [dx]


class LinearAct(Op):
    """y = act(x @ W^T + b)."""
    kind = "linear_act"

    def __init__(self, name, inputs, output, lin: nn.Linear, act: int, lin_name: str):
        super().__init__(name, inputs, output)
        self.act = act
        self.in_f, self.out_f = lin.in_features, lin.out_features
        self.has_bias = lin.bias is not None
        self.k_w, self.k_b = lin_name + ".weight", lin_name + ".bias"
        self.params = [self.k_w] + ([self.k_b] if self.has_bias else [])
        self._init = {self.k_w: lin.weight.detach().float()}
        if self.has_bias:
            self._init[self.k_b] = lin.bias.detach().float()

    def param_specs(self):
        from .store import ParamSpec
        s = [ParamSpec(self.k_w, (self.out_f, self.in_f), self._init[self.k_w], False, self.order)]
        if self.has_bias:
            s.append(ParamSpec(self.k_b, (self.out_f,), self._init[self.k_b], False, self.order))
        return s

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        w = P[self.k_w]
        b = P[self.k_b] if self.has_bias else None
        y = None
        if (TCGEN05["linear"] and self.act in (ACT_NONE, ACT_RELU) and x.dim() == 2 and
                ops.igemm_ok2d(x, self.out_f) and w.dtype == torch.bfloat16):
            y = ops.gemm_ig(x, w, bias=b, relu=self.act == ACT_RELU)
        if y is not None:
            pass
        elif (TCGEN05["linear"] and self.act in (ACT_NONE, ACT_RELU) and x.dim() == 2 and
                ops.gemm_tn_supported(x, w) and self.out_f % 8 == 0):
            y = ops.gemm_tn(x, w, bias=b, relu=self.act == ACT_RELU)
        else:
            y = torch.addmm(b, x, w.t()) if b is not None else x @ w.t()
            if self.act != ACT_NONE:
                y = torch.relu(y) if self.act == ACT_RELU else torch.clamp(y, 0, 6)
        if training:
            ctx["x"] = x
            ctx["y"] = y if self.act != ACT_NONE else None
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        if self.act != ACT_NONE:
            dy = ops.act_bwd(dy, None, ctx["y"], self.act)
        x = ctx["x"]
        w = P[self.k_w]
        native = (TCGEN05["linear"] and dy.dim() == 2 and x.dim() == 2 and dy.is_contiguous() and
                  ops.igemm_ok2d(dy, self.in_f) and ops.igemm_ok2d(x, self.out_f) and
                  w.dtype == torch.bfloat16)
        # dW[out, in] = dy^T x: both operands MN-major, written straight into the gradient buffer
        if not (native and ops.wgrad_gemm_into(dy, x, G[self.k_w], accumulate)):
            _write(G[self.k_w], dy.t() @ x, accumulate)
        if self.has_bias:
            if native:
                if getattr(self, "_bias_ws", None) is None or self._bias_ws.sums.device != dy.device:
                    self._bias_ws = ops.BnWorkspace(self.out_f, dy.device)
                ops.bias_grad_(dy, G[self.k_b], accumulate, self._bias_ws)
            else:
                _write(G[self.k_b], dy.float().sum(0).to(dy.dtype), accumulate)
        if not need[0]:
            return [None]
        dx = ops.gemm_ig(dy, w, b_mn=True) if native else None     # dx = dy @ W: W is the MN-major B
        return [dx if dx is not None else dy @ w]


class MaxPool(Op):
    kind = "maxpool"

    def __init__(self, name, inputs, output, m: nn.MaxPool2d):
        super().__init__(name, inputs, output)
        self.k, self.s = _pair(m.kernel_size), _pair(m.stride if m.stride is not None else m.kernel_size)
        self.p, self.d, self.ceil = _pair(m.padding), _pair(m.dilation), bool(m.ceil_mode)

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        if ops.maxpool_supported(x, self.k, self.s, self.p, self.d, self.ceil):
            y, idx = ops.maxpool_fwd(x, self.k[0], self.s[0], self.p[0], want_idx=training)
            if training:
                ctx["native"], ctx["idx"], ctx["shape"] = True, idx, tuple(x.shape)
            return y
        y, idx = aten.max_pool2d_with_indices(x, self.k, self.s, self.p, self.d, self.ceil)
        if training:
            ctx["native"], ctx["x"], ctx["idx"] = False, x, idx
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        if not need[0]:
            return [None]
        if ctx["native"]:
            return [ops.maxpool_bwd(dy, ctx["idx"], ctx["shape"], self.k[0], self.s[0], self.p[0])]
        return [aten.max_pool2d_with_indices_backward(dy, ctx["x"], self.k, self.s, self.p, self.d,
                                                      self.ceil, ctx["idx"])]


class AvgPool(Op):
    """AvgPool2d / AdaptiveAvgPool2d (general case through aten)."""
    kind = "avgpool"

    def __init__(self, name, inputs, output, m: nn.Module):
        super().__init__(name, inputs, output)
        self.adaptive = isinstance(m, nn.AdaptiveAvgPool2d)
        if self.adaptive:
            self.out = _pair(m.output_size)
        else:
            self.k = _pair(m.kernel_size)
            self.s = _pair(m.stride if m.stride is not None else m.kernel_size)
            self.p = _pair(m.padding)
            self.ceil, self.cip = bool(m.ceil_mode), bool(m.count_include_pad)

    def _identity(self, x) -> bool:
        if self.adaptive:
            return list(x.shape[2:]) == self.out
        return self.k == [1, 1] and self.s == [1, 1] and self.p == [0, 0]

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        if self._identity(x):
            ctx["id"] = True
            return x
        ctx["id"] = False
        ctx["native"] = False
        if ops.avgpool_supported(x):
            if self.adaptive:
                y = ops.avgpool_fwd(x, self.out, 0, 0, 0, True)
            elif self.k[0] == self.k[1] and self.s[0] == self.s[1] and self.p[0] == self.p[1] and not self.ceil:
                oh = (x.shape[2] + 2 * self.p[0] - self.k[0]) // self.s[0] + 1
                ow = (x.shape[3] + 2 * self.p[1] - self.k[1]) // self.s[1] + 1
                y = ops.avgpool_fwd(x, (oh, ow), self.k[0], self.s[0], self.p[0], False, self.cip)
            else:
                y = None
            if y is not None:
                if training:
                    ctx["native"], ctx["shape"] = True, tuple(x.shape)
                return y
        if training:
            ctx["x"] = x
        if self.adaptive:
            return aten._adaptive_avg_pool2d(x, self.out)
        return aten.avg_pool2d(x, self.k, self.s, self.p, self.ceil, self.cip, None)

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        if not need[0]:
            return [None]
        if ctx["id"]:
            return [dy]
        if ctx.get("native"):
            if self.adaptive:
                dx = ops.avgpool_bwd(dy, ctx["shape"], 0, 0, 0, True)
            else:
                dx = ops.avgpool_bwd(dy, ctx["shape"], self.k[0], self.s[0], self.p[0], False, self.cip)
            assert dx is not None
            return [dx]
        if self.adaptive:
            return [aten._adaptive_avg_pool2d_backward(dy, ctx["x"])]
        return [aten.avg_pool2d_backward(dy, ctx["x"], self.k, self.s, self.p, self.ceil, self.cip,
                                         None)]


class GlobalAvgPoolFlatten(Op):
    """avg-pool to 1x1 + flatten, fused ([N,C,H,W] -> [N,C])."""
    kind = "gap_flatten"

    def forward(self, P, B, xs, ctx, training):
        ctx["shape"] = tuple(xs[0].shape)
        return ops.global_avg_pool(xs[0])

    def backward(self, P, G, dys, ctx, accumulate, need):
        if not need[0]:
            return [None]
        return [ops.global_avg_pool_bwd(_sum_grads(dys), ctx["shape"])]


class FlattenOp(Op):
    kind = "flatten"

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        ctx["shape"] = tuple(x.shape)
        ctx["cl"] = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and \
            not x.is_contiguous()
        return x.reshape(x.shape[0], -1)

    def backward(self, P, G, dys, ctx, accumulate, need):
        if not need[0]:
            return [None]
        dx = _sum_grads(dys).reshape(ctx["shape"])
        if ctx["cl"]:
            dx = dx.contiguous(memory_format=torch.channels_last)
        return [dx]


class DropoutOp(Op):
    kind = "dropout"

    def __init__(self, name, inputs, output, m: nn.Dropout):
        super().__init__(name, inputs, output)
        self.p = float(m.p)

    def forward(self, P, B, xs, ctx, training):
        if not training or self.p == 0.0:
            ctx["mask"] = None
            return xs[0]
        nat = ops.dropout_fwd(xs[0], self.p) if xs[0].is_cuda else None
        if nat is not None:
            ctx["mask"], ctx["native"] = nat[1], True
            return nat[0]
        y, mask = aten.native_dropout(xs[0], self.p, True)
        ctx["mask"], ctx["native"] = mask, False
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        if ctx.get("mask") is None:
            return [dy]
        if ctx.get("native"):
            return [ops.dropout_bwd(dy, ctx["mask"], self.p)]
        return [aten.native_dropout_backward(dy, ctx["mask"], 1.0 / (1.0 - self.p))]


class AddAct(Op):
    """y = act(a + b) for joins that are not preceded by a BatchNorm."""
    kind = "add_act"

    def __init__(self, name, inputs, output, act: int):
        super().__init__(name, inputs, output)
        self.act = act

    def forward(self, P, B, xs, ctx, training):
        y = xs[0]
        for x in xs[1:]:
            y = y + x
        if self.act == ACT_RELU:
            y = torch.relu(y)
        elif self.act == ACT_RELU6:
            y = torch.clamp(y, 0, 6)
        if training:
            ctx["y"] = y if self.act != ACT_NONE else None
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        if self.act != ACT_NONE or len(dys) > 1:
            if len(dys) > 2:
                dys = [_sum_grads(dys[:-1]), dys[-1]]
            dz = ops.act_bwd(dys[0], dys[1] if len(dys) > 1 else None, ctx["y"], self.act)
        else:
            dz = dys[0]
        return [dz if n else None for n in need]


class BnAct(Op):
    """y = act(BN(x)) for a BatchNorm that does not directly follow a convolution
    (pre-activation blocks: DenseNet, NASNet).  Same kernels as ``ConvBnAct``."""
    kind = "bn_act"

    def __init__(self, name, inputs, output, bn: nn.BatchNorm2d, act: int, bn_name: str):
        super().__init__(name, inputs, output)
        self.act = act
        self.eps, self.momentum = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
        self.channels = bn.num_features
        self.k_g, self.k_b = bn_name + ".weight", bn_name + ".bias"
        self.k_rm, self.k_rv = bn_name + ".running_mean", bn_name + ".running_var"
        self.params = [self.k_g, self.k_b]
        self._init = {self.k_g: bn.weight.detach().float(), self.k_b: bn.bias.detach().float()}
        self._bufs = {self.k_rm: bn.running_mean.detach().float().clone(),
                      self.k_rv: bn.running_var.detach().float().clone()}
        self.ws: Optional[ops.BnWorkspace] = None

    def param_specs(self):
        from .store import ParamSpec
        return [ParamSpec(self.k_g, (self.channels,), self._init[self.k_g], False, self.order),
                ParamSpec(self.k_b, (self.channels,), self._init[self.k_b], False, self.order)]

    def buffers(self):
        return self._bufs

    def to_device(self, device):
        if torch.device(device).type == "cuda":
            self.ws = ops.BnWorkspace(self.channels, device)

    def forward(self, P, B, xs, ctx, training):
        x = xs[0]
        mom = self.momentum * BN_MOMENTUM_SCALE[0]
        mask = None
        if training and self.act != ACT_NONE and ACT_BITMASK_ON and ops._nhwc_ok(x):
            mask = torch.empty(x.numel() // 8, dtype=torch.uint8, device=x.device)
        y, stat = ops.bn_forward(x, None, P[self.k_g], P[self.k_b], B[self.k_rm], B[self.k_rv],
                                 self.eps, mom, self.act, self.ws, training=training, mask_out=mask)
        if training:
            ctx["x"], ctx["stat"], ctx["mask"] = x, stat, mask
            ctx["y"] = y if self.act != ACT_NONE else None
        return y

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy2 = dys[1] if len(dys) > 1 else None
        if len(dys) > 2:
            dys = [_sum_grads(dys[:-1]), dys[-1]]
            dy2 = dys[1]
        dx, _ = ops.bn_backward(dys[0], dy2, ctx["y"], ctx["x"], ctx["stat"], P[self.k_g], self.act,
                                False, G[self.k_g], G[self.k_b], accumulate, self.ws,
                                mask=ctx.get("mask"))
        return [dx]


class ConcatOp(Op):
    """Channel concatenation (DenseNet / Inception / SqueezeNet / NASNet joins)."""
    kind = "concat"

    def __init__(self, name, inputs, output, dim: int = 1):
        super().__init__(name, inputs, output)
        self.dim = dim

    def forward(self, P, B, xs, ctx, training):
        if training:
            ctx["sizes"] = [x.shape[self.dim] for x in xs]
            ctx["cl"] = xs[0].dim() == 4 and xs[0].is_contiguous(memory_format=torch.channels_last)
        return torch.cat(list(xs), dim=self.dim)

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        parts = torch.split(dy, ctx["sizes"], dim=self.dim)
        fmt = torch.channels_last if ctx["cl"] else torch.contiguous_format
        return [p.contiguous(memory_format=fmt) if n else None for p, n in zip(parts, need)]


class ShiftCropOp(Op):
    """One-pixel shift (drop the first row / column, zero-pad at the end): second path of
    NASNet's factorised reduction."""
    kind = "shift_crop"

    def forward(self, P, B, xs, ctx, training):
        return torch.nn.functional.pad(xs[0][:, :, 1:, 1:], (0, 1, 0, 1))

    def backward(self, P, G, dys, ctx, accumulate, need):
        dy = _sum_grads(dys)
        return [torch.nn.functional.pad(dy[:, :, :-1, :-1], (1, 0, 1, 0))]


class ModuleOp(Op):
    """Fallback: any ``nn.Module`` runs under a local autograd tape, with its parameters
    aliased to views of the flat store.  Keeps the engine complete for arbitrary layers
    (LSTM/attention stages of GNMT, exotic pooling...) at the cost of no fusion."""
    kind = "module"

    def __init__(self, name, inputs, output, module: nn.Module, prefix: str):
        super().__init__(name, inputs, output)
        self.module = module
        self.prefix = prefix
        self._pnames = [n for n, _ in module.named_parameters()]
        self.params = [f"{prefix}.{n}" for n in self._pnames]
        self._bnames = [n for n, _ in module.named_buffers()]

    def param_specs(self):
        from .store import ParamSpec
        return [ParamSpec(f"{self.prefix}.{n}", tuple(p.shape), p.detach().float(), False, self.order)
                for n, p in self.module.named_parameters()]

    def buffers(self):
        return {f"{self.prefix}.{n}": b.detach().clone() for n, b in self.module.named_buffers()}

    def forward(self, P, B, xs, ctx, training):
        self.module.train(training)
        pb = {n: P[f"{self.prefix}.{n}"] for n in self._pnames}
        pb.update({n: B[f"{self.prefix}.{n}"] for n in self._bnames})
        if not training:
            with torch.no_grad():
                return torch.func.functional_call(self.module, pb, tuple(xs))
        with torch.enable_grad():
            # CPU: dense inputs — oneDNN's backward of a strided 1x1 convolution corrupts the heap when
            # handed a non-contiguous activation (hit by a torchgpipe-style Residual with a downsample)
            ins = [(x.detach().contiguous() if not x.is_cuda else x.detach()).requires_grad_(x.is_floating_point())
                   for x in xs]
            leafs = {n: P[f"{self.prefix}.{n}"].detach().requires_grad_(True) for n in self._pnames}
            call = dict(leafs)
            call.update({n: B[f"{self.prefix}.{n}"] for n in self._bnames})
            y = torch.func.functional_call(self.module, call, tuple(ins))
        ctx["ins"], ctx["leafs"], ctx["y"] = ins, leafs, y
        return y.detach()

    def backward(self, P, G, dys, ctx, accumulate, need):
        ins, leafs, y = ctx["ins"], ctx["leafs"], ctx["y"]
        wrt = [i for i, n in zip(ins, need) if n and i.requires_grad] + list(leafs.values())
        grads = torch.autograd.grad(y, wrt, _sum_grads(dys), allow_unused=True) if wrt else []
        out: List[Optional[torch.Tensor]] = []
        gi = 0
        for i, n in zip(ins, need):
            if n and i.requires_grad:
                out.append(grads[gi])
                gi += 1
            else:
                out.append(None)
        for n in self._pnames:
            g = grads[gi]
            gi += 1
            dst = G[f"{self.prefix}.{n}"]
            if g is None:
                if not accumulate:
                    dst.zero_()
            else:
                _write(dst, g, accumulate)
        return out
