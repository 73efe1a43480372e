"""A PyTorch stand-in for the native extension, used on machines without a GPU to execute the
*device branches* of the Python glue (``ops/__init__.py``, ``engine/ops.py``): every function
has the signature, the argument checks and the in-place output semantics of its binding in
``ops/csrc/bindings.cpp`` and computes the result with plain fp32 PyTorch.  Together with
``FakeCuda`` tensors (CPU tensors that answer ``is_cuda == True``) it catches what a numerics
test on the GPU would otherwise be the first to see: wrong argument order, wrong views /
strides / dtypes handed to a kernel, wrong reshapes of what comes back."""
import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
CL = torch.channels_last


class FakeCuda(torch.Tensor):
    """CPU tensor that claims to live on a CUDA device (results of ops on it do, too)."""

    @property
    def is_cuda(self):
        return True


def fake_cuda(t: torch.Tensor) -> torch.Tensor:
    return t.as_subclass(FakeCuda)


def _plain(t):
    return None if t is None else t.as_subclass(torch.Tensor)


def _act_mask(y, act):
    if act == 1:
        return y > 0
    if act == 2:
        return (y > 0) & (y < 6)
    return torch.ones_like(y, dtype=torch.bool)


class FakeExt:
    def __init__(self):
        self.calls = []

    # void gemm_bf16_tn(A, B, D, col_sums?, bias?, relu, ticket?, stat?, rm?, rv?, eps, mom)
    def gemm_bf16_tn(self, A, B, D, col_sums, bias, relu, ticket, stat, rm, rv, eps, mom):
        self.calls.append("gemm_bf16_tn")
        A, B, Dp = _plain(A), _plain(B), _plain(D)
        assert A.dim() == 2 and B.dim() == 2 and D.dim() == 2 and A.dtype == BF16 and B.dtype == BF16 and D.dtype == BF16
        assert A.stride(1) == 1 and B.stride(1) == 1 and D.stride(1) == 1
        M, K = A.shape
        N = B.shape[0]
        assert B.shape[1] == K and tuple(D.shape) == (M, N)
        d = A.float() @ B.float().t()
        if bias is not None:
            d = d + _plain(bias).float()
        if relu:
            d = d.relu()
        Dp.copy_(d.to(BF16))
        if col_sums is not None:
            cs = _plain(col_sums).view(-1)
            df = Dp.float()
            cs[:N] += df.sum(0)
            cs[N:2 * N] += (df * df).sum(0)
        assert ticket is None, "fake: fused finalize not modelled"

    # void gemm_bf16_nt_mn(A[K, M], B[K, N], D[M, N] fp32 +=, splits, workspace?)
    def gemm_bf16_nt_mn(self, A, B, D, splits, workspace):
        self.calls.append("gemm_bf16_nt_mn")
        A, B, Dp = _plain(A), _plain(B), _plain(D)
        assert A.dim() == 2 and B.dim() == 2 and D.dim() == 2 and A.dtype == BF16 and B.dtype == BF16
        assert D.dtype == torch.float32 and A.stride(1) == 1 and B.stride(1) == 1 and D.stride(1) == 1
        K, M = A.shape
        N = B.shape[1]
        assert B.shape[0] == K and tuple(D.shape) == (M, N)
        Dp += A.float().t() @ B.float()

    # bool gemm_bf16_tn_bnfold(C, R?, B, D, Y, mask?, stat, gamma, beta, act, col_sums?, bias?, relu)
    def gemm_bf16_tn_bnfold(self, C, R, B, D, Y, mask, stat, gamma, beta, act, col_sums, bias, relu):
        self.calls.append("gemm_bf16_tn_bnfold")
        C, R, B, Dp, Yp = _plain(C), _plain(R), _plain(B), _plain(D), _plain(Y)
        assert all(t.dim() == 2 and t.dtype == BF16 and t.stride(1) == 1 for t in (C, B, Dp, Yp))
        M, K = C.shape
        N = B.shape[0]
        assert B.shape[1] == K and tuple(Dp.shape) == (M, N) and tuple(Yp.shape) == (M, K)
        assert Yp.stride(0) == C.stride(0)
        if R is not None:
            assert R.stride(0) == C.stride(0) and tuple(R.shape) == (M, K) and R.dtype == BF16
        st = _plain(stat)
        assert st.numel() >= 2 * K and st.dtype == torch.float32 and st.is_contiguous()
        if K % 64 or K > 2048 or N % 8:
            return False
        sc = _plain(gamma).float() * st[K:2 * K]
        sh = _plain(beta).float() - st[:K] * sc
        y = C.float() * sc + sh
        if R is not None:
            y = y + R.float()
        if act == 1:
            y = y.relu()
        elif act == 2:
            y = y.clamp(0, 6)
        yb = y.to(BF16)
        Yp.copy_(yb)
        if mask is not None:
            bits = _act_mask(yb.float(), act).view(M, K // 8, 8).to(torch.int32)
            packed = (bits << torch.arange(8, dtype=torch.int32)).sum(-1).to(torch.uint8)
            _plain(mask).view(-1)[:M * K // 8].copy_(packed.view(-1))
        self.gemm_bf16_tn(yb, B, Dp, col_sums, bias, relu, None, None, None, None, 0.0, 0.0)
        self.calls.pop()
        return True

    # bool conv_bf16_nhwc(x, w, y, pad_h, pad_w, stride, col_sums?, bias?, relu)
    def conv_bf16_nhwc(self, x, w, y, pad_h, pad_w, stride, col_sums, bias, relu):
        self.calls.append("conv_bf16_nhwc")
        x, w, yp = _plain(x), _plain(w), _plain(y)
        assert x.dim() == 4 and w.dim() == 4 and yp.dim() == 4 and x.dtype == BF16 and w.dtype == BF16 and yp.dtype == BF16
        assert x.is_contiguous(memory_format=CL) and w.is_contiguous(memory_format=CL) and yp.is_contiguous(memory_format=CL)
        NB, Cin, H, W = x.shape
        Cout, _, KH, KW = w.shape
        OH, OW = (H + 2 * pad_h - KH) // stride + 1, (W + 2 * pad_w - KW) // stride + 1
        assert w.shape[1] == Cin and tuple(yp.shape) == (NB, Cout, OH, OW)
        if Cin % 64 or Cout % 64 or stride not in (1, 2):
            return False
        d = F.conv2d(x.float(), w.float(), None, stride, (pad_h, pad_w))
        if bias is not None:
            d = d + _plain(bias).float().view(1, -1, 1, 1)
        if relu:
            d = d.relu()
        yp.copy_(d.to(BF16))
        if col_sums is not None:
            cs = _plain(col_sums).view(-1)
            df = yp.float()
            cs[:Cout] += df.sum((0, 2, 3))
            cs[Cout:2 * Cout] += (df * df).sum((0, 2, 3))
        return True

    # ---- persistent implicit-GEMM family (igemm_tcgen05.cu) + one-launch weight gradients
    def _fin(self, cs, N, M, ticket, stat, rm, rv, eps, mom):
        if ticket is None or stat is None:
            return
        assert _plain(ticket).numel() >= (N + 63) // 64 and int(_plain(ticket).abs().sum()) == 0, "finalize tickets"
        v = cs[:R * 2 * N].view(R, 2, N).sum(0)
        mean = v[0] / M
        var = (v[1] / M - mean * mean).clamp_min(0)
        st = _plain(stat).view(-1)
        st[:N] = mean
        st[N:2 * N] = torch.rsqrt(var + eps)
        if rm is not None:
            _plain(rm).add_(mom * (mean - _plain(rm)))
            _plain(rv).add_(mom * (var * (M / max(M - 1, 1)) - _plain(rv)))
        cs.zero_()

    # bool igemm_gemm(A, B, D, b_mn, col_sums?, bias?, relu, ticket?, stat?, rm?, rv?, eps, mom)
    def igemm_gemm(self, A, B, D, b_mn, col_sums, bias, relu, ticket, stat, rm, rv, eps, mom, sk_ws=None, sk_tickets=None):
        self.calls.append("igemm_gemm")
        A, B, Dp = _plain(A), _plain(B), _plain(D)
        assert A.dim() == 2 and B.dim() == 2 and Dp.dim() == 2 and A.dtype == BF16 and B.dtype == BF16 and Dp.dtype == BF16
        assert A.stride(1) == 1 and B.stride(1) == 1 and Dp.stride(1) == 1
        M, K = A.shape
        N = Dp.shape[1]
        assert Dp.shape[0] == M and tuple(B.shape) == ((K, N) if b_mn else (N, K))
        if K % 8 or N % 8:
            return -1
        d = A.float() @ (B.float() if b_mn else B.float().t())
        if bias is not None:
            d = d + _plain(bias).float()
        if relu:
            d = d.relu()
        Dp.copy_(d.to(BF16))
        if col_sums is not None:
            assert not b_mn and N <= 2048
            cs = _plain(col_sums).view(-1)
            df = Dp.float()
            cs[:N] += df.sum(0)
            cs[N:2 * N] += (df * df).sum(0)
            self._fin(cs, N, M, ticket, stat, rm, rv, eps, mom)
        return 1 if (col_sums is not None and ticket is not None and stat is not None) else 0

    # bool gemm2cta(A, B, D, bias?, relu)   (CTA-pair GEMM: same contract as igemm_gemm without stats)
    def gemm2cta(self, A, B, D, bias, relu):
        self.calls.append("gemm2cta")
        A, B, Dp = _plain(A), _plain(B), _plain(D)
        assert A.dim() == 2 and B.dim() == 2 and A.dtype == BF16 and B.dtype == BF16 and Dp.dtype == BF16
        assert A.shape[1] == B.shape[1] and tuple(Dp.shape) == (A.shape[0], B.shape[0])
        if A.shape[1] % 8 or A.stride(0) % 8 or B.stride(0) % 8 or Dp.stride(0) % 8:
            return False
        d = A.float() @ B.float().t()
        if bias is not None:
            d = d + _plain(bias).float()
        Dp.copy_((d.relu() if relu else d).to(BF16))
        return True

    def igemm_set_trace(self, buf):
        self.calls.append("igemm_set_trace")

    # bool igemm_conv_fprop(x, w, y, pad_h, pad_w, stride, col_sums?, bias?, relu, ticket?, stat?, rm?, rv?, eps, mom)
    def igemm_conv_fprop(self, x, w, y, pad_h, pad_w, stride, col_sums, bias, relu, ticket, stat, rm, rv, eps, mom,
                         sk_ws=None, sk_tickets=None):
        self.calls.append("igemm_conv_fprop")
        x, w, yp = _plain(x), _plain(w), _plain(y)
        assert x.dim() == 4 and w.dim() == 4 and yp.dim() == 4 and x.dtype == BF16 and w.dtype == BF16 and yp.dtype == BF16
        assert x.is_contiguous(memory_format=CL) and w.is_contiguous(memory_format=CL) and yp.is_contiguous(memory_format=CL)
        NB, Cin, H, W = x.shape
        Cout, _, KH, KW = w.shape
        OH, OW = (H + 2 * pad_h - KH) // stride + 1, (W + 2 * pad_w - KW) // stride + 1
        assert w.shape[1] == Cin and tuple(yp.shape) == (NB, Cout, OH, OW)
        if Cin % 8 or Cout % 8 or stride not in (1, 2) or KH * KW > 49:
            return -1
        d = F.conv2d(x.float(), w.float(), None, stride, (pad_h, pad_w))
        if bias is not None:
            d = d + _plain(bias).float().view(1, -1, 1, 1)
        if relu:
            d = d.relu()
        yp.copy_(d.to(BF16))
        if col_sums is not None:
            assert Cout <= 2048
            cs = _plain(col_sums).view(-1)
            df = yp.float()
            cs[:Cout] += df.sum((0, 2, 3))
            cs[Cout:2 * Cout] += (df * df).sum((0, 2, 3))
            self._fin(cs, Cout, NB * OH * OW, ticket, stat, rm, rv, eps, mom)
        return 1 if (col_sums is not None and ticket is not None and stat is not None) else 0

    # bool igemm_conv_dgrad(dy, w, dx, pad_h, pad_w, stride)
    def igemm_conv_dgrad(self, dy, w, dx, pad_h, pad_w, stride, sk_ws=None, sk_tickets=None):
        self.calls.append("igemm_conv_dgrad")
        dy, w, dxp = _plain(dy), _plain(w), _plain(dx)
        assert dy.dtype == BF16 and w.dtype == BF16 and dxp.dtype == BF16
        assert dy.is_contiguous(memory_format=CL) and w.is_contiguous(memory_format=CL) and dxp.is_contiguous(memory_format=CL)
        NB, Cin, H, W = dxp.shape
        Cout, _, KH, KW = w.shape
        OH, OW = (H + 2 * pad_h - KH) // stride + 1, (W + 2 * pad_w - KW) // stride + 1
        assert w.shape[1] == Cin and tuple(dy.shape) == (NB, Cout, OH, OW)
        if Cin % 8 or Cout % 8 or stride not in (1, 2):
            return False
        g = torch.ops.aten.convolution_backward(
            dy.float().contiguous(), torch.zeros(NB, Cin, H, W), w.float().contiguous(), None, [stride, stride],
            [pad_h, pad_w], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        dxp.copy_(g.to(BF16))
        return True

    def _wg_check(self, ws, need):
        ws = _plain(ws)
        assert ws.dtype == torch.float32 and ws.numel() >= 1

    # bool wgrad_gemm(A[K, M], B[K, N], ws, tickets, out[M, N], accumulate)
    def wgrad_gemm(self, A, B, ws, out, accumulate):
        self.calls.append("wgrad_gemm")
        A, B, op = _plain(A), _plain(B), _plain(out)
        assert A.dim() == 2 and B.dim() == 2 and op.dim() == 2 and A.dtype == BF16 and B.dtype == BF16
        assert A.stride(1) == 1 and B.stride(1) == 1 and op.stride(1) == 1
        K, M = A.shape
        N = B.shape[1]
        assert B.shape[0] == K and tuple(op.shape) == (M, N)
        self._wg_check(ws, M * op.stride(0))
        if M % 8 or N % 8:
            return False
        g = A.float().t() @ B.float()
        op.copy_((op.float() + g if accumulate else g).to(op.dtype))
        return True

    # bool wgrad_conv(dy, x, ws, tickets, out [Cout, Cin, KH, KW] KRSC memory, accumulate, pad_h, pad_w, stride)
    def wgrad_conv(self, dy, x, ws, out, accumulate, pad_h, pad_w, stride):
        self.calls.append("wgrad_conv")
        dy, x, op = _plain(dy), _plain(x), _plain(out)
        assert dy.dtype == BF16 and x.dtype == BF16 and op.dim() == 4
        assert dy.is_contiguous(memory_format=CL) and x.is_contiguous(memory_format=CL) and op.is_contiguous(memory_format=CL)
        NB, Cin, H, W = x.shape
        Cout, _, KH, KW = op.shape
        assert op.shape[1] == Cin and dy.shape[0] == NB and dy.shape[1] == Cout
        self._wg_check(ws, op.numel())
        if Cin % 8 or Cout % 8 or stride not in (1, 2):
            return False
        g = torch.ops.aten.convolution_backward(
            dy.float().contiguous(), x.float().contiguous(), torch.zeros(Cout, Cin, KH, KW), None, [stride, stride],
            [pad_h, pad_w], [1, 1], False, [0, 0], 1, [False, True, False])[1]
        op.copy_((op.float() + g if accumulate else g).to(op.dtype))
        return True

    # bool im2col_rows(x, a, KH, KW, stride, pad)
    def im2col_rows(self, x, a, KH, KW, stride, pad):
        self.calls.append("im2col_rows")
        x, ap = _plain(x), _plain(a)
        assert x.dim() == 4 and x.dtype == BF16 and ap.dtype == BF16 and ap.dim() == 2 and ap.is_contiguous()
        assert x.is_contiguous(memory_format=CL)
        N, C, H, W = x.shape
        OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
        assert ap.shape[0] == N * OH * OW
        cols = F.unfold(x.float(), (KH, KW), padding=pad, stride=stride)          # [N, C*KH*KW, L]
        cols = cols.view(N, C, KH, KW, OH * OW).permute(0, 4, 2, 3, 1).reshape(N * OH * OW, KH * KW * C)
        ap.zero_()
        ap[:, :KH * KW * C].copy_(cols.to(BF16))
        return True

    # bool conv_wgrad_bf16_nhwc(dy, x, dw[Cout, KH, KW, Cin] fp32 +=, pad_h, pad_w, stride)
    def conv_wgrad_bf16_nhwc(self, dy, x, dw, pad_h, pad_w, stride):
        self.calls.append("conv_wgrad_bf16_nhwc")
        dy, x, dwp = _plain(dy), _plain(x), _plain(dw)
        assert dy.dim() == 4 and x.dim() == 4 and dwp.dim() == 4 and dy.dtype == BF16 and x.dtype == BF16
        assert dwp.dtype == torch.float32 and dwp.is_contiguous()
        assert dy.is_contiguous(memory_format=CL) and x.is_contiguous(memory_format=CL)
        NB, Cin, H, W = x.shape
        Cout, KH, KW, _ = dwp.shape
        assert dwp.shape[3] == Cin and dy.shape[0] == NB and dy.shape[1] == Cout
        g = torch.ops.aten.convolution_backward(
            dy.float().contiguous(), x.float().contiguous(), torch.zeros(Cout, Cin, KH, KW), None, [stride, stride],
            [pad_h, pad_w], [1, 1], False, [0, 0], 1, [False, True, False])[1]
        dwp += g.permute(0, 2, 3, 1)
        return True

    # bool dwconv3x3(mode, a, b, out, H, W, stride, pad)
    def dwconv3x3(self, mode, a, b, out, H, W, stride, pad):
        self.calls.append("dwconv3x3:%d" % mode)
        a, b, op = _plain(a), _plain(b), _plain(out)
        assert a.dim() == 4 and a.dtype == BF16 and a.is_contiguous(memory_format=CL)
        N, C = a.shape[:2]
        if mode == 0:                                     # a = x, b = taps [9, C], out = y
            assert tuple(b.shape) == (9, C) and b.is_contiguous() and (a.shape[2], a.shape[3]) == (H, W)
            w = b.float().t().reshape(C, 1, 3, 3)
            op.copy_(F.conv2d(a.float(), w, None, stride, pad, groups=C).to(BF16))
        elif mode == 1:                                   # a = dy, b = taps, out = dx [N, C, H, W]
            assert tuple(b.shape) == (9, C) and tuple(op.shape) == (N, C, H, W)
            w = b.float().t().reshape(C, 1, 3, 3)
            dx = torch.ops.aten.convolution_backward(
                a.float().contiguous(), torch.zeros(N, C, H, W), w, None, [stride, stride], [pad, pad], [1, 1],
                False, [0, 0], C, [True, False, False])[0]
            op.copy_(dx.to(BF16))
        else:                                             # a = dy, b = x, out = dw [9, C] fp32 +=
            assert b.dtype == BF16 and b.is_contiguous(memory_format=CL) and tuple(b.shape) == (N, C, H, W)
            assert tuple(op.shape) == (9, C) and op.dtype == torch.float32
            dw = torch.ops.aten.convolution_backward(
                a.float().contiguous(), b.float().contiguous(), torch.zeros(C, 1, 3, 3), None, [stride, stride],
                [pad, pad], [1, 1], False, [0, 0], C, [False, True, False])[1]
            op += dw.reshape(C, 9).t()
        return True


R = 8            # DDL_BN_REPLICAS
ACT_BITMASK = 0x100


def _unpack_mask(mask, M, C):
    m = _plain(mask).view(-1)[:M * C // 8].view(M, C // 8, 1).to(torch.int32)
    return ((m >> torch.arange(8, dtype=torch.int32)) & 1).view(M, C).bool()


def _rows(t, M, C):
    """[N, C, H, W] channels-last (or [M, C]) -> [M, C] fp32 copy."""
    t = _plain(t)
    if t.dim() == 4:
        assert t.is_contiguous(memory_format=CL), "kernel expects NHWC memory"
        t = t.permute(0, 2, 3, 1)
    return t.reshape(M, C).float()


def _store_rows(dst, rows):
    d = _plain(dst)
    if d.dim() == 4:
        assert d.is_contiguous(memory_format=CL)
        d.permute(0, 2, 3, 1).copy_(rows.view(d.shape[0], d.shape[2], d.shape[3], d.shape[1]).to(d.dtype))
    else:
        d.copy_(rows.view(d.shape).to(d.dtype))


class FullFakeExt(FakeExt):
    """+ the kernels of the default path (BatchNorm, pooling, loss, optimizer), so that a whole
    model can run through the device branches of the engine on the CPU."""

    def __init__(self):
        super().__init__()
        self._pool_idx = {}

    def set_pdl(self, on):
        return False

    # single-pass cluster BatchNorm kernels (bn_cluster.cu): the stand-in always DECLINES (as the
    # kernels do when a channel slab does not fit a cluster's shared memory), which exercises the
    # two-kernel fall-through of the glue
    def bn_fwd_fused(self, *a):
        self.calls.append("bn_fwd_fused:declined")
        return False

    def bn_bwd_fused(self, *a):
        self.calls.append("bn_bwd_fused:declined")
        return False

    # bool avgpool(backward, in, out, k, s, p, adaptive, count_include_pad)
    def avgpool(self, backward, inp, out, k, s, p, adaptive, cip):
        self.calls.append("avgpool:%d" % int(backward))
        a, op = _plain(inp), _plain(out)
        assert a.dim() == 4 and op.dim() == 4 and a.dtype == BF16 and op.dtype == BF16
        assert a.is_contiguous(memory_format=CL) and op.is_contiguous(memory_format=CL)
        big = op if backward else a
        small = a if backward else op
        x = torch.zeros(big.shape, dtype=torch.float32) if backward else a.float()
        x = x.contiguous().requires_grad_(True)
        y = F.adaptive_avg_pool2d(x, tuple(small.shape[2:])) if adaptive else F.avg_pool2d(x, k, s, p, False, cip)
        assert tuple(y.shape) == tuple(small.shape)
        if backward:
            y.backward(a.float().contiguous())
            op.copy_(x.grad.to(BF16))
        else:
            op.copy_(y.detach().to(BF16))
        return True

    def dropout_fwd(self, x, y, mask, p, seed, salt, step):
        self.calls.append("dropout_fwd")
        xp, yp, mp = _plain(x), _plain(y), _plain(mask)
        g = torch.Generator().manual_seed((seed ^ salt ^ int(_plain(step).item() if step is not None else 0)) & 0x7FFFFFFF)
        keep = torch.rand(xp.numel(), generator=g) >= p
        yp.copy_((xp.float().reshape(-1) * keep / (1 - p)).reshape(xp.shape).to(BF16))
        bits = keep.view(-1, 8).to(torch.int32)
        mp.view(-1)[:xp.numel() // 8].copy_((bits << torch.arange(8, dtype=torch.int32)).sum(-1).to(torch.uint8))

    def dropout_bwd(self, dy, mask, dx, scale):
        self.calls.append("dropout_bwd")
        d, m = _plain(dy), _plain(mask).view(-1)[:_plain(dy).numel() // 8].view(-1, 1).to(torch.int32)
        keep = ((m >> torch.arange(8, dtype=torch.int32)) & 1).view(-1).bool()
        _plain(dx).copy_((d.float().reshape(-1) * keep * scale).reshape(d.shape).to(BF16))

    def bn_replicas(self):
        return R

    # -------------------------------------------------------------------------- optimizer / loss
    def sgd_flat(self, master, mom, grad, w_out, w_out2, hyper):
        self.calls.append("sgd_flat")
        master, mom, grad, w_out, w_out2, hyper = map(_plain, (master, mom, grad, w_out, w_out2, hyper))
        assert master.dtype == torch.float32 and mom.dtype == torch.float32 and hyper.numel() >= 4
        n = master.numel()
        assert mom.numel() == n and grad.numel() == n and w_out.numel() == n
        lr, mu, wd, gs = [float(v) for v in hyper.tolist()[:4]]
        g = grad.float() * gs + wd * master
        mom.mul_(mu).add_(g)
        master.add_(mom, alpha=-lr)
        w_out.copy_(master.to(w_out.dtype))
        if w_out2 is not None:
            w_out2.copy_(master.to(w_out2.dtype))

    def xent(self, logits, target, dlogits, row_stats, ticket, smoothing, grad_scale):
        self.calls.append("xent")
        logits, target = _plain(logits), _plain(target)
        assert logits.dim() == 2 and logits.is_contiguous() and target.dtype == torch.int64
        assert target.numel() == logits.shape[0] and _plain(row_stats).numel() >= logits.shape[0] * 3
        x = logits.float()
        n, c = x.shape
        logp = F.log_softmax(x, dim=1)
        nll = -logp.gather(1, target.view(-1, 1)).squeeze(1)
        loss = ((1.0 - smoothing) * nll + smoothing * (-logp.mean(dim=1))).mean()
        if dlogits is not None:
            onehot = torch.zeros_like(logp).scatter_(1, target.view(-1, 1), 1.0)
            _plain(dlogits).copy_(((logp.exp() - (1.0 - smoothing) * onehot - smoothing / c) * grad_scale).to(logits.dtype))
        greater = (x > x.gather(1, target.view(-1, 1))).sum(dim=1)
        return fake_cuda(torch.stack([loss, (greater == 0).sum().float(), (greater < 5).sum().float()]))

    # -------------------------------------------------------------------------- BatchNorm
    def _finalize(self, s, q, stat, rmean, rvar, M, C, eps, mom):
        mean = s / M
        var = (q / M - mean * mean).clamp_min(0)
        st = _plain(stat).view(-1)
        st[:C] = mean
        st[C:2 * C] = torch.rsqrt(var + eps)
        if rmean is not None:
            _plain(rmean).add_(mom * (mean - _plain(rmean)))
            _plain(rvar).add_(mom * (var * (M / max(M - 1, 1)) - _plain(rvar)))

    def bn_stats(self, x, sums, counter, stat, rmean, rvar, M, C, eps, mom):
        self.calls.append("bn_stats")
        assert _plain(sums).numel() >= R * 2 * C and float(_plain(sums).abs().sum()) == 0.0, "workspace not zero"
        xr = _rows(x, M, C)
        self._finalize(xr.sum(0), (xr * xr).sum(0), stat, rmean, rvar, M, C, eps, mom)

    def bn_finalize(self, sums, stat, rmean, rvar, M, C, eps, mom):
        self.calls.append("bn_finalize")
        sp = _plain(sums)
        v = sp[:R * 2 * C].view(R, 2, C).sum(0)
        self._finalize(v[0], v[1], stat, rmean, rvar, M, C, eps, mom)
        sp.zero_()                                     # the kernel leaves its workspace zeroed

    def bn_apply_sums(self, x, res, y, sums, counter, stat, rmean, rvar, gamma, beta, mask_out, M, C, eps, mom, act):
        """Apply-side finalize: statistics from the producer's replicated sums, which are left zeroed."""
        self.calls.append("bn_apply_sums")
        if (C // 8) % min(C // 8, 256):
            return False
        assert _plain(counter).numel() >= 64 and int(_plain(counter)[48:64].abs().sum()) == 0
        sp = _plain(sums)
        v = sp[:R * 2 * C].view(R, 2, C).sum(0)
        self._finalize(v[0], v[1], stat, rmean, rvar, M, C, eps, mom)
        sp.zero_()
        self.bn_apply(x, res, y, stat, gamma, beta, mask_out, M, C, act)
        self.calls.pop()
        return True

    def bn_apply(self, x, res, y, stat, gamma, beta, mask_out, M, C, act):
        self.calls.append("bn_apply")
        st = _plain(stat).view(-1)
        assert st.is_contiguous() and (mask_out is None or _plain(mask_out).numel() * 8 >= M * C)
        sc = _plain(gamma).float() * st[C:2 * C]
        v = _rows(x, M, C) * sc + (_plain(beta).float() - st[:C] * sc)
        if res is not None:
            v = v + _rows(res, M, C)
        if act == 1:
            v = v.relu()
        elif act == 2:
            v = v.clamp(0, 6)
        vb = v.to(BF16)
        _store_rows(y, vb)
        if mask_out is not None:
            bits = _act_mask(vb.float(), act).view(M, C // 8, 8).to(torch.int32)
            _plain(mask_out).view(-1)[:M * C // 8].copy_(
                (bits << torch.arange(8, dtype=torch.int32)).sum(-1).to(torch.uint8).view(-1))

    def _dz(self, dy, dy2, y, M, C, act_flags):
        d = _rows(dy, M, C)
        if dy2 is not None:
            d = d + _rows(dy2, M, C)
        act = act_flags & 0xff
        if act:
            if act_flags & ACT_BITMASK:
                d = d * _unpack_mask(y, M, C)
            else:
                d = d * _act_mask(_rows(y, M, C), act)
        return d

    def bn_bwd_reduce(self, dy, dy2, y, x, dz_out, stat, sums, counter, gamma, coef, dgamma, dbeta,
                      accumulate, M, C, act):
        self.calls.append("bn_bwd_reduce")
        st = _plain(stat).view(-1)
        d = self._dz(dy, dy2, y, M, C, act)
        if dz_out is not None:
            _store_rows(dz_out, d.to(BF16))
            d = d.to(BF16).float()                     # the kernel sums what it stored
        xhat = (_rows(x, M, C) - st[:C]) * st[C:2 * C]
        s1, s2 = d.sum(0), (d * xhat).sum(0)
        cf = _plain(coef)
        cf[:C] = _plain(gamma).float() * st[C:2 * C]
        cf[C:2 * C] = s1 / M
        cf[2 * C:3 * C] = s2 / M
        for out, val in ((dgamma, s2), (dbeta, s1)):
            o = _plain(out)
            o.copy_((o.float() + val if accumulate else val).to(o.dtype))

    # bool gemm_bf16_tn_bnbwd(dy, dy2?, c, mask?, bt, dx, dc, dz?, stat, coef, act)
    def gemm_bf16_tn_bnbwd(self, dy, dy2, c, mask, bt, dx, dc, dz, stat, coef, act):
        self.calls.append("gemm_bf16_tn_bnbwd")
        dy, dy2, c, bt, dxp, dcp, dzp = map(_plain, (dy, dy2, c, bt, dx, dc, dz))
        for t in (dy, c, bt, dxp, dcp):
            assert t.dim() == 2 and t.dtype == BF16 and t.stride(1) == 1
        M, K = c.shape
        N = bt.shape[0]
        assert bt.shape[1] == K and tuple(dy.shape) == (M, K) and tuple(dcp.shape) == (M, K) and tuple(dxp.shape) == (M, N)
        ld = c.stride(0)
        assert dy.stride(0) == ld and dcp.stride(0) == ld
        for o in (dy2, dzp):
            if o is not None:
                assert tuple(o.shape) == (M, K) and o.stride(0) == ld and o.stride(1) == 1 and o.dtype == BF16
        st, cf = _plain(stat).view(-1), _plain(coef).view(-1)
        assert st.numel() >= 2 * K and cf.numel() >= 3 * K and st.is_contiguous() and cf.is_contiguous()
        assert mask is None or _plain(mask).numel() * 8 >= M * K
        if K % 64 or K > 2048 or N % 8 or (act != 0 and mask is None):
            return False
        d = dy.float()
        if dy2 is not None:
            d = d + dy2.float()
        if act != 0:
            d = d * _unpack_mask(mask, M, K)
        if dzp is not None:
            dzp.copy_(d.to(BF16))
            d = dzp.float()
        xhat = (c.float() - st[:K]) * st[K:2 * K]
        dcv = (cf[:K] * (d - cf[K:2 * K] - xhat * cf[2 * K:3 * K])).to(BF16)
        dcp.copy_(dcv)
        dxp.copy_((dcv.float() @ bt.float().t()).to(BF16))
        return True

    def bn_bwd_apply(self, dz_in, dy, dy2, y, x, dx, stat, coef, M, C, act):
        self.calls.append("bn_bwd_apply")
        st, cf = _plain(stat).view(-1), _plain(coef)
        d = _rows(dz_in, M, C) if dz_in is not None else self._dz(dy, dy2, y, M, C, act)
        xhat = (_rows(x, M, C) - st[:C]) * st[C:2 * C]
        _store_rows(dx, cf[:C] * (d - cf[C:2 * C] - xhat * cf[2 * C:3 * C]))

    def act_bwd(self, dy, dy2, y, dz, act):
        self.calls.append("act_bwd")
        d = _plain(dy).float()
        if dy2 is not None:
            d = d + _plain(dy2).float()
        if act:
            d = d * _act_mask(_plain(y).float(), act)
        _plain(dz).copy_(d.to(BF16))

    def bias_act(self, x, bias, y, M, C, act):
        self.calls.append("bias_act")
        v = _rows(x, M, C)
        if bias is not None:
            v = v + _plain(bias).float()
        _store_rows(y, v.relu() if act == 1 else (v.clamp(0, 6) if act == 2 else v))

    def colsum(self, x, sums, counter, out, accumulate, M, C):
        self.calls.append("colsum")
        o = _plain(out)
        s = _rows(x, M, C).sum(0)
        o.copy_((o.float() + s if accumulate else s).to(o.dtype))

    # -------------------------------------------------------------------------- pooling
    def gap_fwd(self, x, y, N, HW, C):
        self.calls.append("gap_fwd")
        _plain(y).copy_(_rows(x, N * HW, C).view(N, HW, C).mean(1).to(BF16))

    def gap_bwd(self, dy, dx, N, HW, C):
        self.calls.append("gap_bwd")
        d = (_plain(dy).float().view(N, 1, C) / HW).expand(N, HW, C).reshape(N * HW, C)
        _store_rows(dx, d)

    def maxpool_fwd(self, x, y, idx, k, s, p):
        self.calls.append("maxpool_fwd")
        xp = _plain(x)
        assert xp.is_contiguous(memory_format=CL) and _plain(y).is_contiguous(memory_format=CL)
        v, i = F.max_pool2d(xp.float(), k, s, p, return_indices=True)
        _plain(y).copy_(v.to(BF16))
        if idx is not None:
            self._pool_idx[_plain(idx).data_ptr()] = i

    def maxpool_bwd(self, dy, idx, dx, k, s, p):
        self.calls.append("maxpool_bwd")
        i = self._pool_idx[_plain(idx).data_ptr()]
        dxp = _plain(dx)
        assert dxp.is_contiguous(memory_format=CL) and _plain(dy).is_contiguous(memory_format=CL)
        n, c, h, w = dxp.shape
        g = torch.zeros(n, c, h * w).scatter_add_(2, i.view(n, c, -1), _plain(dy).float().reshape(n, c, -1))
        dxp.copy_(g.view(n, c, h, w).to(BF16))
