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
