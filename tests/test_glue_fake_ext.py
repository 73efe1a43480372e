"""The device branches of the Python glue, executed on the CPU: tensors that claim ``is_cuda`` and
a PyTorch stand-in for the extension with the bindings' signatures and argument checks
(``tests/fake_ext.py``).  Catches wrong argument order / views / layouts in the paths that only
run on a GPU (most of them opt-in and still waiting for their first hardware run)."""
import os
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_ext import FakeExt, fake_cuda  # noqa: E402

from ddlbench_b200 import ops  # noqa: E402
from ddlbench_b200.engine import ops as E  # noqa: E402

BF16 = torch.bfloat16
CL = torch.channels_last


@pytest.fixture
def fake(monkeypatch):
    fe = FakeExt()
    monkeypatch.setattr(ops, "ext", lambda: fe)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    # what the glue allocates lives "on the device" as well
    for name in ("empty", "zeros", "empty_like", "zeros_like"):
        real = getattr(torch, name)
        monkeypatch.setattr(torch, name, (lambda real: lambda *a, **k: fake_cuda(real(*a, **k)))(real))
    monkeypatch.setitem(E.TCGEN05, "conv1x1", True)
    monkeypatch.setattr(ops, "IGEMM", False)        # the legacy (round-1) kernels; igemm: fake_ig below
    monkeypatch.setattr(ops, "_WG_WS", {})
    for name in ("CONV3X3_TCGEN05", "WGRAD_TCGEN05", "STEM_IM2COL", "CONV3X3_DGRAD", "DWCONV", "DGRAD1X1_TCGEN05"):
        monkeypatch.setattr(E, name, True)
    return fe


@pytest.fixture
def fake_ig(fake, monkeypatch):
    """Default configuration: the persistent implicit-GEMM family."""
    monkeypatch.setattr(ops, "IGEMM", True)
    return fake


def _t(*shape, cl=True):
    t = torch.randn(*shape).to(BF16)
    if cl and t.dim() == 4:
        t = t.contiguous(memory_format=CL)
    return fake_cuda(t)


def _conv_op(ci, co, k, s, p, groups=1):
    conv = nn.Conv2d(ci, co, k, s, p, groups=groups, bias=False)
    return E.ConvBnAct("op", ["x"], "y", conv, nn.BatchNorm2d(co), E.ACT_RELU, False, "conv", "bn")


def _reference(x, w, dy, s, p, groups=1):
    xf = x.float().as_subclass(torch.Tensor).contiguous().requires_grad_(True)
    wf = w.float().as_subclass(torch.Tensor).contiguous().requires_grad_(True)
    y = F.conv2d(xf, wf, None, s, p, groups=groups)
    y.backward(dy.float().as_subclass(torch.Tensor).contiguous())
    return y.detach(), xf.grad, wf.grad


def _close(a, b, tol=2e-2):
    a = a.as_subclass(torch.Tensor).float()
    b = b.as_subclass(torch.Tensor).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def _close_fro(a, b, tol):
    """Relative Frobenius error: robust against the few ReLU bits that flip under bf16 rounding."""
    a = a.as_subclass(torch.Tensor).float()
    b = b.as_subclass(torch.Tensor).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert ((a - b).norm() / b.norm()).item() <= tol, ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("ci,co,k,s,p,expect", [
    (64, 64, 3, 1, 1, ["conv_bf16_nhwc", "conv_bf16_nhwc", "conv_wgrad_bf16_nhwc"]),     # fwd, dgrad, wgrad
    (64, 128, 3, 2, 1, ["conv_bf16_nhwc", "conv_wgrad_bf16_nhwc"]),                     # strided: library dgrad
    (64, 128, 1, 2, 0, ["conv_bf16_nhwc", "conv_wgrad_bf16_nhwc"]),                     # downsample
    (64, 128, 1, 1, 0, ["gemm_bf16_tn", "gemm_bf16_tn", "gemm_bf16_nt_mn"]),            # fwd, dgrad (W^T), wgrad
])
def test_conv_paths(fake, ci, co, k, s, p, expect):
    torch.manual_seed(0)
    op = _conv_op(ci, co, k, s, p)
    x, w = _t(2, ci, 9, 9), _t(co, ci, k, k)
    sums = torch.zeros(8 * 2 * co)
    y, fused = op._conv_fwd(x, w, col_sums=sums, ctx={})
    dy = _t(*y.shape)
    gw = torch.zeros(co, ci, k, k)
    dx = op._conv_bwd(dy, x, w, True, gw, False)
    y_ref, dx_ref, dw_ref = _reference(x, w, dy, s, p)
    assert fused and fake.calls == expect
    _close(y, y_ref)
    _close(sums[:co], y.as_subclass(torch.Tensor).float().sum((0, 2, 3)), 1e-3)
    _close(dx, dx_ref)
    _close(gw, dw_ref)
    assert y.is_contiguous(memory_format=CL)


@pytest.mark.parametrize("ci,co,k,s,p", [(64, 64, 3, 1, 1), (64, 128, 3, 2, 1), (64, 128, 1, 2, 0),
                                          (64, 128, 1, 1, 0), (24, 40, 3, 1, 1), (16, 8, 5, 2, 2)])
def test_igemm_conv_paths(fake_ig, ci, co, k, s, p):
    """Default path: fprop + fused statistics/finalize, dgrad from the forward filter, one-launch
    weight gradient straight into the (bf16, KRSC) gradient view — for every stride / kernel size."""
    fake = fake_ig
    torch.manual_seed(0)
    op = _conv_op(ci, co, k, s, p)
    x, w = _t(2, ci, 9, 9), _t(co, ci, k, k)
    sums = torch.zeros(8 * 2 * co)
    stat = torch.zeros(2 * co)
    fin = (torch.zeros((co + 63) // 64, dtype=torch.int32), stat, torch.zeros(co), torch.ones(co), 1e-5, 0.1)
    y, fused = op._conv_fwd(x, w, col_sums=sums, bn_finalize=fin, ctx={})
    dy = _t(*y.shape)
    gw = fake_cuda(torch.zeros(co, k, k, ci).to(BF16)).permute(0, 3, 1, 2)     # KRSC view, as in the store
    dx = op._conv_bwd(dy, x, w, True, gw, False)
    y_ref, dx_ref, dw_ref = _reference(x, w, dy, s, p)
    assert fused and fake.calls == ["igemm_conv_fprop", "igemm_conv_dgrad",
                                    "wgrad_gemm" if (k == 1 and s == 1) else "wgrad_conv"]
    _close(y, y_ref)
    yf = y.as_subclass(torch.Tensor).float()
    _close(stat[:co], yf.mean((0, 2, 3)), 1e-2)
    assert float(sums.as_subclass(torch.Tensor).abs().sum()) == 0.0          # finalize re-zeroes the workspace
    _close(dx, dx_ref)
    _close(gw, dw_ref)
    # accumulate (micro-batches after the first)
    fake.calls.clear()
    op._conv_bwd(dy, x, w, False, gw, True)
    _close(gw, 2 * dw_ref, 4e-2)
    assert y.is_contiguous(memory_format=CL) and dx.is_contiguous(memory_format=CL)


def test_igemm_linear_paths(fake_ig, monkeypatch):
    fake = fake_ig
    monkeypatch.setitem(E.TCGEN05, "linear", True)
    torch.manual_seed(3)
    lin = nn.Linear(64, 24)
    op = E.LinearAct("fc", ["x"], "y", lin, E.ACT_RELU, "fc")
    P = {"fc.weight": fake_cuda(lin.weight.detach().to(BF16)), "fc.bias": fake_cuda(lin.bias.detach().to(BF16))}
    G = {"fc.weight": fake_cuda(torch.zeros(24, 64).to(BF16)), "fc.bias": fake_cuda(torch.zeros(24).to(BF16))}
    x = _t(16, 64)
    ctx = {}
    y = op.forward(P, {}, [x], ctx, True)
    dy = _t(16, 24)
    fe = fake
    # bias gradient goes through the colsum kernel, which the conv-only fake does not model
    monkeypatch.setattr(ops, "bias_grad_", lambda dz, out, acc, ws=None: out.copy_(
        dz.as_subclass(torch.Tensor).float().sum(0).to(out.dtype)))
    monkeypatch.setattr(ops, "act_bwd", lambda dy_, dy2, y_, act: fake_cuda(
        (dy_.as_subclass(torch.Tensor).float() * (y_.as_subclass(torch.Tensor) > 0)).to(BF16)))
    monkeypatch.setattr(ops, "BnWorkspace", lambda c, d: type("W", (), {"sums": torch.zeros(1)})())
    dx, = op.backward(P, G, [dy], ctx, False, [True])
    assert fe.calls == ["igemm_gemm", "wgrad_gemm", "igemm_gemm"]
    xr, wr = x.as_subclass(torch.Tensor).float(), P["fc.weight"].as_subclass(torch.Tensor).float()
    y_ref = (xr @ wr.t() + P["fc.bias"].as_subclass(torch.Tensor).float()).relu()
    _close(y, y_ref)
    dz = dy.as_subclass(torch.Tensor).float() * (y_ref > 0)
    _close(dx, dz @ wr, 3e-2)
    _close(G["fc.weight"], dz.t() @ xr, 3e-2)


def test_gemm_ig_dispatches_big_problems_to_the_pair_kernel(fake_ig, monkeypatch):
    """ops.gemm_ig: stats-free K-major GEMMs above GEMM_PAIR_MIN go to the CTA-pair kernel, everything
    else (small, MN-major B, statistics) stays on the single-CTA kernel."""
    fake = fake_ig
    monkeypatch.setattr(ops, "GEMM_PAIR_MIN", (64, 32, 32))
    a, w, b = _t(64, 32), _t(72, 32), _t(72)
    y = ops.gemm_ig(a, w, bias=b, relu=True)
    assert fake.calls == ["gemm2cta"]
    _close(y, (a.as_subclass(torch.Tensor).float() @ w.as_subclass(torch.Tensor).float().t() +
               b.as_subclass(torch.Tensor).float()).relu())
    fake.calls.clear()
    ops.gemm_ig(_t(64, 72), w, b_mn=True)                       # data gradient: B is MN-major
    ops.gemm_ig(_t(16, 32), w)                                  # below the size threshold
    assert fake.calls == ["igemm_gemm", "igemm_gemm"]
    fake.calls.clear()
    monkeypatch.setattr(ops, "GEMM_PAIR", False)
    ops.gemm_ig(a, w)
    assert fake.calls == ["igemm_gemm"]


def test_stem_im2col_paths(fake):
    torch.manual_seed(1)
    op = _conv_op(3, 64, 7, 2, 3)
    x, w = _t(2, 3, 32, 32), _t(64, 3, 7, 7)
    ctx = {}
    y, fused = op._conv_fwd(x, w, col_sums=torch.zeros(8 * 2 * 64), ctx=ctx)
    assert fake.calls == ["im2col_rows", "gemm_bf16_tn"] and "patch" in ctx and ctx["patch"].shape[1] == 152
    dy = _t(*y.shape)
    gw = torch.zeros(64, 3, 7, 7)
    dx = op._conv_bwd(dy, x, w, False, gw, False, patch=ctx["patch"])
    y_ref, _, dw_ref = _reference(x, w, dy, 2, 3)
    assert dx is None and fake.calls[-1] == "gemm_bf16_nt_mn"
    _close(y, y_ref)
    _close(gw, dw_ref)


@pytest.mark.parametrize("s", [1, 2])
def test_depthwise_paths(fake, s):
    torch.manual_seed(2)
    op = _conv_op(16, 16, 3, s, 1, groups=16)
    x, w = _t(2, 16, 10, 10), _t(16, 1, 3, 3)
    y, fused = op._conv_fwd(x, w)
    dy = _t(*y.shape)
    gw = torch.zeros(16, 1, 3, 3)
    dx = op._conv_bwd(dy, x, w, True, gw, False)
    y_ref, dx_ref, dw_ref = _reference(x, w, dy, s, 1, groups=16)
    assert not fused and fake.calls == ["dwconv3x3:0", "dwconv3x3:1", "dwconv3x3:2"]
    _close(y, y_ref)
    _close(dx, dx_ref)
    _close(gw, dw_ref)


@pytest.mark.parametrize("has_res,two_grads", [(False, False), (True, False), (True, True)])
def test_bn_backward_folded_into_dgrad_matches_unfused(monkeypatch, has_res, two_grads):
    """conv1x1 + BN (+ residual) + ReLU: backward with the element-wise BN-backward half folded into
    the data-gradient GEMM (DDLB_BN_BWD_FOLD) against the unfused device path, and both against
    autograd."""
    from fake_ext import FullFakeExt
    torch.manual_seed(4)
    n, ci, co, h = 2, 64, 128, 6
    results = []
    for fold in (False, True):
        with monkeypatch.context() as mp:
            fe = FullFakeExt()
            mp.setattr(ops, "ext", lambda: fe)
            mp.setitem(E.TCGEN05, "conv1x1", True)
            mp.setattr(E, "BN_BWD_FOLD", fold)
            mp.setattr(ops, "IGEMM", False)               # the legacy 1x1 GEMM path this fold belongs to
            mp.setattr(ops, "BN_CLUSTER_FUSED", False)
            for name in ("empty", "zeros", "empty_like", "zeros_like"):
                real = getattr(torch, name)
                mp.setattr(torch, name, (lambda real: lambda *a, **k: fake_cuda(real(*a, **k)))(real))
            g = torch.Generator().manual_seed(9)
            conv = nn.Conv2d(ci, co, 1, bias=False)
            op = E.ConvBnAct("op", ["x", "r"] if has_res else ["x"], "y", conv, nn.BatchNorm2d(co), E.ACT_RELU,
                             has_res, "conv", "bn")
            op.ws = ops.BnWorkspace(co, "cpu")
            rnd = lambda *s: fake_cuda(torch.randn(*s, generator=g).to(BF16))        # noqa: E731
            x = rnd(n, ci, h, h).contiguous(memory_format=CL)
            r = rnd(n, co, h, h).contiguous(memory_format=CL) if has_res else None
            w = rnd(co, ci, 1, 1).contiguous(memory_format=CL)
            gamma, beta = fake_cuda(torch.rand(co, generator=g) + 0.5), fake_cuda(torch.randn(co, generator=g) * 0.1)
            P = {"conv.weight": w, "bn.weight": gamma, "bn.bias": beta}
            B = {"bn.running_mean": fake_cuda(torch.zeros(co)), "bn.running_var": fake_cuda(torch.ones(co))}
            G = {"conv.weight": fake_cuda(torch.zeros(co, ci, 1, 1)), "bn.weight": fake_cuda(torch.zeros(co)),
                 "bn.bias": fake_cuda(torch.zeros(co))}
            ctx = {}
            y = op.forward(P, B, [x, r] if has_res else [x], ctx, True)
            dys = [rnd(*y.shape).contiguous(memory_format=CL) for _ in range(2 if two_grads else 1)]
            outs = op.backward(P, G, dys, ctx, False, [True, True] if has_res else [True])
            assert ("gemm_bf16_tn_bnbwd" in fe.calls) == fold and ("bn_bwd_apply" in fe.calls) != fold
            results.append((y, outs, G, (x, r, w, gamma, beta, dys)))
    (y0, o0, G0, inp), (y1, o1, G1, _) = results
    for a, b in zip(o0, o1):
        _close(a, b, 1e-2)
    for k in G0:
        _close(G0[k], G1[k], 1e-2)
    # and against autograd (fp32 on the bf16-rounded inputs; the device path rounds c, y and dc to bf16
    # and ReLU bits of elements next to zero may flip, hence the loose bound)
    x, r, w, gamma, beta, dys = [t if not isinstance(t, torch.Tensor) else t.as_subclass(torch.Tensor).float() for t in inp[:5]] + [inp[5]]
    xr, wr = x.contiguous().requires_grad_(True), w.contiguous().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = r.contiguous().requires_grad_(True) if has_res else None
    z = F.batch_norm(F.conv2d(xr, wr), None, None, gr, br, True, 0.1, 1e-5)
    yr = torch.relu(z + rr if has_res else z)
    yr.backward(sum(d.as_subclass(torch.Tensor).float() for d in dys))
    _close_fro(o1[0], xr.grad, 6e-2)
    if has_res:
        _close_fro(o1[1], rr.grad, 6e-2)
    _close_fro(G1["conv.weight"], wr.grad, 6e-2)
    _close_fro(G1["bn.weight"], gr.grad, 6e-2)
    _close_fro(G1["bn.bias"], br.grad, 6e-2)


@pytest.mark.parametrize("has_res", [False, True])
def test_bnfold_glue_writes_y_and_mask_in_place(fake, has_res):
    """The consumer GEMM must write y / mask through VIEWS of the tensors the producer allocated
    (a reshape that silently copies would leave them unwritten)."""
    torch.manual_seed(3)
    n, k, h, w_, co = 2, 64, 6, 6, 128
    c = _t(n, k, h, w_)
    res = _t(n, k, h, w_) if has_res else None
    mean, var = torch.randn(k) * 0.1, torch.rand(k) + 0.5
    stat = torch.stack([mean, (var + 1e-5).rsqrt()])
    gamma, beta = torch.rand(k) + 0.5, torch.randn(k) * 0.1
    y = torch.full_like(c, float("nan"))
    mask = torch.zeros(c.numel() // 8, dtype=torch.uint8)
    fold = ops.BnFold(c, res, stat, gamma, beta, 1, y, mask)
    b = _t(co, k)
    d = ops.gemm_tn_bnfold(fold, b)
    assert fake.calls == ["gemm_bf16_tn_bnfold"]
    cf = c.float().as_subclass(torch.Tensor)
    y_ref = (cf - mean.view(1, -1, 1, 1)) * stat[1].view(1, -1, 1, 1) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    if has_res:
        y_ref = y_ref + res.float().as_subclass(torch.Tensor)
    y_ref = y_ref.relu()
    yp = y.as_subclass(torch.Tensor).float()
    assert not torch.isnan(yp).any()
    _close(yp, y_ref)
    d_ref = yp.permute(0, 2, 3, 1).reshape(-1, k) @ b.float().as_subclass(torch.Tensor).t()
    _close(d, d_ref)
    # bit j of byte [pixel][channel // 8] <-> y[pixel, channel] > 0 (the layout bn_backward reads)
    bits = (yp.permute(0, 2, 3, 1).reshape(-1, k // 8, 8) > 0).to(torch.int32)
    packed = (bits << torch.arange(8, dtype=torch.int32)).sum(-1).to(torch.uint8)
    assert torch.equal(mask.view(-1, k // 8), packed)


def test_resnet50_through_device_branches_with_every_optin_switch(monkeypatch):
    """Whole ResNet-50 steps (forward, loss, backward, SGD) through the DEVICE branches of the
    engine.  (1) BN folding into the 1x1 consumers + tcgen05 1x1 weight gradients must reproduce
    the default device path (same rounding points: equal up to the last bit); (2) with every opt-in
    switch on — tcgen05 stem / 3x3 / strided convolutions, conv dgrad and wgrad — every kernel is
    reached, nothing stays pending, and the loss tracks the plain CPU engine.  (Gradients of this
    tiny-batch random network are chaotic — a 2e-3 input perturbation decorrelates them — so the
    per-op tests above carry the numerics; this one carries the plumbing.)"""
    from fake_ext import FullFakeExt
    from ddlbench_b200.engine import StageExecutor
    from ddlbench_b200.engine.plan import compile_plan
    from ddlbench_b200.models import build_model

    def run(fake_device: bool, steps: int = 2):
        torch.manual_seed(0)
        model = build_model("resnet50", "imagenet")
        plan = compile_plan(model)
        ex = StageExecutor(plan, device="cpu", dtype=BF16, channels_last=True, lr=0.002, momentum=0.9,
                           weight_decay=1e-4, tcgen05=fake_device)
        if fake_device:
            for op in plan.ops:                     # what to_device() does on a CUDA device
                if getattr(op, "ws", "absent") is None and hasattr(op, "channels"):
                    op.ws = ops.BnWorkspace(op.channels, "cpu")
        g = torch.Generator().manual_seed(5)
        losses = []
        for _ in range(steps):
            x = torch.randn(4, 3, 64, 64, generator=g)
            t = torch.randint(0, 1000, (4,), generator=g)
            if fake_device:
                x, t = fake_cuda(x), fake_cuda(t)
            outs, ctx = ex.forward([ex.prepare_input(x)])
            assert not E.PENDING_FOLD
            stats, d = ex.loss(outs[0], t)
            ex.backward(ctx, [d])
            ex.store.sgd_step()
            losses.append(float(stats[0]))
        return losses, ex.store.master.as_subclass(torch.Tensor).clone()

    ref_losses, _ = run(False)
    results = {}
    lib_convs = []

    class AtenSpy:
        """torch.ops.aten with the library convolutions recorded (the default path must not use them)."""
        def __getattr__(self, name):
            if name in ("convolution", "convolution_backward"):
                lib_convs.append(name)
            return getattr(torch.ops.aten, name)

    for label, switches in (("igemm", ()), ("default", ()), ("fold+wgrad1x1", ("BN_FOLD", "WGRAD_TCGEN05", "BN_BWD_FOLD")),
                            ("all", ("CONV3X3_TCGEN05", "WGRAD_TCGEN05", "STEM_IM2COL", "CONV3X3_DGRAD", "BN_FOLD",
                                     "DGRAD1X1_TCGEN05", "BN_BWD_FOLD"))):
        with monkeypatch.context() as mp:
            fe = FullFakeExt()
            mp.setattr(ops, "ext", lambda: fe)
            mp.setattr(ops, "IGEMM", label == "igemm")
            mp.setattr(ops, "BN_CLUSTER_FUSED", label == "igemm")     # the legacy folds exclude the cluster kernels
            mp.setattr(ops, "_WG_WS", {})
            if label == "igemm":
                mp.setattr(E, "aten", AtenSpy())
            mp.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
            for name in ("empty", "zeros", "empty_like", "zeros_like"):
                real = getattr(torch, name)
                mp.setattr(torch, name, (lambda real: lambda *a, **k: fake_cuda(real(*a, **k)))(real))
            for name in switches:
                mp.setattr(E, name, True)
            results[label] = run(True, 2 if label in ("all", "igemm") else 1) + (set(fe.calls),)
    (l_d, w_d, used_d), (l_f, w_f, used_f), (l_a, w_a, used_a) = (results[k] for k in ("default", "fold+wgrad1x1", "all"))
    # the shipped configuration: every convolution / FC layer on the persistent implicit-GEMM family,
    # statistics in the convolution epilogue, finalized by the BatchNorm apply kernel itself, no
    # library convolution anywhere in the step
    l_i, w_i, used_i = results["igemm"]
    assert {"igemm_conv_fprop", "igemm_conv_dgrad", "wgrad_conv", "wgrad_gemm", "igemm_gemm", "im2col_rows",
            "bn_apply_sums", "bn_bwd_reduce", "bn_bwd_apply", "xent", "sgd_flat"} <= used_i, used_i
    assert not ({"bn_stats", "bn_finalize", "bn_apply", "gemm_bf16_tn", "conv_bf16_nhwc"} & used_i), used_i
    assert lib_convs == [], lib_convs
    assert l_i == pytest.approx(ref_losses, rel=3e-2) and torch.isfinite(w_i).all()
    assert "gemm_bf16_tn_bnfold" not in used_d and {"gemm_bf16_tn_bnfold", "gemm_bf16_nt_mn",
                                                    "gemm_bf16_tn_bnbwd"} <= used_f
    assert l_f == pytest.approx(l_d, rel=5e-3) and ((w_f - w_d).norm() / w_d.norm()).item() < 1e-3
    for k in ("im2col_rows", "conv_bf16_nhwc", "conv_wgrad_bf16_nhwc", "gemm_bf16_tn_bnfold", "gemm_bf16_nt_mn",
              "gemm_bf16_tn", "bn_finalize", "bn_apply_sums", "bn_bwd_reduce", "bn_bwd_apply",
              "maxpool_fwd", "maxpool_bwd", "gap_fwd", "gap_bwd", "xent", "sgd_flat"):
        assert k in used_a, k
    assert "bn_stats" not in used_a            # every convolution delivered its statistics fused
    assert l_a == pytest.approx(ref_losses, rel=3e-2) and l_d == pytest.approx(ref_losses[:1], rel=3e-2)
    assert torch.isfinite(w_a).all()


@pytest.mark.parametrize("arch,ds,shape,expect", [
    ("mobilenetv2", "cifar10", (4, 3, 32, 32), {"dwconv3x3:0", "dwconv3x3:1", "dwconv3x3:2", "gemm_bf16_tn_bnfold",
                                                "gemm_bf16_nt_mn", "im2col_rows"}),
    ("vgg16", "cifar10", (4, 3, 32, 32), {"conv_bf16_nhwc", "conv_wgrad_bf16_nhwc", "im2col_rows", "bias_act", "colsum"}),
    ("resnet18", "mnist", (4, 1, 28, 28), {"conv_bf16_nhwc", "conv_wgrad_bf16_nhwc", "im2col_rows", "gemm_bf16_nt_mn"}),
])
def test_other_families_device_branches_optin_equals_default(monkeypatch, arch, ds, shape, expect):
    """MobileNetV2 (depthwise kernels, ReLU6 folds), VGG (bias + ReLU fused into the convolution) and
    the MNIST ResNet (1-channel stem): one step through the device branches with every opt-in
    switch on must give the update of the default device path (well-conditioned small networks)."""
    from fake_ext import FullFakeExt
    from ddlbench_b200.engine import StageExecutor
    from ddlbench_b200.engine.plan import compile_plan
    from ddlbench_b200.models import build_model

    def run():
        torch.manual_seed(0)
        plan = compile_plan(build_model(arch, ds))
        ex = StageExecutor(plan, device="cpu", dtype=BF16, channels_last=True, lr=0.002, momentum=0.9,
                           weight_decay=1e-4, tcgen05=True)
        for op in plan.ops:
            if getattr(op, "ws", "absent") is None and hasattr(op, "channels"):
                op.ws = ops.BnWorkspace(op.channels, "cpu")
        g = torch.Generator().manual_seed(5)
        x, t = fake_cuda(torch.randn(*shape, generator=g)), fake_cuda(torch.randint(0, 10, (shape[0],), generator=g))
        outs, ctx = ex.forward([ex.prepare_input(x)])
        assert not E.PENDING_FOLD
        stats, d = ex.loss(outs[0], t)
        ex.backward(ctx, [d])
        ex.store.sgd_step()
        return float(stats[0]), ex.store.master.as_subclass(torch.Tensor).clone()

    out = []
    for switches in ((), ("CONV3X3_TCGEN05", "WGRAD_TCGEN05", "STEM_IM2COL", "CONV3X3_DGRAD", "BN_FOLD", "DWCONV",
                          "DGRAD1X1_TCGEN05", "BN_BWD_FOLD")):
        with monkeypatch.context() as mp:
            fe = FullFakeExt()
            mp.setattr(ops, "ext", lambda: fe)
            mp.setattr(ops, "IGEMM", False)
            mp.setattr(ops, "BN_CLUSTER_FUSED", False)
            mp.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
            for name in ("empty", "zeros", "empty_like", "zeros_like"):
                real = getattr(torch, name)
                mp.setattr(torch, name, (lambda real: lambda *a, **k: fake_cuda(real(*a, **k)))(real))
            for name in switches:
                mp.setattr(E, name, True)
            out.append(run() + (set(fe.calls),))
    (l0, w0, _), (l1, w1, used) = out
    assert expect <= used, expect - used
    assert l1 == pytest.approx(l0, rel=1e-2) and ((w1 - w0).norm() / w0.norm()).item() < 2e-3
