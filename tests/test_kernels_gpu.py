"""Numerics of every hand-written CUDA kernel against a plain PyTorch fp32 reference of the
same op (ops/reference.py).  All tests here need the B200."""
import os
import subprocess
import sys

import pytest
import torch

from ddlbench_b200 import ops
from ddlbench_b200.ops import reference as ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _nhwc(n, c, h, w, scale=1.0, shift=0.0):
    x = (torch.randn(n, c, h, w, device=DEV) * scale + shift).to(torch.bfloat16)
    return x.contiguous(memory_format=torch.channels_last)


def test_extension_loaded():
    e = ops.ext()
    assert "tcgen05" in e.gemm_info()


@pytest.mark.parametrize("n", [8, 1000, 4096 + 3, 1 << 20])
@pytest.mark.parametrize("gdt,wdt", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32),
                                     (torch.bfloat16, torch.float32)])
def test_sgd_flat(n, gdt, wdt):
    torch.manual_seed(0)
    master = torch.randn(n, device=DEV)
    mom = torch.randn(n, device=DEV) * 0.1
    grad = torch.randn(n, device=DEV).to(gdt)
    hyper = torch.tensor([0.1, 0.9, 1e-4, 0.5], device=DEV)
    w = torch.empty(n, device=DEV, dtype=wdt)
    w2 = torch.empty(n, device=DEV, dtype=wdt)
    m_ref, v_ref = master.clone(), mom.clone()
    w_ref = torch.empty(n, device=DEV, dtype=wdt)
    ref.sgd_flat(m_ref, v_ref, grad, hyper, [w_ref])
    ops.sgd_flat_(master, mom, grad, hyper, w, w2)
    torch.testing.assert_close(master, m_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mom, v_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(w.float(), w_ref.float(), rtol=1e-2, atol=1e-3)
    assert torch.equal(w, w2)


@pytest.mark.parametrize("rows,C", [(32, 1000), (7, 10), (128, 32320), (256, 1000)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_xent(rows, C, dtype, smoothing):
    torch.manual_seed(1)
    logits = (torch.randn(rows, C, device=DEV) * 3).to(dtype)
    target = torch.randint(0, C, (rows,), device=DEV)
    for _ in range(2):  # twice: the ticket counter must self-reset
        stats, d = ops.xent(logits, target, smoothing=smoothing, grad_scale=2.0)
    rs, rd = ref.xent(logits, target, smoothing, 2.0 / rows)
    torch.testing.assert_close(stats[0], rs[0], rtol=2e-3, atol=2e-3)
    assert stats[1].item() == rs[1].item() and stats[2].item() == rs[2].item()
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    torch.testing.assert_close(d.float(), rd.float(), rtol=tol, atol=tol / rows)


@pytest.mark.parametrize("n,c,h,w", [(32, 64, 56, 56), (4, 2048, 7, 7), (8, 24, 16, 16),
                                     (2, 960, 7, 7), (3, 1280, 5, 5), (16, 8, 9, 11)])
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("has_res", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_bn_forward_backward(n, c, h, w, act, has_res, fused, monkeypatch):
    # fused=True: single-pass cluster kernels (where a channel slab fits in a cluster's shared
    # memory; (32, 64, 56, 56) does not and exercises the fallback), False: two-kernel path
    monkeypatch.setattr(ops, "BN_CLUSTER_FUSED", fused)
    torch.manual_seed(2)
    x = _nhwc(n, c, h, w, scale=1.5, shift=0.3)
    res = _nhwc(n, c, h, w) if has_res else None
    gamma = (torch.rand(c, device=DEV) + 0.5).to(torch.bfloat16)
    beta = (torch.randn(c, device=DEV) * 0.2).to(torch.bfloat16)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ws = ops.BnWorkspace(c, DEV)
    for _ in range(2):  # workspace must self-zero
        y, stat = ops.bn_forward(x, res, gamma, beta, rm, rv, 1e-5, 0.1, act, ws)
    stat_ref = ref.bn_stat(x, 1e-5)
    ref.bn_update_running(stat_ref, x.numel() // c, 1e-5, 0.1, rm_ref, rv_ref)
    ref.bn_update_running(stat_ref, x.numel() // c, 1e-5, 0.1, rm_ref, rv_ref)
    y_ref = ref.bn_apply(x, res, stat_ref, gamma, beta, act)
    torch.testing.assert_close(stat.view(2, c)[0], stat_ref[0], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(stat.view(2, c)[1], stat_ref[1], rtol=2e-3, atol=1e-3)
    torch.testing.assert_close(rm, rm_ref, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(rv, rv_ref, rtol=2e-3, atol=1e-3)
    torch.testing.assert_close(y.float(), y_ref.float(), rtol=2e-2, atol=2e-2)
    assert y.is_contiguous(memory_format=torch.channels_last)
    # backward
    dy = _nhwc(n, c, h, w)
    dy2 = _nhwc(n, c, h, w) if has_res else None
    dg = torch.empty(c, device=DEV, dtype=torch.bfloat16)
    db = torch.empty(c, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        dx, dz = ops.bn_backward(dy, dy2, y, x, stat, gamma, act, has_res, dg, db, False, ws)
    dx_r, dz_r, dg_r, db_r = ref.bn_backward(dy, dy2, y_ref, x, stat_ref, gamma, act, has_res)
    scale = max(dg_r.abs().max().item(), 1.0)
    torch.testing.assert_close(dg.float(), dg_r, rtol=3e-2, atol=3e-2 * scale)
    torch.testing.assert_close(db.float(), db_r, rtol=3e-2, atol=3e-2 * max(db_r.abs().max().item(), 1.0))
    # elements whose activation mask flips between bf16 y and fp32 y_ref are excluded
    err = (dx.float() - dx_r.float()).abs()
    assert (err > 0.05 * max(dx_r.float().abs().max().item(), 1e-3)).float().mean().item() < 2e-3
    if has_res:
        errz = (dz.float() - dz_r.float()).abs()
        assert (errz > 0.05).float().mean().item() < 2e-3
    if act:
        # 1-bit activation mask written by the forward: same backward as with y, bit for bit
        mask = torch.zeros(x.numel() // 8, dtype=torch.uint8, device=DEV)
        y2, stat2 = ops.bn_forward(x, res, gamma, beta, rm, rv, 1e-5, 0.1, act, ws, mask_out=mask)
        bits = ((mask.view(-1, 1) >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1).bool()
        yf = y2.permute(0, 2, 3, 1).reshape(-1).float()
        expect = (yf > 0) & (yf < 6) if act == 2 else yf > 0
        assert torch.equal(bits.view(-1), expect)
        dg1, db1, dg2, db2 = (torch.empty_like(dg) for _ in range(4))
        dx1, dz1 = ops.bn_backward(dy, dy2, y2, x, stat2, gamma, act, has_res, dg1, db1, False, ws)
        dx2, dz2 = ops.bn_backward(dy, dy2, y2, x, stat2, gamma, act, has_res, dg2, db2, False, ws,
                                   mask=mask)
        # identical up to the (non-deterministic) order of the partial-sum atomics
        assert (dx2.float() - dx1.float()).abs().max().item() <= \
            2e-2 * max(dx1.float().abs().max().item(), 1e-3)
        if has_res:
            assert torch.equal(dz2, dz1)


@pytest.mark.parametrize("m,c,tiles,ok", [(32 * 14 * 14, 1024, 2, True), (32 * 7 * 7, 2048, 2, True),
                                          (32 * 28 * 28, 128, 2, True), (32 * 28 * 28, 512, 1, True),
                                          (32 * 56 * 56, 64, 2, False), (32 * 112 * 112, 64, 1, False)])
def test_bn_cluster_plan(m, c, tiles, ok):
    sv, cl, good, rows, smem = ops.ext().bn_cluster_plan(m, c, tiles)
    assert ok or not good            # 56x56 / 112x112 slabs can never fit a cluster
    if good:
        assert sv in (4, 8, 16) and cl in (1, 2, 4, 8, 16) and rows * cl >= m
        assert smem <= 227 * 1024 and rows * sv * 16 * tiles <= smem


@pytest.mark.parametrize("n,c,h,w", [(32, 1024, 14, 14), (32, 128, 28, 28), (32, 512, 28, 28)])
def test_bn_cluster_large_layers(n, c, h, w, monkeypatch):
    """ResNet-50 layer2/3 shapes at the benchmark batch: cluster path vs fp32 reference."""
    monkeypatch.setattr(ops, "BN_CLUSTER_FUSED", True)
    torch.manual_seed(4)
    x = _nhwc(n, c, h, w, scale=1.3, shift=-0.2)
    res = _nhwc(n, c, h, w)
    gamma = (torch.rand(c, device=DEV) + 0.5).to(torch.bfloat16)
    beta = (torch.randn(c, device=DEV) * 0.2).to(torch.bfloat16)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    mask = torch.zeros(x.numel() // 8, dtype=torch.uint8, device=DEV)
    before = ops.launches
    y, stat = ops.bn_forward(x, res, gamma, beta, rm, rv, 1e-5, 0.1, 1, ops.BnWorkspace(c, DEV),
                             mask_out=mask)
    fits = ops.ext().bn_cluster_plan(n * h * w, c, 2)[2]       # x and res tiles
    assert ops.launches - before == (1 if fits else 2)          # one kernel, not stats + apply
    stat_ref = ref.bn_stat(x, 1e-5)
    y_ref = ref.bn_apply(x, res, stat_ref, gamma, beta, 1)
    torch.testing.assert_close(stat.view(2, c), stat_ref, rtol=2e-3, atol=1e-3)
    torch.testing.assert_close(y.float(), y_ref.float(), rtol=2e-2, atol=2e-2)
    dy, dy2 = _nhwc(n, c, h, w), _nhwc(n, c, h, w)
    dg = torch.empty(c, device=DEV, dtype=torch.bfloat16)
    db = torch.empty(c, device=DEV, dtype=torch.bfloat16)
    before = ops.launches
    dx, dz = ops.bn_backward(dy, dy2, y, x, stat, gamma, 1, True, dg, db, False, None, mask=mask)
    fits = ops.ext().bn_cluster_plan(n * h * w, c, 3)[2]       # dy, dy2 and x tiles
    assert ops.launches - before == (1 if fits else 2)
    dx_r, dz_r, dg_r, db_r = ref.bn_backward(dy, dy2, y, x, stat_ref, gamma, 1, True)
    torch.testing.assert_close(dg.float(), dg_r, rtol=3e-2, atol=3e-2 * max(dg_r.abs().max().item(), 1))
    torch.testing.assert_close(db.float(), db_r, rtol=3e-2, atol=3e-2 * max(db_r.abs().max().item(), 1))
    assert torch.equal(dz.float(), dz_r.to(torch.bfloat16).float())
    err = (dx.float() - dx_r.float()).abs()
    assert (err > 0.05 * max(dx_r.float().abs().max().item(), 1e-3)).float().mean().item() < 2e-3


def test_bn_eval_mode():
    x = _nhwc(4, 32, 8, 8)
    g = torch.ones(32, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(32, device=DEV, dtype=torch.bfloat16)
    rm, rv = torch.randn(32, device=DEV) * 0.1, torch.rand(32, device=DEV) + 0.5
    y, _ = ops.bn_forward(x, None, g, b, rm, rv, 1e-5, 0.1, 1, training=False)
    y_ref = torch.relu((x.float() - rm.view(1, -1, 1, 1)) * torch.rsqrt(rv + 1e-5).view(1, -1, 1, 1))
    torch.testing.assert_close(y.float(), y_ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("act", [0, 1])
def test_bias_act_and_grads(act):
    torch.manual_seed(3)
    x = _nhwc(8, 64, 14, 14)
    bias = torch.randn(64, device=DEV).to(torch.bfloat16)
    y = ops.bias_act(x, bias, act)
    torch.testing.assert_close(y.float(), ref.bias_act(x, bias, act).float(), rtol=2e-2, atol=2e-2)
    dy, dy2 = _nhwc(8, 64, 14, 14), _nhwc(8, 64, 14, 14)
    dz = ops.act_bwd(dy, dy2, y, act)
    torch.testing.assert_close(dz.float(), ref.act_bwd(dy, dy2, y, act).float(), rtol=2e-2, atol=2e-2)
    out = torch.empty(64, device=DEV, dtype=torch.bfloat16)
    ws = ops.BnWorkspace(64, DEV)
    ops.bias_grad_(dz, out, False, ws)
    ops.bias_grad_(dz, out, False, ws)
    torch.testing.assert_close(out.float(), dz.float().sum(dim=(0, 2, 3)), rtol=2e-2, atol=0.5)


def test_normalize_u8():
    x = torch.randint(0, 256, (4, 32, 32, 3), dtype=torch.uint8, device=DEV)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    y = ops.normalize_u8(x, mean, std)
    y_ref = ref.u8_to_norm(x, list(mean), list(std), torch.float32)
    assert y.shape == (4, 3, 32, 32) and y.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y.float(), y_ref, rtol=1e-2, atol=2e-2)
    x1 = torch.randint(0, 256, (5, 28, 28, 1), dtype=torch.uint8, device=DEV)
    y1 = ops.normalize_u8(x1, (0.1307,), (0.3081,))
    torch.testing.assert_close(y1.float(), ref.u8_to_norm(x1, [0.1307], [0.3081], torch.float32),
                               rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("F", [1024, 7, 2])
def test_revert_varlen_and_mask(dtype, F):
    T, B = 13, 5
    x = torch.randn(T, B, F, device=DEV).to(dtype)
    lengths = torch.tensor([13, 1, 7, 4, 10], device=DEV)
    y = ops.revert_varlen(x, lengths)
    assert torch.equal(y, ref.revert_varlen(x, lengths))
    assert torch.equal(ops.set_mask(lengths, T), ref.set_mask(lengths, T))


def test_global_avg_pool():
    x = _nhwc(32, 2048, 7, 7)
    y = ops.global_avg_pool(x)
    torch.testing.assert_close(y.float(), x.float().mean(dim=(2, 3)), rtol=2e-2, atol=1e-2)
    dy = torch.randn(32, 2048, device=DEV).to(torch.bfloat16)
    dx = ops.global_avg_pool_bwd(dy, x.shape)
    torch.testing.assert_close(dx.float(), (dy.float() / 49).view(32, 2048, 1, 1).expand(32, 2048, 7, 7),
                               rtol=2e-2, atol=1e-3)


@pytest.mark.parametrize("shape,k,s,p", [((8, 64, 112, 112), 3, 2, 1), ((4, 128, 32, 32), 2, 2, 0),
                                          ((2, 16, 9, 13), 3, 2, 1), ((2, 8, 7, 7), 3, 1, 1)])
def test_maxpool(shape, k, s, p):
    torch.manual_seed(4)
    x = _nhwc(*shape)
    assert ops.maxpool_supported(x, [k, k], [s, s], [p, p], [1, 1], False)
    y, idx = ops.maxpool_fwd(x, k, s, p)
    y_ref, i_ref = torch.ops.aten.max_pool2d_with_indices(x.float(), [k, k], [s, s], [p, p], [1, 1], False)
    assert torch.equal(y.float(), y_ref)
    dy = torch.randn_like(y)
    dx = ops.maxpool_bwd(dy, idx, x.shape, k, s, p)
    dx_ref = torch.ops.aten.max_pool2d_with_indices_backward(dy.float(), x.float(), [k, k], [s, s],
                                                             [p, p], [1, 1], False, i_ref)
    torch.testing.assert_close(dx.float(), dx_ref, rtol=2e-2, atol=2e-2)


def test_l2norm():
    x = torch.randn(100003, device=DEV).to(torch.bfloat16)
    torch.testing.assert_close(ops.l2norm(x)[0], x.float().norm(), rtol=1e-3, atol=1e-3)


# The tcgen05 GEMM runs in a child process with a timeout: a descriptor / barrier bug must
# fail this test, not take the whole session (or the GPU) down.
def test_tcgen05_gemm_subprocess():
    script = os.path.join(os.path.dirname(__file__), "gemm_check.py")
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, "tcgen05 GEMM check failed"


# Persistent implicit-GEMM family (the default path of every convolution): all 22 ResNet-50 conv
# shapes at batch 32, VGG at 224x224, odd channel counts, FC layers — forward (+fused BatchNorm
# statistics / finalize), data gradient and one-launch weight gradient against fp32 references.
# Child process + timeout: a descriptor / barrier bug fails the test, not the session.
def test_igemm_family_subprocess():
    script = os.path.join(os.path.dirname(__file__), "igemm_check.py")
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-6000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, "igemm family check failed"


# The small-M / long-K layers split K across a thread-block cluster; the default exchange of the
# partial sums goes through L2, this runs the distributed-shared-memory flavour and the round-2
# "slots" scheme on the layers that split (and the one-launch-per-class stride-2 data gradient).
@pytest.mark.parametrize("env", [{"DDLB_IG_SPLIT": "dsmem"}, {"DDLB_IG_SPLIT": "slots"},
                                 {"DDLB_IG_SPLIT": "off", "DDLB_IG_S2_MULTI": "0"}])
def test_igemm_split_variants_subprocess(env):
    script = os.path.join(os.path.dirname(__file__), "igemm_check.py")
    e = dict(os.environ, ONLY="l3.b1,l4.b1,l4.conv2,l3.conv2,odd.3x3.s2", **env)
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600, env=e)
    sys.stdout.write(p.stdout[-3000:])
    sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0, "igemm split-K variant check failed"


# CTA-pair (cta_group::2, 256-row tiles, multicast commit) GEMM against fp32, ragged shapes included.
def test_gemm_pair_subprocess():
    script = os.path.join(os.path.dirname(__file__), "gemm_pair_check.py")
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-3000:])
    sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0, "CTA-pair GEMM check failed"


# Round-1 one-tile-per-CTA implicit-GEMM convolution (kept as the A/B baseline of the persistent
# kernel; validated on hardware at the start of round 2), isolated in a child process.
def test_tcgen05_conv3x3():
    script = os.path.join(os.path.dirname(__file__), "conv_check.py")
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, "tcgen05 convolution check failed"


@pytest.mark.parametrize("n,c,hw,k,s,p", [(4, 3, 224, 7, 2, 3), (2, 3, 32, 3, 1, 1), (3, 1, 28, 3, 1, 1),
                                          (2, 3, 37, 7, 2, 3)])
def test_im2col_rows_and_stem_gemm(n, c, hw, k, s, p):
    torch.manual_seed(0)
    x = torch.randn(n, c, hw, hw, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    a = ops.im2col_rows(x, k, k, s, p)
    a_ref = ref.im2col_rows(x, k, k, s, p)
    assert a.shape == a_ref.shape and torch.equal(a.float(), a_ref.float().to(a.device))
    w = (torch.randn(64, c, k, k, device=DEV) * 0.1).to(torch.bfloat16)
    wm = torch.zeros(64, a.shape[1], device=DEV, dtype=torch.bfloat16)
    wm[:, :k * k * c] = w.permute(0, 2, 3, 1).reshape(64, -1)
    y = ops.gemm_tn(a, wm)
    oh = (hw + 2 * p - k) // s + 1
    y_ref = torch.nn.functional.conv2d(x.float(), w.float(), stride=s, padding=p)
    err = (y.float().view(n, oh, oh, 64).permute(0, 3, 1, 2) - y_ref).abs().max().item()
    assert err < 0.03 * max(1.0, y_ref.abs().max().item())


@pytest.mark.parametrize("n,c,hw,stride", [(8, 96, 28, 1), (4, 144, 29, 2), (2, 32, 112, 2), (3, 960, 7, 1)])
def test_dwconv3x3(n, c, hw, stride):
    torch.manual_seed(0)
    x = torch.randn(n, c, hw, hw, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(c, 1, 3, 3, device=DEV) * 0.3).to(torch.bfloat16)
    assert ops.dwconv_supported(x, w.shape, [stride, stride], [1, 1], [1, 1], c)
    y = ops.dwconv_fwd(x, w, stride, 1)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    y_ref = torch.nn.functional.conv2d(xr, wr, stride=stride, padding=1, groups=c)
    torch.testing.assert_close(y.float(), y_ref.detach(), rtol=2e-2, atol=2e-2)
    dy = torch.randn_like(y_ref).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y_ref.backward(dy.float())
    dx, dw = ops.dwconv_bwd(dy, x, w, stride, 1)
    torch.testing.assert_close(dx.float(), xr.grad, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(dw, wr.grad, rtol=2e-2, atol=2e-2 * max(1.0, wr.grad.abs().max().item()))


@pytest.mark.parametrize("n,c,hw,k,s,p,cip", [(4, 64, 14, 3, 1, 1, True), (2, 192, 35, 3, 1, 1, False),
                                             (3, 32, 17, 2, 2, 0, True), (2, 64, 8, 5, 3, 2, True)])
def test_avgpool_fixed_window(n, c, hw, k, s, p, cip):
    torch.manual_seed(0)
    x = _nhwc(n, c, hw, hw)
    oh = (hw + 2 * p - k) // s + 1
    y = ops.avgpool_fwd(x, (oh, oh), k, s, p, False, cip)
    # reference on the CPU: this build's CUDA avg_pool2d backward disagrees with its own CPU result
    # for channels_last fp32 inputs (tools/avgdbg.py: 1.47 max abs difference)
    xr = x.float().cpu().contiguous().requires_grad_(True)
    y_ref = torch.nn.functional.avg_pool2d(xr, k, s, p, False, cip)
    torch.testing.assert_close(y.float().cpu(), y_ref.detach(), rtol=1e-2, atol=1e-2)
    dy = torch.randn_like(y_ref).to(torch.bfloat16)
    y_ref.backward(dy.float())
    dx = ops.avgpool_bwd(dy.cuda().contiguous(memory_format=torch.channels_last), x.shape, k, s, p, False, cip)
    torch.testing.assert_close(dx.float().cpu(), xr.grad, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("n,c,hw,k,s,p", [(4, 3, 224, 7, 2, 3), (2, 3, 33, 7, 2, 3), (3, 3, 32, 3, 1, 1),
                                         (2, 1, 28, 5, 1, 2)])
def test_im2col_rows(n, c, hw, k, s, p):
    """Patch matrix of the first convolution (csrc/im2col.cu) against the host reference."""
    from ddlbench_b200.ops import reference as ref
    torch.manual_seed(0)
    x = torch.randn(n, c, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    before = ops.launches
    a = ops.im2col_rows(x, k, k, s, p)
    assert ops.launches == before + 1
    want = ref.im2col_rows(x.float().cpu(), k, k, s, p)
    assert a.shape == want.shape
    torch.testing.assert_close(a.float().cpu(), want, rtol=0, atol=0)


@pytest.mark.parametrize("n,c,hw,has_res", [(8, 64, 14, False), (4, 256, 7, True), (2, 2048, 7, True), (32, 128, 28, False)])
def test_bn_apply_finalizes_from_replicated_sums(n, c, hw, has_res):
    """Apply-side finalize (csrc/bn.cu, bn_apply_kernel with `sums`): statistics from the replicated
    sum / sum-of-squares buffers, y = relu(bn(x) + res), stat_out, running statistics, and the
    workspace / ticket left zeroed for the next forward."""
    torch.manual_seed(0)
    x = _nhwc(n, c, hw, hw)
    res = _nhwc(n, c, hw, hw) if has_res else None
    M = n * hw * hw
    xf = x.permute(0, 2, 3, 1).reshape(M, c).float()
    ws = ops.BnWorkspace(c, "cuda")
    sv = ws.sums.view(ops.BN_REPLICAS, 2, c)
    half = M // 3
    sv[0, 0], sv[0, 1] = xf[:half].sum(0), (xf[:half] ** 2).sum(0)          # spread over two replicas
    sv[5, 0], sv[5, 1] = xf[half:].sum(0), (xf[half:] ** 2).sum(0)
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    mask = torch.empty(x.numel() // 8, dtype=torch.uint8, device="cuda")
    y, stat = ops.bn_forward(x, res, gamma, beta, rm, rv, 1e-5, 0.1, ops.ACT_RELU, ws, stats_ready=True,
                             mask_out=mask)
    mean, var = xf.mean(0), xf.var(0, unbiased=False)
    want = (xf - mean) * torch.rsqrt(var + 1e-5) * gamma + beta
    if has_res:
        want = want + res.permute(0, 2, 3, 1).reshape(M, c).float()
    want = want.relu()
    got = y.permute(0, 2, 3, 1).reshape(M, c).float()
    torch.testing.assert_close(got, want, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(stat[0], mean, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(stat[1], torch.rsqrt(var + 1e-5), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(rm, 0.1 * mean, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rv, 0.9 + 0.1 * var * M / (M - 1), rtol=1e-3, atol=1e-4)
    assert float(ws.sums.abs().sum()) == 0.0 and int(ws.counter.abs().sum()) == 0
    bits = ((got > 0).view(M, c // 8, 8).to(torch.int32) << torch.arange(8, device="cuda", dtype=torch.int32)).sum(-1)
    assert torch.equal(mask.view(M, c // 8).to(torch.int32), bits)


@pytest.mark.parametrize("n,c,hw,out", [(4, 64, 13, 6), (2, 512, 14, 7), (3, 128, 9, 1), (2, 64, 10, 4)])
def test_avgpool_adaptive(n, c, hw, out):
    torch.manual_seed(0)
    x = _nhwc(n, c, hw, hw)
    y = ops.avgpool_fwd(x, (out, out), 0, 0, 0, True)
    xr = x.float().cpu().contiguous().requires_grad_(True)
    y_ref = torch.nn.functional.adaptive_avg_pool2d(xr, out)
    torch.testing.assert_close(y.float().cpu(), y_ref.detach(), rtol=1e-2, atol=1e-2)
    dy = torch.randn_like(y_ref).to(torch.bfloat16)
    y_ref.backward(dy.float())
    dx = ops.avgpool_bwd(dy.cuda().contiguous(memory_format=torch.channels_last), x.shape, 0, 0, 0, True)
    torch.testing.assert_close(dx.float().cpu(), xr.grad, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("p", [0.5, 0.1])
def test_dropout_philox(p):
    torch.manual_seed(0)
    x = torch.randn(64, 4096, device=DEV).to(torch.bfloat16)
    y, mask = ops.dropout_fwd(x, p)
    bits = ((mask.view(-1, 1).to(torch.int32) >> torch.arange(8, device=DEV, dtype=torch.int32)) & 1).view(-1).bool()
    keep = bits.float().mean().item()
    assert abs(keep - (1 - p)) < 0.01, keep                     # drop rate
    ref = torch.where(bits.view_as(x), x.float() / (1 - p), torch.zeros_like(x, dtype=torch.float32))
    torch.testing.assert_close(y.float(), ref, rtol=1e-2, atol=1e-2)
    dy = torch.randn_like(x)
    dx = ops.dropout_bwd(dy, mask, p)
    torch.testing.assert_close(dx.float(), torch.where(bits.view_as(x), dy.float() / (1 - p), torch.zeros_like(ref)),
                               rtol=1e-2, atol=1e-2)
    # a new step -> a different mask; same step + same host RNG state -> the same mask (recompute)
    st = torch.get_rng_state()
    _, m1 = ops.dropout_fwd(x, p)
    torch.set_rng_state(st)
    _, m2 = ops.dropout_fwd(x, p)
    assert torch.equal(m1, m2)
    ops.rng_tick(x.device)
    torch.set_rng_state(st)
    _, m3 = ops.dropout_fwd(x, p)
    assert not torch.equal(m1, m3)
    # neighbouring bytes are not correlated (each thread draws two Philox blocks)
    b = mask.view(-1)[:4096].to(torch.float32)
    assert abs(torch.corrcoef(torch.stack([b[:-1], b[1:]]))[0, 1].item()) < 0.1
