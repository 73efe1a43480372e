"""Static engine (explicit forward/backward, flat buffers, fused SGD) against torch autograd."""
import pytest
import torch
import torch.nn.functional as F

from ddlbench_b200.engine import StageExecutor, compile_plan
from ddlbench_b200.models import build_model


def _global_rel_err(ex, model):
    num = den = 0.0
    for name, p in model.named_parameters():
        g = ex.store.g_views[name]
        num += (g - p.grad).pow(2).sum().item()
        den += p.grad.pow(2).sum().item()
    return (num / max(den, 1e-30)) ** 0.5


@pytest.mark.parametrize("arch,ds", [("resnet18", "cifar10"), ("resnet50", "mnist"),
                                     ("vgg11", "cifar10"), ("mobilenetv2", "cifar10")])
def test_engine_matches_autograd(arch, ds):
    # fp64 so that a wiring bug cannot hide behind rounding noise of a deep BatchNorm net
    torch.manual_seed(0)
    m = build_model(arch, ds).double().train()
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float64, lr=0.1, momentum=0.9,
                                  weight_decay=1e-4)
    for name, p in m.named_parameters():
        ex.store.weights()[name].copy_(p.detach())
    x = torch.randn(8, *m.input_shape, dtype=torch.float64)
    t = torch.randint(0, 10, (8,))
    y = m(x)
    loss = F.cross_entropy(y, t)
    loss.backward()
    outs, ctx = ex.forward([ex.prepare_input(x)])
    stats, d = ex.loss(outs[0], t)
    ex.backward(ctx, [d])
    assert torch.allclose(outs[0], y, atol=1e-9)
    assert abs(stats[0].item() - loss.item()) < 1e-5
    assert _global_rel_err(ex, m) < 1e-6
    # BatchNorm running statistics were updated like nn.BatchNorm2d does
    for k, v in ex.buffers.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            ref = dict(m.named_buffers())[k]
            assert torch.allclose(v.double(), ref, atol=1e-5), k


def test_fused_sgd_matches_torch_optim():
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist").train()
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float32, lr=0.1, momentum=0.9,
                                  weight_decay=1e-4)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(4, 1, 28, 28)
    t = torch.randint(0, 10, (4,))
    for _ in range(3):
        opt.zero_grad()
        F.cross_entropy(m(x), t).backward()
        opt.step()
        o, c = ex.forward([x])
        s, d = ex.loss(o[0], t)
        ex.backward(c, [d])
        ex.store.sgd_step()
    for name, p in m.named_parameters():
        assert torch.allclose(ex.store.weights()[name], p, atol=1e-4), name


def test_fusion_patterns_resnet50():
    plan = compile_plan(build_model("resnet50", "imagenet"))
    kinds = [op.kind for op in plan.ops]
    assert kinds.count("conv_bn_act") == 53
    assert "gap_flatten" in kinds and "maxpool" in kinds and kinds[-1] == "linear_act"
    # every residual join is fused into a conv+bn op
    assert "add_act" not in kinds


def test_gradient_accumulation_over_microbatches():
    torch.manual_seed(1)
    m = build_model("vgg11", "mnist").train()
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
    x = torch.randn(8, 1, 28, 28)
    t = torch.randint(0, 10, (8,))
    o, c = ex.forward([x])
    s, d = ex.loss(o[0], t)
    ex.backward(c, [d])
    full = ex.store.g_flat.clone()
    for k, (a, b) in enumerate([(0, 4), (4, 8)]):
        o, c = ex.forward([x[a:b]])
        s, d = ex.loss(o[0], t[a:b], grad_scale=0.5)
        ex.backward(c, [d], accumulate=k > 0, final=k == 1)
    assert torch.allclose(ex.store.g_flat, full, atol=1e-5, rtol=1e-4)


def test_bucket_ready_hook_fires_in_order():
    m = build_model("resnet18", "cifar10").train()
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float32, bucket_elems=1 << 20)
    fired = []
    ex.on_bucket_ready = fired.append
    x = torch.randn(2, 3, 32, 32)
    o, c = ex.forward([x])
    s, d = ex.loss(o[0], torch.tensor([1, 2]))
    ex.backward(c, [d])
    assert fired == list(range(len(ex.store.layout.buckets))) and len(fired) > 3


def test_stage_split_equals_whole():
    torch.manual_seed(2)
    m = build_model("resnet18", "cifar10").train()
    whole = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
    cut = next(i for i, n in enumerate(m.nodes) if n.name == "layer2.0.bn1") + 1  # inside a block
    s0 = StageExecutor.from_model(m, 0, cut, device="cpu", dtype=torch.float32)
    s1 = StageExecutor.from_model(m, cut, None, device="cpu", dtype=torch.float32)
    assert len(s0.plan.out_names) == 2 and s1.plan.in_names == s0.plan.out_names
    x = torch.randn(4, 3, 32, 32)
    t = torch.randint(0, 10, (4,))
    o, c = whole.forward([x]); s, d = whole.loss(o[0], t); whole.backward(c, [d])
    a, ca = s0.forward([x])
    b, cb = s1.forward(a)
    s2, d2 = s1.loss(b[0], t)
    gin = s1.backward(cb, [d2])
    s0.backward(ca, gin)
    assert abs(s[0].item() - s2[0].item()) < 1e-5
    for name in s0.store.layout.entries:
        assert torch.allclose(s0.store.g_views[name], whole.store.g_views[name], atol=1e-5), name
    for name in s1.store.layout.entries:
        assert torch.allclose(s1.store.g_views[name], whole.store.g_views[name], atol=1e-5), name


@pytest.mark.parametrize("arch,hw", [("squeezenet1_1", 48), ("densenet121", 32), ("mobilenet", 32),
                                     ("inception_v3", 75), ("nasnetamobile", 64)])
def test_engine_runs_extra_zoo(arch, hw):
    """Concat joins and pre-activation BatchNorm (``ConcatOp`` / ``BnAct``) against autograd."""
    from ddlbench_b200.models.extra import EXTRA_BUILDERS
    torch.manual_seed(0)
    m = EXTRA_BUILDERS[arch](arch, num_classes=10, input_hw=hw).double().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float64, lr=0.1, momentum=0.9,
                                  weight_decay=1e-4)
    kinds = {op.kind for op in ex.plan.ops}
    assert "module" not in kinds                       # nothing falls back to the autograd op
    x_batch = 2 if arch in ("inception_v3", "nasnetamobile") else 4
    for name, p in m.named_parameters():
        ex.store.weights()[name].copy_(p.detach())
    x = torch.randn(x_batch, 3, hw, hw, dtype=torch.float64)
    t = torch.randint(0, 10, (x_batch,))
    y = m(x)
    F.cross_entropy(y, t).backward()
    outs, ctx = ex.forward([ex.prepare_input(x)])
    _, d = ex.loss(outs[0], t)
    ex.backward(ctx, [d])
    assert torch.allclose(outs[0], y, atol=1e-8)
    assert _global_rel_err(ex, m) < 1e-6


def test_extra_zoo_parameter_counts():
    from ddlbench_b200.models import build_model
    want = {"densenet121": 7978856, "densenet161": 28681000, "squeezenet1_0": 1248424,
            "squeezenet1_1": 1235496, "resnext50": 25028904, "inception_v3": 23834568,
            "mobilenet": 4231976}          # torchvision's counts (mobilenet: Howard et al. 4.2 M)
    for arch, n in want.items():
        assert sum(p.numel() for p in build_model(arch, "imagenet").parameters()) == n, arch
    nas = build_model("nasnetamobile", "imagenet")
    assert 5.0e6 < sum(p.numel() for p in nas.parameters()) < 5.6e6      # 5.3 M in the paper


def test_per_bucket_sgd_overlap_equals_single_update(monkeypatch):
    """DDLB_SGD_OVERLAP: updating bucket by bucket from the bucket-ready hook == one flat update."""
    from ddlbench_b200.harness import Trainer
    x = torch.randn(4, 1, 28, 28)
    t = torch.randint(0, 10, (4,))
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("DDLB_SGD_OVERLAP", flag)
        tr = Trainer("resnet18", "mnist", batch_size=4, device="cpu", dtype=torch.float32, lr=0.05,
                     momentum=0.9, weight_decay=1e-4, cuda_graph=False, seed=7, bucket_elems=1 << 18)
        assert len(tr.ex.store.layout.buckets) > 3
        for _ in range(3):
            tr.step(x, t)
        res[flag] = tr.ex.store.master.clone()
    assert torch.equal(res["0"], res["1"])


@pytest.mark.parametrize("nb,ci,co,hw,k,stride", [(3, 64, 64, 9, 3, 1), (2, 64, 128, 7, 3, 1),
                                                  (5, 64, 64, 14, 3, 1), (2, 128, 64, 10, 5, 1),
                                                  (3, 64, 64, 14, 3, 2), (2, 64, 128, 9, 1, 2),
                                                  (2, 64, 64, 8, 3, 2)])
def test_conv_tcgen05_tiling_logic_emulated(nb, ci, co, hw, k, stride):
    """Host-side emulation of csrc/conv_tcgen05.cu: same tile boxes (from the C++ chooser), same
    (tap, channel-slice) K loop, same shifted / zero-filled input boxes, same row -> pixel
    mapping and clipped stores — must reproduce conv2d.  (Checks the algorithm the kernel
    implements; the tensor-core descriptors themselves need the GPU test.)"""
    import torch.nn.functional as Fn
    from ddlbench_b200 import ops
    ext = ops.load_ext(build_if_missing=False)
    pad = k // 2
    ohw = (hw + 2 * pad - k) // stride + 1
    TW, TH, TN = ext.conv_box(nb, ohw, ohw)
    assert TW * TH * TN == 128
    torch.manual_seed(0)
    x = torch.randn(nb, hw, hw, ci, dtype=torch.float64)                  # NHWC
    w = torch.randn(co, k, k, ci, dtype=torch.float64)                    # KRSC
    wmat = w.reshape(co, k * k * ci)
    BLOCK_N, BLOCK_K = 64, 64
    y = torch.zeros(nb, ohw, ohw, co, dtype=torch.float64)
    tiles_w, tiles_h, tiles_nb = -(-ohw // TW), -(-ohw // TH), -(-nb // TN)

    def box(t, n0, h0, w0, c0):           # 4-D TMA load with zero fill outside the tensor
        out = torch.zeros(TN, TH, TW, BLOCK_K, dtype=torch.float64)
        for a in range(TN):
            for b in range(TH):
                for c in range(TW):
                    # element strides of the tensor map: every `stride`-th pixel from the start
                    n, h, ww = n0 + a, h0 + b * stride, w0 + c * stride
                    if 0 <= n < nb and 0 <= h < hw and 0 <= ww < hw:
                        out[a, b, c] = t[n, h, ww, c0:c0 + BLOCK_K]
        return out.reshape(128, BLOCK_K)  # row = (ln * TH + lh) * TW + lw

    for tile_pix in range(tiles_w * tiles_h * tiles_nb):
        w0 = (tile_pix % tiles_w) * TW
        h0 = ((tile_pix // tiles_w) % tiles_h) * TH
        n0 = (tile_pix // (tiles_w * tiles_h)) * TN
        for tile_n in range(co // BLOCK_N):
            acc = torch.zeros(128, BLOCK_N, dtype=torch.float64)
            for kb in range(k * k * (ci // BLOCK_K)):
                tap, cb = divmod(kb, ci // BLOCK_K)
                kh, kw = divmod(tap, k)
                a = box(x, n0, h0 * stride + kh - pad, w0 * stride + kw - pad, cb * BLOCK_K)
                bmat = wmat[tile_n * BLOCK_N:(tile_n + 1) * BLOCK_N,
                            tap * ci + cb * BLOCK_K: tap * ci + (cb + 1) * BLOCK_K]
                acc += a @ bmat.t()
            for row in range(128):
                lw, lh, ln = row % TW, (row // TW) % TH, row // (TW * TH)
                if w0 + lw < ohw and h0 + lh < ohw and n0 + ln < nb:     # clipped store
                    y[n0 + ln, h0 + lh, w0 + lw, tile_n * BLOCK_N:(tile_n + 1) * BLOCK_N] = acc[row]
    ref = Fn.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), stride=stride,
                    padding=pad).permute(0, 2, 3, 1)
    torch.testing.assert_close(y, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("nb,ci,co,hw,k,stride,pad", [(2, 16, 24, 9, 3, 2, 1), (2, 8, 16, 8, 3, 2, 1), (3, 8, 8, 7, 1, 2, 0),
                                                      (2, 8, 16, 10, 5, 2, 2), (2, 16, 8, 6, 3, 1, 1)])
def test_igemm_dgrad_parity_classes_emulated(nb, ci, co, hw, k, stride, pad):
    """Host emulation of ``ddl_igemm_conv_dgrad`` (csrc/igemm_tcgen05.cu): a stride-2 data gradient is
    one stride-1 problem per output parity class (a, b) over the taps with (a + pad - kh) and
    (b + pad - kw) even, reading dY at the class pixel + tap offset (nh/2, nw/2) (zero outside:
    the TMA fill), multiplying by the KRSC filter block of the tap read MN-major (W[:, tap, :] as
    [Cout][Cin], no transpose), and storing through the strided view dX[:, a::2, b::2, :].  Classes
    without taps stay zero (the memset).  Must equal autograd's input gradient."""
    import torch.nn.functional as Fn
    torch.manual_seed(0)
    oh = (hw + 2 * pad - k) // stride + 1
    x = torch.randn(nb, ci, hw, hw, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, k, k, dtype=torch.float64)
    y = Fn.conv2d(x, w, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    dyh = dy.permute(0, 2, 3, 1)                                   # NHWC [nb, oh, oh, co]
    wk = w.permute(0, 2, 3, 1).reshape(co, k * k * ci)             # KRSC filter matrix
    dx = torch.zeros(nb, hw, hw, ci, dtype=torch.float64)
    classes = [(0, 0)] if stride == 1 else [(a, b) for a in range(2) for b in range(2)]
    launched = 0
    for a, b in classes:
        Ha, Wb = (hw - a + stride - 1) // stride, (hw - b + stride - 1) // stride
        taps = []
        for kh in range(k):
            for kw in range(k):
                nh, nw = a + pad - kh, b + pad - kw
                if stride == 1:
                    taps.append((pad - kh, pad - kw, (kh * k + kw) * ci))
                elif nh % 2 == 0 and nw % 2 == 0:
                    taps.append((nh // 2, nw // 2, (kh * k + kw) * ci))
        if not taps or Ha <= 0 or Wb <= 0:
            continue
        launched += 1
        out = torch.zeros(nb, Ha, Wb, ci, dtype=torch.float64)
        for dh, dw, bcol in taps:
            shifted = torch.zeros(nb, Ha, Wb, co, dtype=torch.float64)     # dY box at the tap offset
            for i in range(Ha):
                for j in range(Wb):
                    p_, q_ = i + dh, j + dw
                    if 0 <= p_ < oh and 0 <= q_ < oh:
                        shifted[:, i, j] = dyh[:, p_, q_]
            out += shifted @ wk[:, bcol:bcol + ci]                         # B MN-major: [K = Cout][N = Cin]
        dx[:, a::stride, b::stride] = out
    assert launched >= 1
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), x.grad, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("S,total_kb", [(2, 9), (4, 16), (4, 10), (2, 72)])
def test_igemm_cluster_splitk_exchange_emulated(S, total_kb):
    """Host emulation of the cluster split-K epilogue (csrc/igemm_tcgen05.cu, ``csplit``): slice z
    accumulates k-blocks [z*kbps, (z+1)*kbps), owns W = 128 / S output columns, parks the 32-column
    groups it does NOT own in its slot of the workspace laid out [column quad][row][4] (L2
    flavour) or in the owner's receive buffer [slot][column quad][row][4] with slot = source index
    among the other slices (DSMEM flavour); every owner then adds the S - 1 partials of its columns
    to its own.  Both flavours must reproduce the full reduction, and the host's slice count must
    leave no slice empty."""
    torch.manual_seed(0)
    BM, BN = 128, 128
    W = BN // S
    kbps = -(-total_kb // S)
    assert -(-total_kb // kbps) == S, "launch_ig would have fallen back to fewer slices"
    partial_kb = torch.randn(total_kb, BM, BN, dtype=torch.float64)
    acc = [partial_kb[z * kbps:min((z + 1) * kbps, total_kb)].sum(0) for z in range(S)]
    # L2 flavour: ws[src][quad][row][4]
    ws = torch.zeros(S, BN // 4, BM, 4, dtype=torch.float64)
    # DSMEM flavour: rbuf[owner][slot][quad in slice][row][4]
    rbuf = torch.zeros(S, S - 1, W // 4, BM, 4, dtype=torch.float64)
    for rank in range(S):
        for g in range(BN // 32):
            dest = (g * 32) // W
            if dest == rank:
                continue
            for j in range(8):
                quad = acc[rank][:, g * 32 + 4 * j: g * 32 + 4 * j + 4]          # [row][4]
                ws[rank, g * 8 + j] = quad
                slot = rank - (1 if rank > dest else 0)
                rbuf[dest, slot, ((g * 32) % W) // 4 + j] = quad
    for via_l2 in (True, False):
        out = torch.zeros(BM, BN, dtype=torch.float64)
        for rank in range(S):
            for g in range(W // 32):
                f = acc[rank][:, rank * W + g * 32: rank * W + g * 32 + 32].clone()
                for sl in range(S - 1):
                    srck = sl + (1 if sl >= rank else 0)
                    for j in range(8):
                        q = ws[srck, (rank * W + g * 32) // 4 + j] if via_l2 else rbuf[rank, sl, g * 8 + j]
                        f[:, 4 * j:4 * j + 4] += q
                out[:, rank * W + g * 32: rank * W + g * 32 + 32] = f
        torch.testing.assert_close(out, partial_kb.sum(0), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("tiles_m,tiles_n,group_m,clusters", [(32, 32, 8, 74), (7, 2, 8, 14), (25, 1, 8, 25), (3, 17, 8, 51),
                                                             (9, 5, 4, 74)])
def test_gemm_pair_tile_walk_and_operand_split_emulated(tiles_m, tiles_n, group_m, clusters):
    """Host emulation of csrc/gemm2cta_tcgen05.cu: (1) ``tile_coords`` — the persistent walk of cluster c
    over tiles c, c + clusters, ... with m-tiles grouped for L2 reuse visits every (m, n) tile exactly
    once; (2) the CTA-pair operand split — CTA r stages rows [256 tm + 128 r, +128) of A and rows
    [BN tn + BN/2 r, +BN/2) of B, the pair's MMA is A_pair (256 x K) times B_pair^T (K x BN), CTA r's
    TMEM holds output rows 128 r .. 128 r + 127: reassembling the tiles must give A B^T."""
    def tile_coords(t):
        per_group = group_m * tiles_n
        g = t // per_group
        first = g * group_m
        gsz = min(group_m, tiles_m - first)
        r = t - g * per_group
        return first + r % gsz, r // gsz
    seen = set()
    for c in range(min(clusters, tiles_m * tiles_n)):
        for t in range(c, tiles_m * tiles_n, clusters):
            tm, tn = tile_coords(t)
            assert 0 <= tm < tiles_m and 0 <= tn < tiles_n and (tm, tn) not in seen
            seen.add((tm, tn))
    assert len(seen) == tiles_m * tiles_n
    if tiles_m * tiles_n > 40:
        return
    torch.manual_seed(0)
    BN, K = 64, 24
    M, N = tiles_m * 256 - 100, tiles_n * BN - 8                           # ragged edges: TMA zero fill / clipped store
    A = torch.randn(M, K, dtype=torch.float64)
    B = torch.randn(N, K, dtype=torch.float64)
    Ap = torch.zeros(tiles_m * 256, K, dtype=torch.float64); Ap[:M] = A
    Bp = torch.zeros(tiles_n * BN, K, dtype=torch.float64); Bp[:N] = B
    D = torch.zeros(tiles_m * 256, tiles_n * BN, dtype=torch.float64)
    for tm, tn in seen:
        a_cta = [Ap[tm * 256 + 128 * r: tm * 256 + 128 * (r + 1)] for r in range(2)]
        b_cta = [Bp[tn * BN + (BN // 2) * r: tn * BN + (BN // 2) * (r + 1)] for r in range(2)]
        b_pair = torch.cat(b_cta)                                            # both halves feed both tensor cores
        for r in range(2):                                                   # CTA r's accumulator = its 128 rows
            D[tm * 256 + 128 * r: tm * 256 + 128 * (r + 1), tn * BN:(tn + 1) * BN] = a_cta[r] @ b_pair.t()
    torch.testing.assert_close(D[:M, :N], A @ B.t(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n,c,hw,k,s,p", [(2, 3, 20, 7, 2, 3), (1, 3, 9, 3, 1, 1), (2, 1, 12, 3, 1, 1)])
def test_im2col_kernel_index_math_emulated(n, c, hw, k, s, p):
    """The staging / gather arithmetic of csrc/im2col.cu (rows[] fill, k -> (kh, rem), base offset)
    replayed on the host against the reference patch matrix."""
    from ddlbench_b200.ops import reference as ref
    torch.manual_seed(0)
    x = torch.randn(n, c, hw, hw)
    xh = x.permute(0, 2, 3, 1).contiguous().reshape(-1)                 # NHWC flat
    H = W = hw
    OH = (H + 2 * p - k) // s + 1
    OW = (W + 2 * p - k) // s + 1
    K, R = k * k * c, k * c
    kpad = (K + 7) // 8 * 8
    rowlen = (W + 2 * p) * c
    out = torch.zeros(n * OH * OW, kpad)
    G = 4                                                   # kIm2colRows: output rows per CTA
    groups = (OH + G - 1) // G
    koff = [(kk // R) * rowlen + (kk % R) if kk < K else -1 for kk in range(kpad)]
    for blk in range(n * groups):
        nn_, g = divmod(blk, groups)
        oh0 = g * G
        nrows = min(G, OH - oh0)
        in_rows = (nrows - 1) * s + k
        ih0 = oh0 * s - p
        rows = torch.zeros(in_rows * rowlen + 1)            # + the zero slot
        for r in range(in_rows):
            ih = ih0 + r
            for j in range(rowlen):
                iwc = j - p * c
                if 0 <= ih < H and 0 <= iwc < W * c:
                    rows[r * rowlen + j] = xh[(nn_ * H + ih) * W * c + iwc]
        for px in range(nrows * OW):
            r, ow = divmod(px, OW)
            base = r * s * rowlen + ow * s * c
            for kk in range(kpad):
                src = base + koff[kk] if koff[kk] >= 0 else in_rows * rowlen
                out[(nn_ * OH + oh0) * OW + px, kk] = rows[src]
    torch.testing.assert_close(out, ref.im2col_rows(x, k, k, s, p))


@pytest.mark.parametrize("nb,ci,co,hw,k,stride", [(3, 64, 64, 9, 3, 1), (2, 64, 128, 8, 3, 2),
                                                  (2, 128, 64, 7, 1, 2)])
def test_conv_wgrad_tcgen05_logic_emulated(nb, ci, co, hw, k, stride):
    """Host emulation of conv_wgrad_bf16_kernel (csrc/gemm_wgrad_tcgen05.cu): per filter tap, the
    reduction over 64-pixel boxes of dY and the tap-shifted (strided, zero-filled) X boxes must
    give the weight gradient of conv2d."""
    import torch.nn.functional as Fn
    torch.manual_seed(0)
    pad = k // 2
    ohw = (hw + 2 * pad - k) // stride + 1
    x = torch.randn(nb, hw, hw, ci, dtype=torch.float64)
    dy = torch.randn(nb, ohw, ohw, co, dtype=torch.float64)
    # same chooser as the C++ helper, for 64-pixel boxes (mirrors choose_pixel_box)
    best, box = -1.0, None
    tw = 1
    while tw <= 64:
        th = 1
        while tw * th <= 64:
            tn = 64 // (tw * th)
            eff = (ohw / (-(-ohw // tw) * tw)) * (ohw / (-(-ohw // th) * th)) * (nb / (-(-nb // tn) * tn))
            score = eff + 1e-3 * tw / 64.0 + 1e-4 * th / 64.0
            if score > best:
                best, box = score, (tw, th, tn)
            th *= 2
        tw *= 2
    TW, TH, TN = box
    tiles_w, tiles_h, tiles_nb = -(-ohw // TW), -(-ohw // TH), -(-nb // TN)

    def load(t, size, n0, h0, w0, step):       # 4-D TMA box -> [64 pixels, C]
        out = torch.zeros(TN, TH, TW, t.shape[3], dtype=torch.float64)
        for a in range(TN):
            for b in range(TH):
                for c in range(TW):
                    n, h, ww = n0 + a, h0 + b * step, w0 + c * step
                    if 0 <= n < nb and 0 <= h < size and 0 <= ww < size:
                        out[a, b, c] = t[n, h, ww]
        return out.reshape(64, -1)

    dw = torch.zeros(co, k, k, ci, dtype=torch.float64)
    for tap in range(k * k):
        kh, kw = divmod(tap, k)
        for pt in range(tiles_w * tiles_h * tiles_nb):
            w0 = (pt % tiles_w) * TW
            h0 = ((pt // tiles_w) % tiles_h) * TH
            n0 = (pt // (tiles_w * tiles_h)) * TN
            a = load(dy, ohw, n0, h0, w0, 1)                                        # [64, Cout]
            b = load(x, hw, n0, h0 * stride + kh - pad, w0 * stride + kw - pad, stride)   # [64, Cin]
            dw[:, kh, kw, :] += a.t() @ b
    xr = x.permute(0, 3, 1, 2)
    wr = torch.zeros(co, ci, k, k, dtype=torch.float64, requires_grad=True)
    Fn.conv2d(xr, wr, stride=stride, padding=pad).backward(dy.permute(0, 3, 1, 2))
    torch.testing.assert_close(dw.permute(0, 3, 1, 2), wr.grad, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("hw,stride,pad", [(7, 1, 1), (8, 2, 1), (9, 2, 1)])
def test_dwconv_kernel_index_math_emulated(hw, stride, pad):
    """The tap / bounds / divisibility arithmetic of csrc/dwconv.cu (forward, dgrad, wgrad)
    replayed on the host against autograd."""
    import torch.nn.functional as Fn
    torch.manual_seed(0)
    n, c = 2, 8
    x = torch.randn(n, c, hw, hw, dtype=torch.float64, requires_grad=True)
    w = torch.randn(c, 1, 3, 3, dtype=torch.float64, requires_grad=True)
    y_ref = Fn.conv2d(x, w, stride=stride, padding=pad, groups=c)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    oh_n = ow_n = (hw + 2 * pad - 3) // stride + 1
    wt = w.detach().reshape(c, 9).t()                       # [9, C] tap-major
    xh, dyh = x.detach().permute(0, 2, 3, 1), dy.permute(0, 2, 3, 1)
    y = torch.zeros(n, oh_n, ow_n, c, dtype=torch.float64)
    dx = torch.zeros(n, hw, hw, c, dtype=torch.float64)
    dw = torch.zeros(9, c, dtype=torch.float64)
    for b in range(n):
        for oh in range(oh_n):
            for ow in range(ow_n):
                for kh in range(3):
                    ih = oh * stride - pad + kh
                    if ih < 0 or ih >= hw:
                        continue
                    for kw in range(3):
                        iw = ow * stride - pad + kw
                        if iw < 0 or iw >= hw:
                            continue
                        y[b, oh, ow] += xh[b, ih, iw] * wt[kh * 3 + kw]
                        dw[kh * 3 + kw] += dyh[b, oh, ow] * xh[b, ih, iw]
        for ih in range(hw):
            for iw in range(hw):
                for kh in range(3):
                    th = ih + pad - kh
                    if th < 0 or th % stride or th // stride >= oh_n:
                        continue
                    for kw in range(3):
                        tw = iw + pad - kw
                        if tw < 0 or tw % stride or tw // stride >= ow_n:
                            continue
                        dx[b, ih, iw] += dyh[b, th // stride, tw // stride] * wt[kh * 3 + kw]
    torch.testing.assert_close(y.permute(0, 3, 1, 2), y_ref.detach())
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), x.grad)
    torch.testing.assert_close(dw.t().reshape(c, 1, 3, 3), w.grad)


def test_every_zoo_model_compiles_to_fused_ops_only():
    """All reference architectures x data sets (and the extra zoo) map onto native engine ops:
    nothing falls back to the autograd ``ModuleOp``."""
    from ddlbench_b200.models.extra import EXTRA_ARCHS
    from ddlbench_b200.models.zoo import ARCHS, DATASETS
    for ds in DATASETS:
        for arch in ARCHS:
            plan = compile_plan(build_model(arch, ds))
            kinds = {op.kind for op in plan.ops}
            assert "module" not in kinds, (arch, ds, kinds)
            assert plan.ops[0].kind in ("conv_bn_act", "conv_bias_act"), (arch, ds)
    for arch in EXTRA_ARCHS:
        kinds = {op.kind for op in compile_plan(build_model(arch, "imagenet")).ops}
        assert "module" not in kinds, (arch, kinds)


def test_bn_fold_plan_and_executor_equivalence(monkeypatch):
    """DDLB_BN_FOLD: BatchNorm applications deferred into the consuming 1x1-convolution GEMM.  On
    the CPU the consumer materialises the value with the reference op, so this checks the plan
    marking, the pending-value bookkeeping and the backward against autograd."""
    from ddlbench_b200.engine import ops as E
    from ddlbench_b200.engine.plan import compile_plan as cp
    monkeypatch.setattr(E, "BN_FOLD", True)
    torch.manual_seed(0)
    m = build_model("resnet50", "cifar10").double().train()
    plan = cp(m)
    deferred = [op for op in plan.ops if getattr(op, "defer_apply", False)]
    folding = [op for op in plan.ops if getattr(op, "folds_input", False)]
    # per bottleneck: conv2's BN feeds conv3 (1x1), the block output feeds the next conv1 (1x1);
    # conv1's BN feeds the 3x3 and the last block feeds the pooling head: not deferred; the CIFAR
    # stem (no max-pool) feeds layer1.0.conv1 directly: deferred
    assert len(deferred) == len(folding) == 16 + 15 + 1
    assert all(op.wshape[2:] == (1, 1) for op in folding)
    ex = StageExecutor(plan, device="cpu", dtype=torch.float64, lr=0.1, momentum=0.9, weight_decay=1e-4)
    for name, p in m.named_parameters():
        ex.store.weights()[name].copy_(p.detach())
    x = torch.randn(4, 3, 32, 32, dtype=torch.float64)
    t = torch.randint(0, 10, (4,))
    y = m(x)
    F.cross_entropy(y, t).backward()
    outs, ctx = ex.forward([ex.prepare_input(x)])
    assert not E.PENDING_FOLD                                   # every deferred value was written
    _, d = ex.loss(outs[0], t)
    ex.backward(ctx, [d])
    assert torch.allclose(outs[0], y, atol=1e-9)
    assert _global_rel_err(ex, m) < 1e-6
    # a stage cut right after a deferred producer: the value must be written before it leaves
    cut = next(i for i, n in enumerate(m.nodes) if n.name == "layer1.0.relu2") + 1
    ex_a = StageExecutor(cp(m, 0, cut), device="cpu", dtype=torch.float64)
    ex_b = StageExecutor(cp(m, cut, None), device="cpu", dtype=torch.float64, first_stage=False)
    for name, p in m.named_parameters():
        for e in (ex_a, ex_b):
            if name in e.store.weights():
                e.store.weights()[name].copy_(p.detach())
    mid, _ = ex_a.forward([ex_a.prepare_input(x)])
    out_b, _ = ex_b.forward(mid)
    assert torch.allclose(out_b[0], y, atol=1e-9) and not E.PENDING_FOLD
