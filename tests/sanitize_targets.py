"""Small launches of every round-2 kernel path for compute-sanitizer (memcheck / racecheck /
synccheck / initcheck): implicit GEMM (one-tile variant, chunked variant, cluster split-K with
the L2 or (DDLB_IG_SPLIT=dsmem) distributed-shared-memory exchange, MN-major data gradient, the
one-launch stride-2 parity classes, fused finalize ticket), the CTA-pair GEMM, weight gradients (direct store, planes +
reduce) and the cluster BatchNorm kernels (DSMEM).  Sizes are chosen to hit each variant while staying
small enough for the ~50x slowdown of the tools."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402

DEV, BF, CL = "cuda", torch.bfloat16, torch.channels_last


def conv_case(n, ci, co, hw, k, s, p):
    x = (torch.randn(n, ci, hw, hw, device=DEV)).to(BF).contiguous(memory_format=CL)
    w = (torch.randn(co, ci, k, k, device=DEV) * 0.05).to(BF).contiguous(memory_format=CL)
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device=DEV)
    stat = torch.zeros(2 * co, device=DEV)
    tk = torch.zeros((co + 63) // 64, dtype=torch.int32, device=DEV)
    y = ops.conv_fprop(x, w, s, [p, p], col_sums=sums, bn_finalize=(tk, stat, None, None, 1e-5, 0.1))
    dy = torch.randn_like(y)
    dx = ops.conv_dgrad(dy, w, x.shape, s, [p, p])
    gw = torch.zeros((co, k, k, ci), device=DEV, dtype=BF).permute(0, 3, 1, 2)
    assert ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)
    y_ref = F.conv2d(x.float(), w.float(), None, s, p)
    torch.cuda.synchronize()
    err = (y.float() - y_ref).abs().max().item()
    assert err < 0.05 * max(1.0, y_ref.abs().max().item()), err
    return dx


def main():
    torch.manual_seed(0)
    conv_case(4, 64, 64, 16, 1, 1, 0)        # one-tile variant (1 k-block), stats, separate finalize
    conv_case(4, 128, 256, 12, 1, 1, 0)      # one-tile variant, 2 k-blocks, 2 n-tiles
    conv_case(2, 64, 64, 12, 3, 1, 1)        # chunked variant, fused finalize ticket, conv wgrad planes
    conv_case(2, 512, 128, 7, 3, 1, 1)       # few tiles + long K: cluster split-K, 4 slices (32-column stores)
    conv_case(8, 256, 128, 7, 1, 1, 0)       # cluster split-K, 2 slices
    conv_case(2, 64, 128, 12, 3, 2, 1)       # stride 2: fprop element strides, dgrad parity views
    conv_case(2, 128, 256, 12, 1, 2, 0)      # 1x1 stride 2: dgrad with memset + one parity class
    # FC layers
    a, w = torch.randn(8, 256, device=DEV).to(BF), (torch.randn(104, 256, device=DEV) * 0.05).to(BF)
    y = ops.gemm_ig(a, w, bias=torch.randn(104, device=DEV).to(BF), relu=True)
    dy = torch.randn_like(y)
    ops.gemm_ig(dy, w, b_mn=True)
    g = torch.zeros((104, 256), device=DEV, dtype=BF)
    assert ops.wgrad_gemm_into(dy, a, g, False)
    # CTA-pair (cta_group::2) GEMM: both BLOCK_N variants, ragged M / N / K, bias + ReLU
    for m, n, k in ((300, 128, 136), (520, 328, 200)):
        a2, w2 = torch.randn(m, k, device=DEV).to(BF), (torch.randn(n, k, device=DEV) * 0.05).to(BF)
        y2 = ops.gemm_pair(a2, w2, bias=torch.randn(n, device=DEV).to(BF), relu=True)
        ref2 = (a2.float() @ w2.float().t()).relu()
        assert y2 is not None
        del ref2
    # element-wise layers: patch matrix of the stem, average pools, Philox dropout
    xi = torch.randn(2, 3, 33, 33, device=DEV).to(BF).contiguous(memory_format=CL)
    ops.im2col_rows(xi, 7, 7, 2, 3)
    xp = torch.randn(2, 64, 13, 13, device=DEV).to(BF).contiguous(memory_format=CL)
    yp = ops.avgpool_fwd(xp, (6, 6), 0, 0, 0, True)
    ops.avgpool_bwd(torch.randn_like(yp), xp.shape, 0, 0, 0, True)
    yq = ops.avgpool_fwd(xp, (7, 7), 3, 2, 1, False, True)
    ops.avgpool_bwd(torch.randn_like(yq), xp.shape, 3, 2, 1, False, True)
    xd = torch.randn(4, 1024, device=DEV).to(BF)
    yd, md = ops.dropout_fwd(xd, 0.5)
    ops.dropout_bwd(torch.randn_like(yd), md, 0.5)
    # cluster BatchNorm (thread-block clusters + DSMEM), forward and backward
    x = torch.randn(8, 64, 14, 14, device=DEV).to(BF).contiguous(memory_format=CL)
    gamma, beta = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)
    rm, rv = torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
    ws = ops.BnWorkspace(64, DEV)
    mask = torch.empty(x.numel() // 8, dtype=torch.uint8, device=DEV)
    yb, st = ops.bn_forward(x, None, gamma, beta, rm, rv, 1e-5, 0.1, ops.ACT_RELU, ws, mask_out=mask)
    dg, db = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    ops.bn_backward(torch.randn_like(yb), None, yb, x, st, gamma, ops.ACT_RELU, False, dg, db, False, ws, mask=mask)
    torch.cuda.synchronize()
    print("sanitize targets done")


if __name__ == "__main__":
    main()
