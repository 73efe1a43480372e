"""Small launches of every round-2 kernel path for compute-sanitizer (memcheck / racecheck /
synccheck / initcheck): implicit GEMM (one-tile variant, chunked variant, split-K slots, MN-major
data gradient, stride-2 parity views, fused finalize ticket), weight gradients (direct store, planes +
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
    conv_case(2, 512, 128, 7, 3, 1, 1)       # few tiles + long K: split-K slots
    conv_case(2, 64, 128, 12, 3, 2, 1)       # stride 2: fprop element strides, dgrad parity views
    conv_case(2, 128, 256, 12, 1, 2, 0)      # 1x1 stride 2: dgrad with memset + one parity class
    # FC layers
    a, w = torch.randn(8, 256, device=DEV).to(BF), (torch.randn(104, 256, device=DEV) * 0.05).to(BF)
    y = ops.gemm_ig(a, w, bias=torch.randn(104, device=DEV).to(BF), relu=True)
    dy = torch.randn_like(y)
    ops.gemm_ig(dy, w, b_mn=True)
    g = torch.zeros((104, 256), device=DEV, dtype=BF)
    assert ops.wgrad_gemm_into(dy, a, g, False)
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
