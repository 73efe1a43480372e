"""GNMT on the GPU: the emulated bidirectional encoder layer and the attention mask run on the
native revert_varlen / set_mask kernels; logits and gradients must match the CPU (reference
ops) model.  (Runs last: file name sorts after the kernel tests.)"""
import pytest
import torch

from ddlbench_b200 import ops
from ddlbench_b200.models.gnmt import GNMT, PAD, LabelSmoothing

pytestmark = pytest.mark.gpu


def test_gnmt_gpu_matches_cpu(monkeypatch):
    # exact fp32 on the device side too (cuDNN's LSTM would otherwise use TF32)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(0)
    m = GNMT(vocab_size=64, hidden_size=32, num_layers=4, dropout=0.0)
    T, B = 9, 4
    lens = torch.tensor([9, 7, 4, 2])
    src = torch.randint(4, 64, (T, B))
    for b, l in enumerate(lens):
        src[l:, b] = PAD
    tgt = torch.randint(4, 64, (T + 1, B))
    crit = LabelSmoothing(PAD, 0.1)
    ref = m(src, lens, tgt[:-1])
    crit(ref, tgt[1:]).backward()
    g_ref = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    mg = m.cuda()
    before = ops.launches
    out = mg(src.cuda(), lens.cuda(), tgt[:-1].cuda())
    assert ops.launches > before, "native sequence kernels did not run"
    crit(out, tgt[1:].cuda()).backward()
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-2, atol=5e-3)
    for n, p in mg.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), g_ref[n], rtol=2e-2, atol=1e-2, msg=n)
