"""GPipe: micro-batched fill-drain pipeline == large-batch SGD; balance helpers."""
import glob
import os
import subprocess
import sys

import pytest
import torch

from ddlbench_b200.engine import StageExecutor
from ddlbench_b200.models import build_model
from ddlbench_b200.parallel.gpipe import (GPipe, balance_by_size, balance_by_time,
                                          solve_min_max_partition)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_min_max_partition():
    assert solve_min_max_partition([1, 1, 1, 1], 2) == [2, 2]
    assert solve_min_max_partition([5, 1, 1, 1, 1, 1], 2) == [1, 5]
    assert sum(solve_min_max_partition([3, 1, 4, 1, 5, 9, 2, 6], 3)) == 8


def test_balance_helpers_cover_all_layers():
    m = build_model("vgg11", "cifar10")
    b = balance_by_size(3, m)
    assert len(b) == 3 and sum(b) == len(m.nodes)
    bt = balance_by_time(2, m, torch.randn(4, 3, 32, 32), device="cpu", trials=1)
    assert len(bt) == 2 and sum(bt) == len(m.nodes) and min(bt) >= 1


@pytest.mark.parametrize("checkpoint", ["except_last", "never", "always"])
def test_gpipe_equals_large_batch_sgd_single_process(checkpoint):
    """No-BatchNorm model (VGG): 4 micro-batches through 3 partitions == one SGD step on the
    whole mini-batch."""
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist")
    ref = StageExecutor.from_model(m, device="cpu", dtype=torch.float32, lr=0.1, momentum=0.9,
                                   weight_decay=1e-4)
    pipe = GPipe(m, balance_by_size(3, m), chunks=4, checkpoint=checkpoint,
                 devices=["cpu"] * 3, dtype=torch.float32, lr=0.1, momentum=0.9,
                 weight_decay=1e-4)
    x = torch.randn(16, 1, 28, 28)
    t = torch.randint(0, 10, (16,))
    for _ in range(2):
        o, c = ref.forward([x])
        s, d = ref.loss(o[0], t)
        ref.backward(c, [d])
        ref.store.sgd_step()
        sp = pipe.train_step(x, t)
        assert abs(float(sp[0]) - float(s[0])) < 1e-4
    sd_ref, sd = ref.state_dict(), pipe.state_dict()
    assert set(sd) == set(sd_ref)
    for k in sd:
        assert torch.allclose(sd[k], sd_ref[k], atol=1e-5, rtol=1e-4), k


def test_gpipe_with_batchnorm_and_skip_cut_runs():
    """ResNet cut inside residual blocks: skip tensors cross partitions, BatchNorm statistics
    are per micro-batch, re-computation must not double-update the running statistics."""
    torch.manual_seed(0)
    m = build_model("resnet18", "cifar10")
    n = len(m.nodes)
    pipe = GPipe(m, [11, 20, n - 31], chunks=2, devices=["cpu"] * 3, dtype=torch.float32,
                 lr=0.05, momentum=0.9)
    x = torch.randn(8, 3, 32, 32)
    t = torch.randint(0, 10, (8,))
    l0 = float(pipe.train_step(x, t)[0])
    rm = pipe.state_dict()["layers.bn1.running_mean"].clone()
    # one step with momentum 0.1 from zeros: |running_mean| = 0.1 * (sum of 2 micro-batch means) roughly;
    # a double update during recompute would make it ~2x larger than the no-checkpoint run
    pipe2 = GPipe(m, [11, 20, n - 31], chunks=2, checkpoint="never", devices=["cpu"] * 3,
                  dtype=torch.float32, lr=0.05, momentum=0.9)
    pipe2.train_step(x, t)
    rm2 = pipe2.state_dict()["layers.bn1.running_mean"]
    assert torch.allclose(rm, rm2, atol=1e-6)
    for _ in range(5):
        l1 = float(pipe.train_step(x, t)[0])
    assert l1 < l0
    st = pipe.eval_step(x, t)
    assert st.shape == (3,)


def test_gpipe_multiprocess_matches_single_process(tmp_path):
    env = dict(os.environ, GP_OUT=str(tmp_path), OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
           "--master-addr", "127.0.0.1", "--master-port", "29741",
           os.path.join(ROOT, "tests", "gpipe_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=420)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    res = [torch.load(f) for f in sorted(glob.glob(os.path.join(str(tmp_path), "rank*.pt")))]
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist")
    pipe = GPipe(m, balance_by_size(3, m), chunks=4, devices=["cpu"] * 3, dtype=torch.float32,
                 lr=0.1, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(16, 1, 28, 28, generator=g)
    t = torch.randint(0, 10, (16,), generator=g)
    losses = [float(pipe.train_step(x, t)[0]) for _ in range(3)]
    last = [r for r in res if r["losses"]][0]
    assert last["losses"] == pytest.approx(losses, rel=1e-3, abs=1e-3)
    sd = pipe.state_dict()
    for r in res:
        for k, v in r["state"].items():
            assert torch.allclose(v, sd[k], atol=1e-3, rtol=1e-2), k


def test_checkpoint_recompute_reuses_dropout_masks():
    """torchgpipe restores the RNG state around a recompute; without it a checkpointed VGG (two
    Dropout layers in the classifier) would back-propagate through different masks than its
    forward.  checkpoint='always' must give the same update as checkpoint='never'."""
    import torch
    from ddlbench_b200.models import build_model
    from ddlbench_b200.parallel.gpipe import GPipe, balance_by_size
    torch.manual_seed(0)
    m = build_model("vgg11", "cifar10")
    x, t = torch.randn(8, 3, 32, 32), torch.randint(0, 10, (8,))
    out = []
    for mode in ("never", "always"):
        torch.manual_seed(1)
        pipe = GPipe(m, balance_by_size(2, m), chunks=2, checkpoint=mode, devices=["cpu"] * 2,
                     dtype=torch.float32, lr=0.1, momentum=0.0)
        torch.manual_seed(7)
        pipe.train_step(x, t)
        out.append(torch.cat([ex.store.master.clone() for ex in pipe.execs.values()]))
    assert torch.allclose(out[0], out[1], atol=1e-6), (out[0] - out[1]).abs().max()
