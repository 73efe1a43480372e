"""The static engine on the GPU (bf16, channels-last, native kernels) against the same
engine on CPU (fp32 reference ops), plus CUDA-graph replay of the whole training step."""
import pytest
import torch

from ddlbench_b200 import ops
from ddlbench_b200.engine import StageExecutor
from ddlbench_b200.harness import Trainer
from ddlbench_b200.models import build_model

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12)).item()


@pytest.mark.parametrize("arch,ds,tc", [("resnet18", "cifar10", False), ("resnet50", "cifar10", True),
                                         ("vgg11", "cifar10", False), ("mobilenetv2", "cifar10", True)])
def test_engine_gpu_matches_cpu(arch, ds, tc):
    torch.manual_seed(0)
    m = build_model(arch, ds)
    cpu = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
    gpu = StageExecutor.from_model(m, device="cuda", dtype=torch.bfloat16, tcgen05=tc)
    x = torch.randn(16, *m.input_shape)
    t = torch.randint(0, 10, (16,))
    oc, cc = cpu.forward([cpu.prepare_input(x)])
    sc, dc = cpu.loss(oc[0], t)
    cpu.backward(cc, [dc])
    before = ops.launches
    og, cg = gpu.forward([gpu.prepare_input(x)])
    sg, dg = gpu.loss(og[0], t.cuda())
    gpu.backward(cg, [dg])
    torch.cuda.synchronize()
    assert ops.launches > before, "native kernels did not run"
    assert abs(sc[0].item() - sg[0].item()) < 0.08 * max(1.0, sc[0].item())
    # A random-init deep BatchNorm net amplifies bf16 rounding (ReLU masks flip), so the
    # yardstick is stock PyTorch bf16 autograd on the same weights: our kernels must be no
    # further from the fp32 gradients than the library path is.
    import torch.nn.functional as F
    mb = build_model(arch, ds)
    mb.load_state_dict(m.state_dict())
    mb = mb.cuda().to(torch.bfloat16).to(memory_format=torch.channels_last).train()
    yb = mb(x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    F.cross_entropy(yb.float(), t.cuda()).backward()
    tb = dict(mb.named_parameters())
    num = num_t = den = 0.0
    for name in cpu.store.layout.entries:
        r = cpu.store.g_views[name]
        a = gpu.store.g_views[name].float().cpu()
        b = tb[name].grad.float().cpu()
        num += (a - r).pow(2).sum().item()
        num_t += (b - r).pow(2).sum().item()
        den += r.pow(2).sum().item()
    err, err_t = (num / den) ** 0.5, (num_t / den) ** 0.5
    assert err < 1.25 * err_t + 0.02, (err, err_t)


def test_trainer_cuda_graph_matches_eager_and_learns():
    losses = {}
    for graph in (False, True):
        tr = Trainer("resnet18", "cifar10", batch_size=32, device="cuda", dtype=torch.bfloat16,
                     lr=0.05, momentum=0.9, weight_decay=1e-4, cuda_graph=graph, seed=3)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(32, 3, 32, 32, generator=g).cuda()
        t = torch.randint(0, 10, (32,), generator=g).cuda()
        ls = []
        for _ in range(12):
            ls.append(float(tr.step(x, t)[0]))
        losses[graph] = ls
    # same batch every step: the loss must fall, with and without graph replay
    assert losses[False][-1] < losses[False][0] * 0.7
    assert losses[True][-1] < losses[True][0] * 0.7
    assert all(l == l for l in losses[True])
    # the warm-up steps inside the capture are undone (state snapshot / restore): the graph run
    # starts from the same weights as the eager run, so the first losses agree
    assert losses[True][0] == pytest.approx(losses[False][0], rel=2e-2)
    assert losses[True][1] == pytest.approx(losses[False][1], rel=0.1)


def test_trainer_resnet50_imagenet_step_u8():
    from ddlbench_b200.data.synthetic import SyntheticImages
    # small LR: with lr=0.1 a single repeated batch of 8 transiently blows the loss up
    # (observed identically with and without our kernels) before it collapses
    tr = Trainer("resnet50", "imagenet", batch_size=8, device="cuda", dtype=torch.bfloat16,
                 cuda_graph=True, tcgen05=True, lr=0.005, momentum=0.9)
    src = SyntheticImages("imagenet", 8, num_batches=2, device="cuda")
    losses = []
    for i in range(12):
        x, t = src.batch(0)
        losses.append(float(tr.step_u8(x, t, src.mean, src.std)[0]))
    assert all(l == l for l in losses) and min(losses[-3:]) < losses[0]
    assert tr.launches_per_step > 100


def test_resnet50_imagenet_gradients_against_fp32():
    """ResNet-50 at the benchmark's input shape (3x224x224): every parameter gradient of the bf16
    engine (tcgen05 convolutions, fused BatchNorm statistics, split-K / strided data gradients,
    one-launch weight gradients) against the same engine in fp32 on the CPU.  Yardstick as above:
    no further from fp32 than stock PyTorch bf16 autograd on the same weights."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    m = build_model("resnet50", "imagenet")
    n = 8
    x = torch.randn(n, *m.input_shape)
    t = torch.randint(0, 1000, (n,))
    cpu = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
    oc, cc = cpu.forward([cpu.prepare_input(x)])
    sc, dc = cpu.loss(oc[0], t)
    cpu.backward(cc, [dc])
    gpu = StageExecutor.from_model(m, device="cuda", dtype=torch.bfloat16, tcgen05=True)
    before = ops.launches
    og, cg = gpu.forward([gpu.prepare_input(x)])
    sg, dg = gpu.loss(og[0], t.cuda())
    gpu.backward(cg, [dg])
    torch.cuda.synchronize()
    assert ops.launches - before > 200
    assert abs(sc[0].item() - sg[0].item()) < 0.05 * max(1.0, sc[0].item())
    mb = build_model("resnet50", "imagenet")
    mb.load_state_dict(m.state_dict())
    mb = mb.cuda().to(torch.bfloat16).to(memory_format=torch.channels_last).train()
    yb = mb(x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    F.cross_entropy(yb.float(), t.cuda()).backward()
    tb = dict(mb.named_parameters())
    import collections
    acc = collections.defaultdict(lambda: [0.0, 0.0, 0.0])      # group -> [ours^2, torch^2, ref^2]
    for name in cpu.store.layout.entries:
        r = cpu.store.g_views[name]
        a = gpu.store.g_views[name].float().cpu()
        b = tb[name].grad.float().cpu()
        part = name.split(".")[1] if name.count(".") >= 2 else name.split(".")[0]
        grp = part.split("_")[0] if part.startswith("layer") else "stem/fc"
        for key in (grp, "all"):
            acc[key][0] += (a - r).pow(2).sum().item()
            acc[key][1] += (b - r).pow(2).sum().item()
            acc[key][2] += r.pow(2).sum().item()
    report = {k: (round((v[0] / v[2]) ** 0.5, 4), round((v[1] / v[2]) ** 0.5, 4)) for k, v in sorted(acc.items())}
    print("resnet50 224x224 gradient error vs fp32, per stage (ours, torch bf16 autograd):", report)
    # batch 8 at random init is an ill-conditioned point (bf16 BatchNorm noise is amplified through
    # 50 layers: stock PyTorch bf16 is itself > 100 % off in the early stages) — the claim is that
    # the engine is no worse than stock bf16 autograd, stage by stage
    for k, (err, err_t) in report.items():
        assert err < 1.25 * err_t + 0.02, (k, err, err_t)
