"""torchgpipe skip API (skippable / stash / pop / Namespace) and its conversion to a layer graph —
the way the reference writes its pipeline-ready ResNets
(benchmark/imagenet/gpipemodels/resnet/block.py:31-74)."""
from collections import OrderedDict

import pytest
import torch
import torch.nn as nn

from ddlbench_b200.engine import StageExecutor
from ddlbench_b200.models.graph import INPUT
from ddlbench_b200.parallel.gpipe import GPipe
from ddlbench_b200.parallel.skip import (Namespace, Skippable, pop, sequential_to_graph, skippable, stash,
                                         verify_skippables)


@skippable(stash=['identity'])
class Identity(nn.Module):
    def forward(self, tensor):
        yield stash('identity', tensor)
        return tensor


@skippable(pop=['identity'])
class Residual(nn.Module):
    def __init__(self, downsample=None):
        super().__init__()
        self.downsample = downsample

    def forward(self, input):
        identity = yield pop('identity')
        if self.downsample is not None:
            identity = self.downsample(identity)
        return input + identity


def basicblock(inplanes, planes, stride=1, downsample=None):
    """The reference's basicblock, verbatim in structure."""
    layers = OrderedDict()
    ns = Namespace()
    layers['identity'] = Identity().isolate(ns)
    layers['conv1'] = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
    layers['bn1'] = nn.BatchNorm2d(planes)
    layers['relu1'] = nn.ReLU()
    layers['conv2'] = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
    layers['bn2'] = nn.BatchNorm2d(planes)
    layers['residual'] = Residual(downsample).isolate(ns)
    layers['relu3'] = nn.ReLU()
    return nn.Sequential(layers)


def tiny_resnet():
    down = nn.Sequential(nn.Conv2d(8, 16, 1, 2, bias=False), nn.BatchNorm2d(16))
    return nn.Sequential(OrderedDict([
        ("stem", nn.Conv2d(3, 8, 3, 1, 1, bias=False)), ("bn", nn.BatchNorm2d(8)), ("relu", nn.ReLU()),
        ("layer1", nn.Sequential(basicblock(8, 8), basicblock(8, 8))),
        ("layer2", nn.Sequential(basicblock(8, 16, 2, down))),
        ("pool", nn.AdaptiveAvgPool2d((1, 1))), ("flat", nn.Flatten()), ("fc", nn.Linear(16, 10))]))


def _manual_forward(seq, x):
    """Reference semantics: run the Sequential with a skip tracker."""
    tracker = {}
    for m in [m for _, m in __import__("ddlbench_b200.parallel.skip", fromlist=["_flatten"])._flatten(seq)]:
        if isinstance(m, Skippable):
            popped = {n: tracker.pop(m.key(n)) for n in m.poppable_names}
            x, st = m.dispatch(x, popped)
            for n, t in st.items():
                tracker[m.key(n)] = t
        else:
            x = m(x)
    assert not tracker
    return x


def test_graph_equals_sequential_with_skip_tracker():
    torch.manual_seed(0)
    seq = tiny_resnet().eval()
    x = torch.randn(4, 3, 16, 16)
    g = sequential_to_graph(seq, x[:1]).eval()
    assert torch.allclose(g(x), _manual_forward(seq, x), atol=1e-6)
    # the skips became ordinary second input edges (no stash node: Identity stashes its input)
    res = [n for n in g.nodes if n.name.endswith("residual")]
    assert len(res) == 3 and all(len(n.inputs) == 2 for n in res)
    assert res[0].inputs[1] == "relu"                     # block 1 pops the stem output
    assert g.nodes[-1].name == "fc" and g.nodes[0].inputs == (INPUT,)
    # a cut in the middle of a block carries the skip tensor explicitly
    cut = [i for i, n in enumerate(g.nodes) if n.name == "layer1.0.bn1"][0]
    assert set(g.live_after(cut)) == {"layer1.0.bn1", "relu"}


def test_verify_skippables_rules():
    ns = Namespace()
    with pytest.raises(TypeError, match="no earlier layer stashes"):
        verify_skippables(nn.Sequential(Residual().isolate(ns)))
    with pytest.raises(TypeError, match="never popped"):
        verify_skippables(nn.Sequential(Identity().isolate(ns)))
    with pytest.raises(TypeError, match="already stashed"):
        verify_skippables(nn.Sequential(Identity().isolate(ns), Identity().isolate(ns), Residual().isolate(ns)))
    # different namespaces do not collide
    a, b = Namespace(), Namespace()
    verify_skippables(nn.Sequential(Identity().isolate(a), Identity().isolate(b), Residual().isolate(b),
                                    Residual().isolate(a)))


def test_stash_of_an_intermediate_tensor():
    @skippable(stash=['mid'])
    class Mid(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(6, 6), nn.Linear(6, 6)

        def forward(self, x):
            h = torch.relu(self.a(x))
            yield stash('mid', h)
            return self.b(h)

    @skippable(pop=['mid'])
    class Join(nn.Module):
        def forward(self, x):
            h = yield pop('mid')
            return x * h

    torch.manual_seed(1)
    seq = nn.Sequential(Mid(), nn.Tanh(), Join())
    x = torch.randn(3, 6)
    g = sequential_to_graph(seq, x[:1])
    assert any(".stash.mid" in n.name for n in g.nodes)
    assert torch.allclose(g(x), _manual_forward(seq, x), atol=1e-6)


def test_gpipe_accepts_a_skippable_sequential_and_matches_autograd():
    """GPipe(nn.Sequential with stash/pop, ...) partitions the converted graph ANYWHERE (here in the
    middle of residual blocks) and its update equals plain large-batch SGD."""
    torch.manual_seed(0)
    seq = tiny_resnet()
    x, t = torch.randn(8, 3, 16, 16), torch.randint(0, 10, (8,))
    graph = sequential_to_graph(seq, x[:1])
    n = len(graph.nodes)
    pipe = GPipe(seq, [5, 9, n - 14], chunks=1, checkpoint="never", devices=["cpu"] * 3, dtype=torch.float32,
                 lr=0.1, momentum=0.0, sample_input=x[:1])
    before = {k: v.clone() for k, v in pipe.state_dict().items()}
    pipe.train_step(x, t)
    after = pipe.state_dict()
    # reference: autograd on the same graph module
    ref = sequential_to_graph(seq, x[:1])
    ref.load_state_dict(graph.state_dict())
    loss = nn.functional.cross_entropy(ref(x), t)
    grads = torch.autograd.grad(loss, list(ref.parameters()))
    changed = 0
    for (k, p), g_ in zip(ref.named_parameters(), grads):
        key = [kk for kk in after if kk.replace(".", "_").endswith(k.split("layers.")[-1].replace(".", "_"))]
        if not key:
            continue
        exp = before[key[0]] - 0.1 * g_
        assert torch.allclose(after[key[0]], exp, atol=2e-5), k
        changed += 1
    assert changed >= 10


def test_executor_runs_converted_model():
    torch.manual_seed(2)
    seq = tiny_resnet()
    x, t = torch.randn(4, 3, 16, 16), torch.randint(0, 10, (4,))
    g = sequential_to_graph(seq, x[:1])
    ex = StageExecutor.from_model(g, device="cpu", dtype=torch.float32, lr=0.05, momentum=0.9)
    losses = []
    for _ in range(5):
        o, c = ex.forward([x])
        s, d = ex.loss(o[0], t)
        ex.backward(c, [d])
        ex.store.sgd_step()
        losses.append(float(s[0]))
    assert losses[-1] < losses[0]
