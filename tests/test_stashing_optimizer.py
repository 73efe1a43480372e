"""nn.Module weight-stashing optimizers (API of pipedream-fork/runtime/{optimizer,sgd,adam}.py).
Semantic invariant of the reference's tests/backprop/sgd_with_stashing.py: the gradient w.r.t.
the input of minibatch b equals the gradient computed with the weight version b's forward saw,
for the first `num_versions` in-flight minibatches."""
import copy

import pytest
import torch
import torch.nn as nn

from ddlbench_b200.parallel.pipedream.optimizer import (AdamWithWeightStashing,
                                                        SGDWithWeightStashing)


def _run(num_versions, opt_cls=SGDWithWeightStashing, **kw):
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    w0 = copy.deepcopy(model.state_dict())
    opt = opt_cls([model], model.parameters(), num_versions=num_versions, lr=0.5, **kw)
    xs = [torch.ones(2, 8, requires_grad=True) for _ in range(3)]
    outs = [model(x).sum() for x in xs]          # three forwards with the SAME weights
    grads = []
    for x, o in zip(xs, outs):
        opt.zero_grad()
        opt.load_old_params()
        o.backward()
        opt.load_new_params()
        opt.step()
        grads.append(x.grad.clone())
    changed = any(not torch.equal(model.state_dict()[k], w0[k]) for k in w0)
    return grads, changed


def test_sgd_stashing_semantics():
    g1, c1 = _run(1)
    g2, c2 = _run(2)
    g3, c3 = _run(3)
    assert c1 and c2 and c3
    assert not torch.allclose(g1[0], g1[1])
    assert torch.allclose(g2[0], g2[1]) and not torch.allclose(g2[1], g2[2])
    assert torch.allclose(g3[0], g3[1]) and torch.allclose(g3[1], g3[2])
    assert torch.allclose(g1[0], g3[0])


def test_stashed_sgd_equals_plain_sgd_with_one_version():
    torch.manual_seed(1)
    a = nn.Sequential(nn.Linear(6, 6), nn.Tanh(), nn.Linear(6, 2))
    b = copy.deepcopy(a)
    oa = SGDWithWeightStashing([a], a.parameters(), num_versions=1, lr=0.1, momentum=0.9,
                               weight_decay=1e-3)
    ob = torch.optim.SGD(b.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-3)
    x = torch.randn(5, 6)
    for _ in range(4):
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6)


def test_adam_stashing_and_macrobatch():
    g, changed = _run(2, AdamWithWeightStashing)
    assert changed and torch.allclose(g[0], g[1])
    torch.manual_seed(0)
    model = nn.Linear(4, 2)
    opt = SGDWithWeightStashing([model], model.parameters(), num_versions=4, lr=0.1, macrobatch=True)
    assert opt.num_versions == 2 and opt.update_interval == 4
    w0 = model.weight.detach().clone()
    for i in range(4):
        opt.zero_grad()
        opt.load_old_params()
        model(torch.ones(1, 4)).sum().backward()
        opt.load_new_params()
        opt.step()
        if i < 3:
            assert torch.equal(model.weight.detach(), w0)     # no update before the 4th minibatch
    assert not torch.equal(model.weight.detach(), w0)
