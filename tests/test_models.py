import pytest
import torch

from ddlbench_b200.models import ARCHS, build_model, zoo


@pytest.mark.parametrize("arch", ["resnet18", "resnet50", "vgg11", "vgg16", "mobilenetv2"])
@pytest.mark.parametrize("ds", ["mnist", "cifar10"])
def test_forward_shapes_small(arch, ds):
    m = build_model(arch, ds).eval()
    with torch.no_grad():
        y = m(torch.randn(2, *m.input_shape))
    assert y.shape == (2, 10)


def test_imagenet_param_counts_match_torchvision():
    import torchvision.models as tv
    for a, f in [("resnet18", tv.resnet18), ("resnet50", tv.resnet50), ("vgg11", tv.vgg11),
                 ("mobilenetv2", tv.mobilenet_v2)]:
        n1 = sum(p.numel() for p in f().parameters())
        n2 = sum(p.numel() for p in build_model(a, "imagenet").parameters())
        assert n1 == n2, a


def test_resnet152_node_count_and_factory_names():
    m = zoo.inet_resnet152()
    assert len(m.nodes) == 515 and m.num_classes == 1000
    assert zoo.mnist_vgg16().input_shape == (1, 28, 28)
    assert zoo.cifar10_mobilenetv2().num_classes == 10
    assert zoo.highres_resnet50().input_shape == (3, 512, 512)


def test_sequential_and_segment_views_equal_plain():
    torch.manual_seed(0)
    m = build_model("resnet50", "cifar10").eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        y = m(x)
        assert torch.equal(y, m.as_sequential()(x))
        # cut INSIDE a residual block: the skip tensor must travel with the activation
        cut = next(i for i, n in enumerate(m.nodes) if n.name == "layer2.0.bn2") + 1
        s1, s2 = m.segment(0, cut), m.segment(cut, len(m.nodes))
        assert len(s1.out_names) == 2
        o = s1(x)
        assert torch.equal(y, s2(*o))


def test_block_boundaries_are_single_tensor_cuts():
    m = build_model("mobilenetv2", "cifar10")
    for i in m.block_boundaries():
        assert len(m.live_after(i)) == 1


def test_vgg_mnist_drops_last_pool():
    a = build_model("vgg16", "mnist")
    b = build_model("vgg16", "cifar10")
    assert len(a.nodes) == len(b.nodes) - 1
