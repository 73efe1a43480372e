"""Real-data (JPEG tree) input mode and the synthetic-tree generator."""
import os

import numpy as np
import torch

from ddlbench_b200.config import RunConfig
from ddlbench_b200.data.imagefolder import JpegFolder, scan_tree
from ddlbench_b200.data.synthetic import generate_jpeg_tree
from ddlbench_b200.harness.loop import run


def test_jpeg_tree_roundtrip_and_sharding(tmp_path):
    generate_jpeg_tree(str(tmp_path), "cifar10", train=80, val=20, workers=1, verbose=False)
    samples, classes = scan_tree(os.path.join(tmp_path, "CIFAR10", "train"))
    assert len(samples) == 80 and classes[0] == "class_0" and len(classes) == 10
    a = JpegFolder(str(tmp_path), "cifar10", 8, rank=0, world=2, workers=2)
    b = JpegFolder(str(tmp_path), "cifar10", 8, rank=1, world=2, workers=2)
    assert len(a) == len(b) == 5
    assert not set(a._order.tolist()) & set(b._order.tolist())      # disjoint shards
    x, t = a.batch(0)
    assert x.shape == (8, 32, 32, 3) and x.dtype == torch.uint8 and t.dtype == torch.int64
    x2, _ = a.batch(0)
    assert torch.equal(x, x2)                                        # deterministic augmentation
    a.set_epoch(1)
    assert not np.array_equal(a._order, JpegFolder(str(tmp_path), "cifar10", 8, rank=0, world=2)._order)
    v = JpegFolder(str(tmp_path), "cifar10", 4, split="val", shuffle=False)
    assert v.augment is False and len(v) == 5


def test_run_with_real_data_mode(tmp_path, capsys):
    generate_jpeg_tree(str(tmp_path), "mnist", train=64, val=32, workers=1, verbose=False)
    cfg = RunConfig(benchmark="mnist", framework="pytorch", arch="resnet18", batch_size=16, epochs=1,
                    device="cpu", synthetic=False, datadir=str(tmp_path), cuda_graph=False,
                    log_interval=2, eval_steps=2)
    out = run(cfg)
    assert out["throughput"] > 0
    text = capsys.readouterr().out
    assert "train | 1/1 epoch" in text and "valid accuracy" in text


def test_decoded_tensor_transforms_match_pil_path(tmp_path):
    """The nvJPEG path applies the loader's transforms to decoded tensors: same random draws and
    geometry as the PIL path (resampling filters differ slightly, hence a mean-error bound), also
    for images that are NOT already of the training size."""
    from PIL import Image
    rs = np.random.RandomState(0)
    yy, xx = np.mgrid[0:300, 0:400]
    img = np.stack([(xx * 0.5 + yy * 0.2) % 256, (yy * 0.7) % 256, ((xx + yy) * 0.3) % 256], -1).astype(np.uint8)
    img = (img.astype(np.int32) + rs.randint(0, 8, img.shape)).clip(0, 255).astype(np.uint8)
    root = tmp_path / "IMAGENET" / "train" / "class_0"
    root.mkdir(parents=True)
    for i in range(4):
        Image.fromarray(img).save(root / ("%d.png" % i))             # lossless: both paths see the same pixels
    for split, augment in (("train", True), ("train", False)):
        f = JpegFolder(str(tmp_path), "imagenet", 2, split=split, augment=augment, workers=1)
        for seed in range(3):
            ref = f._load_one(str(root / "0.png"), np.random.RandomState(seed))
            d = torch.from_numpy(img).permute(2, 0, 1).contiguous()
            out = f._transform_decoded(d, np.random.RandomState(seed))
            assert tuple(out.shape) == (224, 224, 3) and out.dtype == torch.uint8
            err = (out.float() - torch.from_numpy(ref).float()).abs().mean().item()
            assert err < 3.0, (augment, seed, err)
    generate_jpeg_tree(str(tmp_path / "c"), "cifar10", train=8, val=0, workers=1, verbose=False)
    f = JpegFolder(str(tmp_path / "c"), "cifar10", 2, workers=1)
    path = f.samples[0][0]
    d = torch.from_numpy(np.asarray(Image.open(path).convert("RGB"))).permute(2, 0, 1).contiguous()
    for seed in range(3):
        ref = f._load_one(path, np.random.RandomState(seed))
        assert torch.equal(f._transform_decoded(d, np.random.RandomState(seed)), torch.from_numpy(ref))


def test_native_mnist_and_cifar_files(tmp_path):
    """Real-data mode reads MNIST idx files and CIFAR-10 python batches (what the reference's
    ``datasets.MNIST`` / ``datasets.CIFAR10`` calls read) when they are what DATADIR holds."""
    import gzip
    import pickle
    import struct
    from ddlbench_b200.data import make_source
    from ddlbench_b200.data.imagefolder import ArrayFolder, has_native_format
    rs = np.random.RandomState(0)
    raw = tmp_path / "MNIST" / "raw"
    raw.mkdir(parents=True)
    for stem, n in (("train", 64), ("t10k", 16)):
        x = rs.randint(0, 256, (n, 28, 28), dtype=np.uint8)
        y = (np.arange(n) % 10).astype(np.uint8)
        with gzip.open(raw / (stem + "-images-idx3-ubyte.gz"), "wb") as f:       # gz, as downloaded
            f.write(struct.pack(">IIII", 2051, n, 28, 28) + x.tobytes())
        with open(raw / (stem + "-labels-idx1-ubyte"), "wb") as f:               # and unpacked
            f.write(struct.pack(">II", 2049, n) + y.tobytes())
        if stem == "train":
            mnist_x = x
    cdir = tmp_path / "cifar-10-batches-py"
    cdir.mkdir()
    cx = rs.randint(0, 256, (50, 3072), dtype=np.uint8)
    for i in range(1, 6):
        with open(cdir / ("data_batch_%d" % i), "wb") as f:
            pickle.dump({"data": cx[(i - 1) * 10:i * 10], "labels": list(range(10))}, f)
    with open(cdir / "test_batch", "wb") as f:
        pickle.dump({"data": cx[:10], "labels": list(range(10))}, f)
    assert has_native_format(str(tmp_path), "mnist") and has_native_format(str(tmp_path), "cifar10")
    assert not has_native_format(str(tmp_path), "imagenet")
    m = ArrayFolder(str(tmp_path), "mnist", 8, shuffle=False, workers=1)
    x, t = m.batch(0)
    assert x.shape == (8, 28, 28, 1) and t.tolist() == list(range(8))
    assert np.array_equal(x[..., 0].numpy(), mnist_x[:8])                       # no augmentation for MNIST
    assert len(ArrayFolder(str(tmp_path), "mnist", 8, split="val", workers=1)) == 2
    c = ArrayFolder(str(tmp_path), "cifar10", 10, shuffle=False, augment=False, workers=1)
    x, t = c.batch(0)
    assert x.shape == (10, 32, 32, 3) and np.array_equal(x.numpy(), cx[:10].reshape(10, 3, 32, 32).transpose(0, 2, 3, 1))
    a = ArrayFolder(str(tmp_path), "cifar10", 10, rank=0, world=2, workers=1)
    b = ArrayFolder(str(tmp_path), "cifar10", 10, rank=1, world=2, workers=1)
    assert len(a) == len(b) == 2 and not set(a._order.tolist()) & set(b._order.tolist())
    cfg = RunConfig.from_env(benchmark="cifar10", framework="pytorch", arch="resnet18", synthetic=False,
                             datadir=str(tmp_path), batch_size=10, device="cpu")
    assert isinstance(make_source(cfg, 10, "cpu"), ArrayFolder)
