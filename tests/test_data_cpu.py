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


def test_decode_worker_processes_match_the_thread_pool_bit_for_bit(tmp_path):
    """``decode_procs=N``: plain ``python decode_worker.py`` children write into the shared ring; the
    batches must be IDENTICAL to the thread pool's (same per-sample generator, same transforms), for
    the augmented training split, the validation split, across epochs and for non-square inputs."""
    from PIL import Image
    generate_jpeg_tree(str(tmp_path), "imagenet", train=48, val=16, workers=1, verbose=False)
    rng = np.random.RandomState(0)                    # a few files that are NOT of the training size
    d = os.path.join(tmp_path, "IMAGENET", "train", "class_0")
    for k, (w, h) in enumerate(((300, 260), (500, 375), (231, 480))):
        Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8), "RGB").save(os.path.join(d, "odd_%d.JPEG" % k))
    for split in ("train", "val"):
        thr = JpegFolder(str(tmp_path), "imagenet", 8, split=split, seed=5, workers=2)
        prc = JpegFolder(str(tmp_path), "imagenet", 8, split=split, seed=5, workers=2, decode_procs=3)
        try:
            for epoch in range(2):
                thr.set_epoch(epoch)
                prc.set_epoch(epoch)
                for i in range(len(thr)):
                    (xa, ta), (xb, tb) = thr.batch(i), prc.batch(i)
                    assert torch.equal(ta, tb) and torch.equal(xa, xb), (split, epoch, i)
            assert prc._procs_error is None and prc._ppool is not None and prc._ppool.n == 3
            ring = prc._ring_path
            assert os.path.exists(ring)
        finally:
            thr.close()
            prc.close()
        assert not os.path.exists(ring)                # the ring file is removed with the loader


def test_decode_worker_pool_failure_falls_back_to_threads(tmp_path, capfd):
    """Workers that die (or never start) must not take the loader down: the remaining batches come
    from the thread pool, and the reason is kept for the bench line (``decode_procs_error``)."""
    generate_jpeg_tree(str(tmp_path), "cifar10", train=64, val=8, workers=1, verbose=False)
    ref = JpegFolder(str(tmp_path), "cifar10", 8, seed=2, workers=2)
    src = JpegFolder(str(tmp_path), "cifar10", 8, seed=2, workers=2, decode_procs=2)
    try:
        x0, _ = src.batch(0)
        assert src._procs_error is None and torch.equal(x0, ref.batch(0)[0])
        with src._lock:                               # drop prefetched work, then kill the workers
            for f in src._pending.values():
                f.cancel()
            src._pending.clear()
        for p in src._ppool._procs:
            p.kill()
        for i in range(1, len(src)):
            assert torch.equal(src.batch(i)[0], ref.batch(i)[0])
        assert src._procs_error is not None and src._ppool is None
        assert "decode processes disabled" in capfd.readouterr().err
    finally:
        src.close()
        ref.close()
    # a pool that cannot start: same fallback
    import ddlbench_b200.data.decode_worker as dw
    real = dw.DecodePool

    class Boom:
        def __init__(self, n):
            raise OSError("no processes today")
    dw.DecodePool = Boom
    try:
        bad = JpegFolder(str(tmp_path), "cifar10", 8, seed=2, workers=2, decode_procs=4)
        assert torch.equal(bad.batch(0)[0], ref.batch(0)[0] if False else JpegFolder(str(tmp_path), "cifar10", 8, seed=2, workers=2).batch(0)[0])
        assert "no processes today" in bad._procs_error
        bad.close()
    finally:
        dw.DecodePool = real


def test_decode_worker_module_is_importable_without_torch():
    import subprocess
    import sys
    code = ("import sys, importlib.util; spec = importlib.util.spec_from_file_location('dw', sys.argv[1]); "
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); "
            "assert 'torch' not in sys.modules; print('ok')")
    import ddlbench_b200.data.decode_worker as dw
    r = subprocess.run([sys.executable, "-c", code, dw.__file__], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
