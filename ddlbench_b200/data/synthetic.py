"""Synthetic datasets of the benchmark's four input shapes.

The reference generates uniform-random JPEG files on disk — ``DATADIR/<SET>/{train,val}/
class_<k>/img_<i>.JPEG`` — with one process per CPU core
(``benchmark/generate_synthetic_data.py:21-107``) and then pays JPEG decode + augmentation in
``DataLoader`` workers inside the timed region.  This module provides:

* ``SyntheticImages``: an in-memory dataset of uint8 HWC images + int64 labels of the named
  shape (device-resident, or pinned host memory for the end-to-end path that includes the
  H2D copy every step);
* ``generate_jpeg_tree``: the on-disk JPEG ImageFolder layout, for parity with the
  reference's ``-s`` real-data code path and for running the unmodified reference.
"""
from __future__ import annotations

import os
import time
from multiprocessing import Pool
from typing import Iterator, Optional, Tuple

import torch

from ..config import DATASET_SIZES
from ..models.zoo import DATASETS

MEAN_STD = {
    "mnist": ((0.1307,), (0.3081,)),
    "cifar10": ((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)),
    "imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "highres": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
}


class SyntheticImages:
    """``num_batches`` distinct random batches, cycled.  Images are uint8 [B, H, W, C]."""

    def __init__(self, dataset: str, batch_size: int, num_batches: int = 4, device="cpu",
                 pinned: bool = False, seed: int = 1, length: Optional[int] = None):
        c, hw, ncls, _ = DATASETS[dataset]
        self.dataset, self.batch_size = dataset, batch_size
        self.shape = (batch_size, hw, hw, c)
        self.num_classes = ncls
        gen = torch.Generator().manual_seed(seed)
        self.images, self.labels = [], []
        for _ in range(num_batches):
            img = torch.randint(0, 256, self.shape, dtype=torch.uint8, generator=gen)
            lab = torch.randint(0, ncls, (batch_size,), dtype=torch.int64, generator=gen)
            if pinned and torch.cuda.is_available():
                img, lab = img.pin_memory(), lab.pin_memory()
            elif torch.device(device).type != "cpu":
                img, lab = img.to(device), lab.to(device)
            self.images.append(img)
            self.labels.append(lab)
        # iterations per epoch if the full synthetic set were used (drop_last)
        self.length = length if length is not None else DATASET_SIZES[dataset][0] // batch_size
        self.mean, self.std = MEAN_STD[dataset]

    def __len__(self) -> int:
        return self.length

    def batch(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        k = i % len(self.images)
        return self.images[k], self.labels[k]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        for i in range(self.length):
            yield self.batch(i)

    def bytes_per_step(self) -> int:
        img, lab = self.batch(0)
        return img.numel() * img.element_size() + lab.numel() * lab.element_size()


# ------------------------------------------------------------------------------- JPEG tree
_SETS = {  # name -> (train, val, size, mode, classes)
    "MNIST": (60000, 10000, 28, "L", 10),
    "CIFAR10": (50000, 10000, 32, "RGB", 10),
    "IMAGENET": (1281167, 50000, 224, "RGB", 1000),
    "HIGHRES": (50000, 10000, 512, "RGB", 1000),
}


def _gen_class(args) -> int:
    root, cls, count, size, mode, seed = args
    from PIL import Image
    g = torch.Generator().manual_seed(seed)
    d = os.path.join(root, "class_%d" % cls)
    os.makedirs(d, exist_ok=True)
    ch = 1 if mode == "L" else 3
    for i in range(count):
        arr = torch.randint(0, 256, (size, size, ch), dtype=torch.uint8, generator=g).numpy()
        img = Image.fromarray(arr[:, :, 0] if ch == 1 else arr, mode=mode)
        img.save(os.path.join(d, "img_%d.JPEG" % i), quality=90)
    return count


def generate_jpeg_tree(datadir: str, name: str, train: Optional[int] = None,
                       val: Optional[int] = None, workers: Optional[int] = None,
                       classes: Optional[int] = None, verbose: bool = True) -> float:
    """Write ``datadir/<NAME>/{train,val}/class_k/img_i.JPEG``; returns images/sec.  ``train``
    / ``val`` override the image counts (the full IMAGENET set is 1.28 M files)."""
    ntr, nva, size, mode, ncls = _SETS[name.upper()]
    ntr = ntr if train is None else train
    nva = nva if val is None else val
    ncls = ncls if classes is None else classes
    t0 = time.time()
    total = 0
    for split, n in (("train", ntr), ("val", nva)):
        root = os.path.join(datadir, name.upper(), split)
        os.makedirs(root, exist_ok=True)
        per = max(1, n // ncls)
        jobs = [(root, c, per, size, mode, 1000 * (split == "val") + c) for c in range(ncls)]
        w = workers or os.cpu_count() or 1
        if w > 1 and len(jobs) > 1:
            with Pool(min(w, len(jobs))) as pool:
                total += sum(pool.map(_gen_class, jobs))
        else:
            total += sum(_gen_class(j) for j in jobs)
    rate = total / max(time.time() - t0, 1e-9)
    if verbose:
        print("Generated %d %s images (%.1f imgs/sec)" % (total, name.upper(), rate))
    return rate
