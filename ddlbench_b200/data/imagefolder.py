"""Real-data input mode (``run.sh -s``): the ``DATADIR/<SET>/{train,val}/class_k/*.JPEG`` tree.

The reference reads this tree with torchvision ``ImageFolder`` + ``DataLoader(num_workers=
CORES_GPU, pin_memory=True, shuffle=True)`` and per-sample CPU transforms producing fp32 CHW
tensors (``benchmark/imagenet/imagenet_pytorch.py:68-110``, ``cifar10_pytorch.py:164-186``,
``mnist_pytorch.py:136-150`` incl. the grayscale loader).  Here the loader's product is the
same thing the synthetic source yields — a PINNED uint8 [B, H, W, C] batch + int64 labels — so
everything downstream (double-buffered H2D copy stream, fused u8->bf16 normalise kernel,
captured step) is shared.  Decode runs in a thread pool (PIL releases the GIL inside libjpeg)
with ``prefetch`` batches in flight, or — ``decode_procs=N`` / ``DDLB_DECODE_PROCS=N`` — in N worker
PROCESSES that write into a shared ring of batch slots (``decode_worker.py``: the thread pool is
bound by the GIL at ~900 images/s, processes scale with the cores; any failure of the process pool
falls back to the threads); ``gpu_decode=True`` ships the compressed bytes instead and decodes with
nvJPEG on the device (``torchvision.io.decode_jpeg(device=cuda)``).

Sharding for data parallelism follows ``DistributedSampler(num_replicas, rank)`` semantics
(``imagenet_horovod.py:113-123``): a per-epoch seeded permutation, padded to a multiple of the
world size, strided by rank.
"""
from __future__ import annotations

import os
import queue
import sys
import tempfile
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..models import DATASETS
from . import decode_worker
from .synthetic import MEAN_STD

_EXT = (".jpeg", ".jpg", ".png", ".ppm", ".bmp")


def scan_tree(root: str) -> Tuple[List[Tuple[str, int]], List[str]]:
    """(path, class index) samples + sorted class names — ImageFolder's contract."""
    classes = sorted(d.name for d in os.scandir(root) if d.is_dir())
    samples = []
    for ci, c in enumerate(classes):
        for dirpath, _, files in sorted(os.walk(os.path.join(root, c))):
            for f in sorted(files):
                if f.lower().endswith(_EXT):
                    samples.append((os.path.join(dirpath, f), ci))
    if not samples:
        raise FileNotFoundError("no images under %s" % root)
    return samples, classes


def _remove_stale_rings(base: str) -> None:
    """Ring files of loaders whose process died without ``close()`` (name = ddlb_ring_<pid>_*)."""
    try:
        for name in os.listdir(base):
            if not name.startswith("ddlb_ring_"):
                continue
            try:
                pid = int(name.split("_")[2])
                os.kill(pid, 0)                       # raises if the owner is gone
            except (ValueError, IndexError, PermissionError):
                continue
            except ProcessLookupError:
                try:
                    os.unlink(os.path.join(base, name))
                except OSError:
                    pass
    except OSError:
        pass


class JpegFolder:
    """Drop-in for ``SyntheticImages`` backed by an image tree."""

    def __init__(self, datadir: str, dataset: str, batch_size: int, split: str = "train",
                 shuffle: bool = True, seed: int = 1, rank: int = 0, world: int = 1,
                 workers: Optional[int] = None, pinned: bool = False, augment: Optional[bool] = None,
                 prefetch: int = 3, gpu_decode: bool = False, device="cpu",
                 decode_procs: Optional[int] = None):
        c, hw, ncls, _ = DATASETS[dataset]
        name = {"highres": "HIGHRES", "imagenet": "IMAGENET", "cifar10": "CIFAR10",
                "mnist": "MNIST"}[dataset]
        self.root = os.path.join(datadir, name, split)
        if getattr(self, "_scan_override", None) is not None:
            self.samples, self.classes = self._scan_override, [str(k) for k in range(ncls)]
        else:
            self.samples, self.classes = scan_tree(self.root)
        self.dataset, self.batch_size, self.split = dataset, batch_size, split
        self.channels, self.size, self.num_classes = c, hw, ncls
        self.shape = (batch_size, hw, hw, c)
        self.mean, self.std = MEAN_STD[dataset]
        self.shuffle, self.seed, self.rank, self.world = shuffle, seed, rank, world
        self.augment = (split == "train") if augment is None else augment
        self.pinned = pinned and torch.cuda.is_available()
        self.gpu_decode = gpu_decode and torch.cuda.is_available()
        self.device = torch.device(device)
        nthreads = workers or min(32, os.cpu_count() or 4)
        self.prefetch = max(1, prefetch)
        # two levels: `pool` assembles whole batches (`prefetch` of them in flight), `img_pool`
        # decodes the images of a batch in parallel (PIL drops the GIL inside libjpeg / resize)
        self.pool = ThreadPoolExecutor(max_workers=self.prefetch + 1)
        self.img_pool = ThreadPoolExecutor(max_workers=nthreads)
        self.epoch = 0
        self._lock = threading.Lock()
        self._pending: Dict[int, object] = {}
        # optional process pool for the decode (started lazily on the first batch)
        if decode_procs is None:
            decode_procs = int(os.environ.get("DDLB_DECODE_PROCS", "0") or 0)
        self.decode_procs = 0 if (self.gpu_decode or getattr(self, "_scan_override", None) is not None) \
            else max(0, int(decode_procs))
        self._ppool: Optional[decode_worker.DecodePool] = None
        self._ring = None
        self._ring_path = ""
        self._free_slots: "queue.Queue[int]" = queue.Queue()
        self._procs_error: Optional[str] = None
        self._procs_lock = threading.Lock()
        self._closed = False
        self._order = self._make_order()
        self.length = len(self._order) // batch_size           # drop_last
        if self.length == 0:
            raise ValueError("batch size %d exceeds the %d samples of this shard"
                             % (batch_size, len(self._order)))

    # ---------------------------------------------------------------- sampling
    def _make_order(self) -> np.ndarray:
        n = len(self.samples)
        idx = (np.random.RandomState(self.seed + self.epoch).permutation(n) if self.shuffle
               else np.arange(n))
        if self.world > 1:
            pad = (-n) % self.world
            idx = np.concatenate([idx, idx[:pad]])[self.rank::self.world]
        return idx

    def set_epoch(self, epoch: int) -> None:
        with self._lock:
            self.epoch = epoch
            self._order = self._make_order()
            for f in self._pending.values():
                f.cancel()
            self._pending.clear()

    def __len__(self) -> int:
        return self.length

    # ---------------------------------------------------------------- decode
    def _load_one(self, path: str, rng: np.random.RandomState) -> np.ndarray:
        return decode_worker.load_image(path, self.dataset, self.size, self.channels, self.augment, rng)

    def _augment(self, a: np.ndarray, rng: np.random.RandomState) -> np.ndarray:
        return decode_worker.augment_array(a, self.dataset, self.size, self.augment, rng)

    # ---------------------------------------------------------------- process pool
    def _start_procs(self) -> bool:
        """Ring file + worker processes.  False (and the thread pool stays in charge) on any failure."""
        with self._procs_lock:             # several batch threads arrive here at once
            return self._start_procs_locked()

    def _start_procs_locked(self) -> bool:
        if self._ppool is not None:
            return True
        if self._procs_error is not None or self.decode_procs <= 0:
            return False
        try:
            slots = self.prefetch + 2
            shape = (slots,) + tuple(self.shape)
            need = int(np.prod(shape))
            base = tempfile.gettempdir()
            try:        # tmpfs when it has room (a full tmpfs would SIGBUS the writers), else a disk-backed map
                if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK):
                    st = os.statvfs("/dev/shm")
                    if st.f_bavail * st.f_frsize > 4 * need + (64 << 20):
                        base = "/dev/shm"
            except OSError:
                pass
            _remove_stale_rings(base)
            fd, path = tempfile.mkstemp(prefix="ddlb_ring_%d_" % os.getpid(), dir=base)
            try:
                os.ftruncate(fd, need)
            finally:
                os.close(fd)
            self._ring_path, self._ring_shape = path, shape
            self._ring = np.memmap(path, dtype=np.uint8, mode="r+", shape=shape)
            # plain `python decode_worker.py` children (no CUDA context, no inherited threads, no
            # re-import of the parent's __main__): see decode_worker.DecodePool
            pool = decode_worker.DecodePool(self.decode_procs)
            for k in range(slots):
                self._free_slots.put(k)
            self._ppool = pool
            return True
        except Exception as e:  # noqa: BLE001 - the loader must keep working on threads
            self._fail_procs("start: %s: %s" % (type(e).__name__, e))
            return False

    def _fail_procs(self, why: str) -> None:
        if self._closed:                   # batches still in flight while the loader shuts down
            return
        if self._procs_error is None:
            self._procs_error = why[:300]
            print("JpegFolder: decode processes disabled (%s); using the thread pool" % self._procs_error,
                  file=sys.stderr, flush=True)
        pool, self._ppool = self._ppool, None
        if pool is not None:
            try:
                pool.close()
            except Exception:  # noqa: BLE001
                pass

    def _decode_batch_procs(self, sel, epoch: int, out: np.ndarray) -> bool:
        """Decode the batch ``sel`` into ``out`` with the worker processes; False -> caller falls back."""
        pool = self._ppool
        if pool is None or self._closed:
            return False
        try:
            slot = self._free_slots.get(timeout=60)
        except queue.Empty:
            self._fail_procs("no free ring slot for 60 s")
            return False
        try:
            items = [(k, int(j), self.samples[j][0]) for k, j in enumerate(sel)]
            # 2-4 images per request: `prefetch` batches are in flight, so even 16 requests per batch
            # keep 64 workers busy, and the parent's per-request cost (pickle + pipe + future) halves
            chunk = max(2 if len(items) >= 16 else 1, min(4, len(items) // max(self.decode_procs, 1)))
            futs = [pool.submit(dict(ring_path=self._ring_path, ring_shape=self._ring_shape, slot=slot,
                                     items=items[c:c + chunk], dataset=self.dataset, size=self.size,
                                     channels=self.channels, augment=self.augment, seed=self.seed,
                                     epoch=epoch)) for c in range(0, len(items), chunk)]
            import time
            deadline = time.time() + 60.0                       # one budget for the whole batch
            for f in futs:
                f.result(timeout=max(0.1, deadline - time.time()))
            np.copyto(out, self._ring[slot])
            return True
        except Exception as e:  # noqa: BLE001 - broken pool, worker crash, timeout
            self._fail_procs("%s: %s" % (type(e).__name__, e))
            return False
        finally:
            self._free_slots.put(slot)

    def _transform_decoded(self, d: torch.Tensor, rng: np.random.RandomState) -> torch.Tensor:
        """The transforms of ``_load_one`` on an already decoded [C, H, W] uint8 tensor (device
        side of the nvJPEG path; the same random draws, bilinear resampling with antialiasing).
        Returns [s, s, C] uint8."""
        import torch.nn.functional as F
        s = self.size
        _, h, w = d.shape

        def resize(t, oh, ow):
            return F.interpolate(t[None].float(), size=(oh, ow), mode="bilinear", align_corners=False,
                                 antialias=True)[0]

        if self.dataset in ("imagenet", "highres"):
            if self.augment and self.dataset == "imagenet":
                area = w * h * rng.uniform(0.08, 1.0)
                ar = np.exp(rng.uniform(np.log(3 / 4), np.log(4 / 3)))
                cw = int(min(w, max(1, round(np.sqrt(area * ar)))))
                ch = int(min(h, max(1, round(np.sqrt(area / ar)))))
                x0, y0 = rng.randint(0, w - cw + 1), rng.randint(0, h - ch + 1)
                d = resize(d[:, y0:y0 + ch, x0:x0 + cw], s, s)
            elif (h, w) != (s, s):
                r = (s * 256 // 224) / min(w, h)
                nh, nw = max(s, round(h * r)), max(s, round(w * r))
                d = resize(d, nh, nw)
                d = d[:, (nh - s) // 2:(nh - s) // 2 + s, (nw - s) // 2:(nw - s) // 2 + s]
        elif (h, w) != (s, s):
            d = resize(d, s, s)
        if d.is_floating_point():
            d = d.round().clamp_(0, 255).to(torch.uint8)
        if self.augment:
            if self.dataset == "cifar10":
                pad = torch.zeros((d.shape[0], s + 8, s + 8), dtype=torch.uint8, device=d.device)
                pad[:, 4:4 + s, 4:4 + s] = d
                y0, x0 = rng.randint(0, 9), rng.randint(0, 9)
                d = pad[:, y0:y0 + s, x0:x0 + s]
            if self.dataset != "mnist" and rng.rand() < 0.5:
                d = d.flip(2)
        return d.permute(1, 2, 0)

    def _make_batch(self, i: int, order: np.ndarray, epoch: int):
        sel = order[(i % self.length) * self.batch_size:(i % self.length + 1) * self.batch_size]
        labels = torch.tensor([self.samples[j][1] for j in sel], dtype=torch.int64)
        if self.gpu_decode:
            from torchvision.io import ImageReadMode, decode_jpeg, read_file
            raw = [read_file(self.samples[j][0]) for j in sel]
            mode = ImageReadMode.GRAY if self.channels == 1 else ImageReadMode.RGB
            rngs = [decode_worker.sample_rng(self.seed, epoch, j) for j in sel]
            try:
                dec = decode_jpeg(raw, device=self.device, mode=mode)
                img = torch.stack([self._transform_decoded(d, r) for d, r in zip(dec, rngs)]).contiguous()
            except RuntimeError:
                # a file nvJPEG does not take (CMYK, progressive corner cases, not a JPEG): decode
                # this batch on the host, as the default path does
                img = torch.stack([torch.from_numpy(self._load_one(self.samples[j][0], r))
                                   for j, r in zip(sel, rngs)]).to(self.device, non_blocking=True)
            return img, labels.to(self.device, non_blocking=True)
        img = torch.empty(self.shape, dtype=torch.uint8)
        if self.pinned:
            img, labels = img.pin_memory(), labels.pin_memory()
        out = img.numpy()

        if self.decode_procs > 0 and self._start_procs() and self._decode_batch_procs(sel, epoch, out):
            return img, labels

        def one(kj):
            k, j = kj
            out[k] = self._load_one(self.samples[j][0], decode_worker.sample_rng(self.seed, epoch, j))

        list(self.img_pool.map(one, enumerate(sel)))
        return img, labels

    def _submit(self, i: int) -> None:
        if i not in self._pending:
            self._pending[i] = self.pool.submit(self._make_batch, i, self._order, self.epoch)

    def batch(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        with self._lock:
            for k in range(i, i + self.prefetch):
                self._submit(k)
            fut = self._pending.pop(i)
            for k in [k for k in self._pending if k < i]:
                self._pending.pop(k).cancel()
        return fut.result()

    def __iter__(self):
        for i in range(self.length):
            yield self.batch(i)

    def bytes_per_step(self) -> int:
        return int(np.prod(self.shape)) + self.batch_size * 8

    def close(self) -> None:
        self._closed = True
        self.pool.shutdown(wait=False, cancel_futures=True)
        self.img_pool.shutdown(wait=False, cancel_futures=True)
        pool, self._ppool = self._ppool, None
        if pool is not None:
            pool.close()
        self._ring = None
        if self._ring_path:
            try:
                os.unlink(self._ring_path)
            except OSError:
                pass
            self._ring_path = ""

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ------------------------------------------------------------------------------------------
# MNIST / CIFAR-10 in the formats torchvision's dataset classes read (what the reference's
# scripts use in real-data mode: ``datasets.MNIST(datadir)``, ``datasets.CIFAR10(datadir)``,
# ``cifar10_pytorch.py:163-181``, ``mnist_pytorch.py:175-190``)
# ------------------------------------------------------------------------------------------
def _open_maybe_gz(path: str):
    import gzip
    if os.path.exists(path):
        return open(path, "rb")
    if os.path.exists(path + ".gz"):
        return gzip.open(path + ".gz", "rb")
    raise FileNotFoundError(path)


def _mnist_bases(root: str):
    """Where the idx files may sit below DATADIR: ``datasets.MNIST(root)`` puts them in
    ``root/MNIST/raw``; the PipeDream runtime passes ``data_dir + '/MNIST'`` as that root
    (``main_with_runtime.py:275``), the benchmark scripts pass DATADIR itself (``mnist_pytorch.py:175``)."""
    return (os.path.join(root, "MNIST", "raw"), os.path.join(root, "MNIST", "MNIST", "raw"),
            os.path.join(root, "raw"), root)


def load_mnist_idx(root: str, train: bool) -> Tuple[np.ndarray, np.ndarray]:
    """``<root>/MNIST/raw/{train,t10k}-{images-idx3,labels-idx1}-ubyte[.gz]`` -> ([N, 28, 28] uint8, [N])."""
    import struct
    stem = "train" if train else "t10k"
    for base in _mnist_bases(root):
        if os.path.exists(os.path.join(base, stem + "-images-idx3-ubyte")) or \
                os.path.exists(os.path.join(base, stem + "-images-idx3-ubyte.gz")):
            break
    else:
        raise FileNotFoundError("no MNIST idx files under %s" % root)
    with _open_maybe_gz(os.path.join(base, stem + "-images-idx3-ubyte")) as f:
        magic, n, h, w = struct.unpack(">IIII", f.read(16))
        assert magic == 2051, "not an idx3 image file"
        x = np.frombuffer(f.read(n * h * w), dtype=np.uint8).reshape(n, h, w)
    with _open_maybe_gz(os.path.join(base, stem + "-labels-idx1-ubyte")) as f:
        magic, n2 = struct.unpack(">II", f.read(8))
        assert magic == 2049 and n2 == n, "label file does not match"
        y = np.frombuffer(f.read(n), dtype=np.uint8).astype(np.int64)
    return x, y


def load_cifar10_batches(root: str, train: bool) -> Tuple[np.ndarray, np.ndarray]:
    """``<root>/cifar-10-batches-py/{data_batch_1..5,test_batch}`` -> ([N, 32, 32, 3] uint8, [N])."""
    import pickle
    d = os.path.join(root, "cifar-10-batches-py")
    xs, ys = [], []
    for name in (["data_batch_%d" % i for i in range(1, 6)] if train else ["test_batch"]):
        with open(os.path.join(d, name), "rb") as f:
            e = pickle.load(f, encoding="latin1")
        xs.append(np.asarray(e["data"], dtype=np.uint8))
        ys += list(e["labels"] if "labels" in e else e["fine_labels"])
    x = np.vstack(xs).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)
    return np.ascontiguousarray(x), np.asarray(ys, dtype=np.int64)


def has_native_format(datadir: str, dataset: str) -> bool:
    if dataset == "cifar10":
        return os.path.isdir(os.path.join(datadir, "cifar-10-batches-py"))
    if dataset == "mnist":
        return any(os.path.exists(os.path.join(b, "train-images-idx3-ubyte" + ext))
                   for b in _mnist_bases(datadir) for ext in ("", ".gz"))
    return False


class ArrayFolder(JpegFolder):
    """The real MNIST / CIFAR-10 files held in memory; same sharding, augmentation, prefetch and
    batch interface as ``JpegFolder`` (a "sample" is an index into the array)."""

    def __init__(self, datadir: str, dataset: str, batch_size: int, split: str = "train", **kw):
        self._arrays = (load_cifar10_batches if dataset == "cifar10" else load_mnist_idx)(datadir, split == "train")
        kw.pop("gpu_decode", None)
        self._scan_override = [(i, int(l)) for i, l in enumerate(self._arrays[1])]
        super().__init__(datadir, dataset, batch_size, split=split, gpu_decode=False, **kw)

    def _load_one(self, index, rng: np.random.RandomState) -> np.ndarray:
        return self._augment(self._arrays[0][index], rng)
