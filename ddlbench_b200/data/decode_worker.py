"""Image decode + augmentation, importable WITHOUT torch (numpy + PIL only): the body of
``JpegFolder._load_one`` lives here so that it can also run in plain worker processes.

Why processes: PIL releases the GIL inside libjpeg / resize, but the Python glue around every image
does not, and a thread pool tops out near 900 images/s regardless of the core count (measured: 4,
8 and 16 threads on an 8-core host: 885 / 778 / 879 images/s); worker processes scale with the
cores (1 / 2 / 4 / 8 processes: 614 / 1 086 / 2 407 / 4 766 images/s on the same host).  The
reference gets its parallelism the same way (``DataLoader(num_workers=CORES_GPU)``,
``benchmark/imagenet/imagenet_pytorch.py:95-110``).

Workers write the decoded uint8 pixels straight into a ring of batch slots — a file in
``/dev/shm`` mapped by both sides (``np.memmap``), so no pixel data goes through a pipe and there is
no ``multiprocessing.shared_memory`` bookkeeping to leak.

The workers are plain ``python decode_worker.py`` children driven over stdin / stdout pickle
frames (``DecodePool`` below), NOT ``multiprocessing`` processes: nothing is forked from a process
that holds a CUDA context and NCCL threads, and — unlike spawn / forkserver — the children never
re-import the parent's ``__main__`` module (an unguarded user script would otherwise re-run
itself in every worker)."""
from __future__ import annotations

import numpy as np

_RINGS = {}


def sample_rng(seed: int, epoch: int, index: int) -> np.random.RandomState:
    """The per-sample generator (same in every decode mode: threads, processes, nvJPEG)."""
    return np.random.RandomState((seed * 1000003 + epoch * 7919 + int(index)) & 0x7FFFFFFF)


def augment_array(a: np.ndarray, dataset: str, size: int, augment: bool, rng) -> np.ndarray:
    """[H, W(, C)] uint8 of the training size -> augmented [s, s, C] (CIFAR pad-crop, flip)."""
    s = size
    if a.ndim == 2:
        a = a[:, :, None]
    if augment:
        if dataset == "cifar10":                      # RandomCrop(32, padding=4)
            p = np.zeros((s + 8, s + 8, a.shape[2]), np.uint8)
            p[4:4 + s, 4:4 + s] = a
            y0, x0 = rng.randint(0, 9), rng.randint(0, 9)
            a = p[y0:y0 + s, x0:x0 + s]
        if dataset != "mnist" and rng.rand() < 0.5:
            a = a[:, ::-1]
    return np.ascontiguousarray(a)


def load_image(path: str, dataset: str, size: int, channels: int, augment: bool, rng) -> np.ndarray:
    """Decode one file and apply the reference's transforms (``imagenet_pytorch.py:68-92``:
    RandomResizedCrop(224) + flip for training, Resize(256) + CenterCrop(224) for validation)."""
    from PIL import Image
    with Image.open(path) as im:
        im = im.convert("L" if channels == 1 else "RGB")
        s = size
        if dataset in ("imagenet", "highres"):
            if augment and dataset == "imagenet":
                w, h = im.size
                area = w * h * rng.uniform(0.08, 1.0)
                ar = np.exp(rng.uniform(np.log(3 / 4), np.log(4 / 3)))
                cw = int(min(w, max(1, round(np.sqrt(area * ar)))))
                ch = int(min(h, max(1, round(np.sqrt(area / ar)))))
                x0, y0 = rng.randint(0, w - cw + 1), rng.randint(0, h - ch + 1)
                im = im.resize((s, s), Image.BILINEAR, box=(x0, y0, x0 + cw, y0 + ch))
            elif im.size != (s, s):
                w, h = im.size
                r = (s * 256 // 224) / min(w, h)
                im = im.resize((max(s, round(w * r)), max(s, round(h * r))), Image.BILINEAR)
                w, h = im.size
                im = im.crop(((w - s) // 2, (h - s) // 2, (w - s) // 2 + s, (h - s) // 2 + s))
        elif im.size != (s, s):
            im = im.resize((s, s), Image.BILINEAR)
        a = np.asarray(im, dtype=np.uint8)
    return augment_array(a, dataset, size, augment, rng)


def _ring(path: str, shape) -> np.ndarray:
    key = (path, tuple(shape))
    arr = _RINGS.get(key)
    if arr is None:
        _RINGS.clear()                                # one ring per loader; old maps are dropped
        arr = np.memmap(path, dtype=np.uint8, mode="r+", shape=tuple(shape))
        _RINGS[key] = arr
    return arr


def decode_into(ring_path: str, ring_shape, slot: int, items, dataset: str, size: int, channels: int,
                augment: bool, seed: int, epoch: int) -> int:
    """Worker entry: ``items`` = [(position in the batch, sample index, path), ...]; pixels go to
    ``ring[slot, position]``."""
    ring = _ring(ring_path, ring_shape)
    for pos, index, path in items:
        ring[slot, pos] = load_image(path, dataset, size, channels, augment, sample_rng(seed, epoch, index))
    return len(items)


# ------------------------------------------------------------------------------ worker process
def _serve() -> None:
    """``python decode_worker.py``: read pickled requests from stdin, answer on stdout.
    Request: (task id, "decode", kwargs of decode_into) | (id, "ping", None) | None (exit)."""
    import pickle
    import sys
    rd, wr = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr                       # stray prints must not corrupt the protocol
    while True:
        try:
            req = pickle.load(rd)
        except EOFError:
            return
        if req is None:
            return
        tid, what, kw = req
        try:
            out = decode_into(**kw) if what == "decode" else 0
            ans = (tid, True, out)
        except BaseException as e:  # noqa: BLE001 - reported to the parent, the worker lives on
            ans = (tid, False, "%s: %s" % (type(e).__name__, e))
        pickle.dump(ans, wr, protocol=pickle.HIGHEST_PROTOCOL)
        wr.flush()


class DecodePool:
    """N decode worker processes; ``submit(kwargs)`` returns a ``concurrent.futures.Future``."""

    def __init__(self, n: int, start_timeout: float = 60.0):
        import os
        import subprocess
        import sys
        import threading
        self._threading = threading
        self.n = n
        self._procs, self._locks, self._pending, self._outstanding = [], [], [], []
        self._tid = 0
        self._tid_lock = threading.Lock()
        self.broken: "str | None" = None
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        script = os.path.abspath(__file__)
        for _ in range(n):
            p = subprocess.Popen([sys.executable, "-u", script], stdin=subprocess.PIPE,
                                 stdout=subprocess.PIPE, env=env, close_fds=True)
            self._procs.append(p)
            self._locks.append(threading.Lock())
            self._pending.append({})
            self._outstanding.append(0)
        self._readers = [threading.Thread(target=self._reader, args=(w,), daemon=True) for w in range(n)]
        for t in self._readers:
            t.start()
        import time
        deadline = time.time() + start_timeout                 # ONE budget for all workers together
        try:
            for f in [self._send(w, "ping", None) for w in range(n)]:
                f.result(timeout=max(0.1, deadline - time.time()))
        except BaseException:
            self.close()
            raise

    def _reader(self, w: int) -> None:
        import pickle
        rd = self._procs[w].stdout
        try:
            while True:
                tid, ok, val = pickle.load(rd)
                with self._locks[w]:
                    fut = self._pending[w].pop(tid, None)
                    self._outstanding[w] -= 1
                if fut is not None:
                    if ok:
                        fut.set_result(val)
                    else:
                        fut.set_exception(RuntimeError(val))
        except Exception as e:  # noqa: BLE001 - EOF / broken pipe: the worker is gone
            self.broken = self.broken or ("worker %d: %s" % (w, type(e).__name__))
            with self._locks[w]:
                futs, self._pending[w] = list(self._pending[w].values()), {}
            for fut in futs:
                if not fut.done():
                    fut.set_exception(RuntimeError("decode worker exited"))

    def _send(self, w: int, what: str, kw):
        import pickle
        from concurrent.futures import Future
        fut = Future()
        with self._tid_lock:
            self._tid += 1
            tid = self._tid
        with self._locks[w]:
            self._pending[w][tid] = fut
            self._outstanding[w] += 1
            try:
                pickle.dump((tid, what, kw), self._procs[w].stdin, protocol=pickle.HIGHEST_PROTOCOL)
                self._procs[w].stdin.flush()
            except Exception as e:  # noqa: BLE001
                self._pending[w].pop(tid, None)
                self._outstanding[w] -= 1
                self.broken = self.broken or ("worker %d: %s" % (w, type(e).__name__))
                fut.set_exception(RuntimeError("decode worker pipe closed"))
        return fut

    def submit(self, kw):
        if self.broken:
            raise RuntimeError(self.broken)
        w = min(range(self.n), key=lambda i: self._outstanding[i])      # least loaded
        return self._send(w, "decode", kw)

    def close(self) -> None:
        import pickle
        for w, p in enumerate(self._procs):
            got = self._locks[w].acquire(timeout=1.0)          # never block on a stuck writer
            try:
                if got:
                    pickle.dump(None, p.stdin)
                    p.stdin.flush()
                    p.stdin.close()
                else:
                    p.kill()
            except Exception:  # noqa: BLE001
                pass
            finally:
                if got:
                    self._locks[w].release()
        for p in self._procs:
            try:
                p.wait(timeout=2)
            except Exception:  # noqa: BLE001
                p.kill()
        self._procs = []


if __name__ == "__main__":
    _serve()
