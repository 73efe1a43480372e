"""Copy an N-class subset of an ImageNet-layout tree (``<root>/{train,val}/<wnid>/*.JPEG``) to a
scratch directory — the utility of ``benchmark/imagenet/sampler.py:10-110`` (a LISA-cluster
helper there; generic here).  Class choice is seeded; images per class can be capped.

    python tools/imagenet_subset.py /data/imagenet /scratch/imagenet100 --classes 100 --per-class 300
"""
import argparse
import os
import random
import shutil
from concurrent.futures import ThreadPoolExecutor


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--per-class", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--splits", nargs="+", default=["train", "val"])
    ap.add_argument("--symlink", action="store_true", help="link instead of copying")
    ap.add_argument("--workers", type=int, default=16)
    a = ap.parse_args(argv)
    classes = sorted(d for d in os.listdir(os.path.join(a.src, a.splits[0]))
                     if os.path.isdir(os.path.join(a.src, a.splits[0], d)))
    random.Random(a.seed).shuffle(classes)
    chosen = sorted(classes[:a.classes])
    jobs = []
    for split in a.splits:
        for c in chosen:
            sdir, ddir = os.path.join(a.src, split, c), os.path.join(a.dst, split, c)
            if not os.path.isdir(sdir):
                continue
            os.makedirs(ddir, exist_ok=True)
            files = sorted(os.listdir(sdir))
            if a.per_class:
                files = files[:a.per_class]
            jobs += [(os.path.join(sdir, f), os.path.join(ddir, f)) for f in files]

    def put(job):
        s, d = job
        if os.path.lexists(d):
            return
        if a.symlink:
            os.symlink(os.path.abspath(s), d)
        else:
            shutil.copyfile(s, d)
    with ThreadPoolExecutor(a.workers) as ex:
        list(ex.map(put, jobs))
    print("copied %d files of %d classes to %s" % (len(jobs), len(chosen), a.dst))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
