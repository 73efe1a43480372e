"""Synthetic JPEG ImageFolder tree shared by BOTH arms of bench.py (stand-alone: imports nothing
from ddlbench_b200 or the reference).  Layout = what the reference's
``benchmark/generate_synthetic_data.py`` writes: ``<root>/IMAGENET/{train,val}/class_k/img_i.JPEG``,
224x224 RGB noise at quality 90.  ``unique`` distinct files are encoded (in parallel); the rest of
``n_train`` are hard links to them, so a long epoch costs no extra encode time or disk while the
readers still open / decode / augment every sample."""
import os
from concurrent.futures import ProcessPoolExecutor


def _encode(args):
    path, seed = args
    import numpy as np
    from PIL import Image
    rng = np.random.RandomState(seed)
    arr = rng.randint(0, 256, (224, 224, 3), dtype=np.uint8)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    Image.fromarray(arr, "RGB").save(path, quality=90)
    return path


def make_tree(root, n_train, n_val, classes=8, unique=2048, workers=None):
    jobs, links = [], []
    for split, n in (("train", n_train), ("val", n_val)):
        per = max(1, (n + classes - 1) // classes)
        uq = max(1, min(per, unique // classes))
        for c in range(classes):
            d = os.path.join(root, "IMAGENET", split, "class_%d" % c)
            os.makedirs(d, exist_ok=True)
            for i in range(per):
                path = os.path.join(d, "img_%d.JPEG" % i)
                if i < uq:
                    jobs.append((path, (1000 * (split == "val") + c) * 100003 + i))
                else:
                    links.append((os.path.join(d, "img_%d.JPEG" % (i % uq)), path))
    w = workers or min(32, os.cpu_count() or 4)
    if w > 1 and len(jobs) > 8:
        with ProcessPoolExecutor(max_workers=w) as ex:
            list(ex.map(_encode, jobs, chunksize=16))
    else:
        for j in jobs:
            _encode(j)
    for src, dst in links:
        try:
            os.link(src, dst)
        except OSError:
            import shutil
            shutil.copyfile(src, dst)
    return len(jobs), len(links)
