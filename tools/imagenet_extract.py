"""Unpack the ILSVRC2012 archives into the ``train/<wnid>/`` + ``val/<wnid>/`` ImageFolder layout
with md5 verification — the job of ``pipedream-fork/scripts/imagenet.py`` (there is no network in
a sandbox, so ``download_imagenet.py``'s fetch step has no counterpart; point this at archives
you already have).

    python tools/imagenet_extract.py --train ILSVRC2012_img_train.tar --val ILSVRC2012_img_val.tar \
        --devkit ILSVRC2012_devkit_t12.tar.gz --out /data/imagenet [--no-verify]
"""
import argparse
import hashlib
import io
import os
import tarfile

MD5 = {"ILSVRC2012_img_train.tar": "1d675b47d978889d74fa0da5fadfb00e",
       "ILSVRC2012_img_val.tar": "29b22e2961454d5413ddabcf34fc5622",
       "ILSVRC2012_devkit_t12.tar.gz": "fa75699e90414af021442c21a62c3abf"}


def md5sum(path: str, chunk: int = 1 << 22) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(chunk), b""):
            h.update(block)
    return h.hexdigest()


def verify(path: str) -> None:
    want = MD5.get(os.path.basename(path))
    if want and md5sum(path) != want:
        raise RuntimeError("md5 mismatch for %s" % path)


def extract_train(tar_path: str, out: str) -> int:
    n = 0
    with tarfile.open(tar_path) as outer:
        for member in outer:
            if not member.name.endswith(".tar"):
                continue
            wnid = os.path.splitext(os.path.basename(member.name))[0]
            d = os.path.join(out, "train", wnid)
            os.makedirs(d, exist_ok=True)
            inner = tarfile.open(fileobj=io.BytesIO(outer.extractfile(member).read()))
            inner.extractall(d)
            n += len(inner.getnames())
    return n


def val_labels(devkit: str):
    """validation ground truth (ILSVRC2012 ids, 1-based) -> wnids via the devkit's meta.mat."""
    import scipy.io
    with tarfile.open(devkit) as t:
        gt = t.extractfile([m for m in t if m.name.endswith("validation_ground_truth.txt")][0]).read()
        meta = t.extractfile([m for m in t if m.name.endswith("meta.mat")][0]).read()
    syn = scipy.io.loadmat(io.BytesIO(meta), squeeze_me=True)["synsets"]
    id2wnid = {int(s[0]): str(s[1]) for s in syn if int(s[4]) == 0}
    return [id2wnid[int(x)] for x in gt.decode().split()]


def extract_val(tar_path: str, devkit: str, out: str) -> int:
    labels = val_labels(devkit)
    with tarfile.open(tar_path) as t:
        members = sorted((m for m in t if m.isfile()), key=lambda m: m.name)
        for m, wnid in zip(members, labels):
            d = os.path.join(out, "val", wnid)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, os.path.basename(m.name)), "wb") as f:
                f.write(t.extractfile(m).read())
    return len(members)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--train")
    ap.add_argument("--val")
    ap.add_argument("--devkit")
    ap.add_argument("--out", required=True)
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args(argv)
    for p in (a.train, a.val, a.devkit):
        if p and not a.no_verify:
            verify(p)
    if a.train:
        print("train images:", extract_train(a.train, a.out))
    if a.val:
        if not a.devkit:
            raise SystemExit("--val needs --devkit for the labels")
        print("val images:", extract_val(a.val, a.devkit, a.out))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
