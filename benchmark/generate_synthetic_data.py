"""Write the synthetic JPEG trees (``$DATADIR/<SET>/{train,val}/class_k/img_i.JPEG``) — same role and
command line as the reference's script of this name (``benchmark/generate_synthetic_data.py:74-116``):

    DATADIR=/scratch python generate_synthetic_data.py mnist cifar10 imagenet highres

(every argument is a dataset name, the output directory comes from ``DATADIR``).  Extensions: ``all``
for the four sets, a first argument that is not a dataset name is taken as the output directory, and
``TRAIN_IMAGES`` / ``VAL_IMAGES`` cap the image counts.  The training harness does not need these
files (synthetic batches are generated in memory); they exist for the real-data code path (``-s``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200.data.synthetic import generate_jpeg_tree  # noqa: E402

NAMES = ["mnist", "cifar10", "imagenet", "highres"]


def main(argv):
    args = list(argv)
    datadir = os.environ.get("DATADIR", ".")
    if args and args[0] not in NAMES + ["all"]:
        datadir = args.pop(0)
    names = NAMES if (not args or "all" in args) else [a for a in NAMES if a in args]
    cap = {k: (int(os.environ[v]) if os.environ.get(v) else None)
           for k, v in (("train", "TRAIN_IMAGES"), ("val", "VAL_IMAGES"))}
    for name in names:
        print("Start %s" % name, flush=True)
        rate = generate_jpeg_tree(datadir, name, train=cap["train"], val=cap["val"], verbose=False)
        print("took | imgs / sec: %.1f" % rate, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
