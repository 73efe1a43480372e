"""Write the synthetic JPEG trees (``DATADIR/<SET>/{train,val}/class_k/img_i.JPEG``) — same role and
arguments as the reference's script of this name: ``python generate_synthetic_data.py <datadir>
[mnist|cifar10|imagenet|highres|all]``.  The training harness does not need them (synthetic batches
are generated in memory); they exist for the real-data code path (``-s``) and for parity."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200.data.synthetic import generate_jpeg_tree  # noqa: E402

if __name__ == "__main__":
    datadir = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("DATADIR", ".")
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    for name in (["mnist", "cifar10", "imagenet", "highres"] if which == "all" else [which]):
        generate_jpeg_tree(datadir, name)
