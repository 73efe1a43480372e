"""Install the reference for the ``--impl reference`` arm.

``python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref /root/reference`` fails: "Directory '/root/reference' is not installable. Neither
'setup.py' nor 'pyproject.toml' found" — DDLBench is a tree of scripts, not a package.  The
"install" is therefore a verbatim copy of the tree into ``baseline/_ref`` (git-ignored; it ships
with the gpurun snapshot)."""
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
if __name__ == "__main__":
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git"))
    print("copied %s -> %s" % (SRC, DST))
