"""In-tree build of the native extension (``ddlbench_b200/ops/_ddlb_ext*.so``).

Every ``.cu`` is compiled straight with ``nvcc -gencode arch=compute_100a,code=sm_100a
-lineinfo`` (cross-compiles on a GPU-less box in seconds because the kernels do not include
torch headers); ``bindings.cpp`` is compiled with the host compiler against torch and the
objects are linked into one shared object that lives next to this file so it travels with
the repo snapshot to the GPU box.  Objects are cached by content hash under ``build/``.

Usage:  ``python -m ddlbench_b200.ops.build [--force] [--verbose]``
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE.parent.parent / "build" / "ddlb_ext"
EXT_NAME = "_ddlb_ext"

CUDA_SOURCES = ["sgd.cu", "xent.cu", "bn.cu", "bn_cluster.cu", "allreduce.cu", "misc.cu", "pool.cu", "gemm_tcgen05.cu", "conv_tcgen05.cu", "im2col.cu",
                "gemm_wgrad_tcgen05.cu", "dwconv.cu",
                "gemm_bnfold_tcgen05.cu", "gemm_bnbwd_tcgen05.cu", "igemm_tcgen05.cu", "elementwise.cu",
                "gemm2cta_tcgen05.cu"]
CPP_SOURCES = ["bindings.cpp"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def cuda_home() -> str:
    for cand in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if cand and Path(cand, "bin", "nvcc").exists():
            return cand
    nvcc = shutil.which("nvcc")
    if nvcc:
        return str(Path(nvcc).resolve().parent.parent)
    raise RuntimeError("nvcc not found (set CUDA_HOME)")


def so_path() -> Path:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return HERE / f"{EXT_NAME}{suffix}"


def _hash(paths: List[Path], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def _headers() -> List[Path]:
    return sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")))


def _run(cmd: List[str], verbose: bool) -> str:
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("build step failed:\n  %s\n%s" % (" ".join(cmd), p.stdout))
    return p.stdout


def build(force: bool = False, verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as ce

    BUILD.mkdir(parents=True, exist_ok=True)
    home = cuda_home()
    nvcc = str(Path(home, "bin", "nvcc"))
    hdrs = _headers()
    objs: List[Path] = []
    jobs = []

    def cu_job(src: Path):
        tag = _hash([src] + hdrs, " ".join(NVCC_FLAGS))
        obj = BUILD / f"{src.stem}.{tag}.o"
        if force or not obj.exists():
            out = _run([nvcc] + NVCC_FLAGS + ["-I", str(CSRC), "-c", str(src), "-o", str(obj)],
                       verbose)
            (BUILD / f"{src.stem}.ptxas.txt").write_text(out)
        return obj

    torch_inc = ce.include_paths()
    py_inc = sysconfig.get_paths()["include"]
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-DTORCH_EXTENSION_NAME={EXT_NAME}",
                 "-DTORCH_API_INCLUDE_EXTENSION_H",
                 f"-D_GLIBCXX_USE_CXX11_ABI={int(torch.compiled_with_cxx11_abi())}",
                 "-Wno-deprecated-declarations"]

    def cpp_job(src: Path):
        tag = _hash([src] + hdrs, " ".join(cxx_flags) + torch.__version__)
        obj = BUILD / f"{src.stem}.{tag}.o"
        if force or not obj.exists():
            cmd = [os.environ.get("CXX", "g++")] + cxx_flags
            for inc in torch_inc + [py_inc, str(Path(home, "include")), str(CSRC)]:
                cmd += ["-isystem", inc]
            cmd += ["-c", str(src), "-o", str(obj)]
            _run(cmd, verbose)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for s in CUDA_SOURCES:
            jobs.append(ex.submit(cu_job, CSRC / s))
        for s in CPP_SOURCES:
            jobs.append(ex.submit(cpp_job, CSRC / s))
        objs = [j.result() for j in jobs]

    out = so_path()
    link_tag = _hash(objs, "link")
    stamp = BUILD / "link.stamp"
    if force or not out.exists() or not stamp.exists() or stamp.read_text() != link_tag:
        torch_lib = str(Path(torch.__file__).parent / "lib")
        cmd = [os.environ.get("CXX", "g++"), "-shared", "-o", str(out)] + [str(o) for o in objs]
        cmd += ["-L", torch_lib, "-L", str(Path(home, "lib64")),
                "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
                "-lcudart", f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{Path(home, 'lib64')}"]
        _run(cmd, verbose)
        stamp.write_text(link_tag)
    # drop objects of older source revisions (the cache is keyed by content hash, so they only
    # accumulate and inflate the snapshot that travels to the GPU box)
    keep = {o.name for o in objs}
    for old in BUILD.glob("*.o"):
        if old.name not in keep:
            try:
                old.unlink()
            except OSError:
                pass
    return out


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    p = build(force="--force" in argv, verbose="--verbose" in argv)
    print(p)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
