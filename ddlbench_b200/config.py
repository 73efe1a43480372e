"""Typed run configuration.

The reference passes configuration through four stacked channels — bash getopts
(``run/run/run.sh:16-47``), positional args (``run/run/run_template.sh:14-28``),
environment variables read at import (``benchmark/imagenet/imagenet_pytorch.py:32-36``)
and per-script argparse.  Here there is one dataclass; the environment contract
(``DATADIR, EPOCHS, BATCH_SIZE, LOGINTER, CORES_GPU, MICROBATCHES``) is still
honoured so reference-style invocations keep working.
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass, field
from typing import Dict, Optional, Tuple

FRAMEWORKS = ("pytorch", "horovod", "gpipe", "pipedream")
BENCHMARKS = ("mnist", "cifar10", "imagenet", "highres")

# per-process batch sizes hard-coded by the reference job script
# (run/run/run_template.sh:186-201, 213-228, 244-263, 378-390)
BATCH_SIZES: Dict[str, Dict[str, int]] = {
    "pytorch": {"mnist": 128, "cifar10": 64, "imagenet": 32, "highres": 32},
    "horovod": {"mnist": 128, "cifar10": 64, "imagenet": 32, "highres": 32},
    # gpipe: BATCH_SIZE is the MICRO-batch (benchmark/imagenet/imagenet_gpipe.py:41-45)
    "gpipe": {"mnist": 128, "cifar10": 64, "imagenet": 24, "highres": 4},
    "pipedream": {"mnist": 512, "cifar10": 256, "imagenet": 128, "highres": 64},
}
MICROBATCHES: Dict[str, int] = {"mnist": 24, "cifar10": 32, "imagenet": 12, "highres": 12}

# (lr, momentum, weight_decay) per script — SURVEY Appendix A
HYPER: Dict[str, Dict[str, Tuple[float, float, float]]] = {
    "pytorch": {"mnist": (0.01, 0.5, 0.0), "cifar10": (0.1, 0.9, 5e-4),
                "imagenet": (0.1, 0.9, 1e-4), "highres": (0.1, 0.9, 1e-4)},
    # horovod: base lr is multiplied by world size at run time
    "horovod": {"mnist": (0.01, 0.5, 0.0), "cifar10": (0.1, 0.9, 5e-4),
                "imagenet": (0.0125, 0.9, 5e-5), "highres": (0.0125, 0.9, 5e-5)},
    "gpipe": {"mnist": (0.01, 0.5, 0.0), "cifar10": (0.001, 0.9, 0.0),
              "imagenet": (0.1, 0.9, 1e-4), "highres": (0.1, 0.9, 1e-4)},
    "pipedream": {"mnist": (0.1, 0.9, 1e-4), "cifar10": (0.1, 0.9, 1e-4),
                  "imagenet": (0.1, 0.9, 1e-4), "highres": (0.1, 0.9, 1e-4)},
}

# synthetic dataset sizes: train / val images (benchmark/generate_synthetic_data.py:76-107)
DATASET_SIZES: Dict[str, Tuple[int, int]] = {
    "mnist": (60000, 10000), "cifar10": (50000, 10000),
    "imagenet": (1281167, 50000), "highres": (50000, 10000),
}


def _env_int(name: str, default: Optional[int]) -> Optional[int]:
    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


@dataclass
class RunConfig:
    benchmark: str = "mnist"
    framework: str = "pytorch"
    arch: str = "resnet18"
    gpus: int = 1                 # GPUs per node (-g)
    nodes: int = 1                # -n
    log_interval: int = 25        # -p / LOGINTER
    epochs: int = 3               # EPOCHS (run_template.sh:71)
    batch_size: Optional[int] = None      # BATCH_SIZE (per process; micro-batch for gpipe)
    microbatches: Optional[int] = None    # MICROBATCHES (gpipe chunks)
    cores_gpu: int = 4            # CORES_GPU
    datadir: str = ""             # DATADIR
    synthetic: bool = True        # run.sh -s disables
    lr: Optional[float] = None
    momentum: Optional[float] = None
    weight_decay: Optional[float] = None
    dtype: str = "bf16"           # compute dtype on GPU
    device: str = "auto"          # "auto" | "cuda" | "cpu"
    steps_per_epoch: Optional[int] = None  # cap iterations/epoch (synthetic runs); None = full dataset
    eval_steps: Optional[int] = None
    cuda_graph: bool = True
    fused: bool = True            # use the hand-written sm_100a kernels where available
    allreduce: str = "fused"      # "fused" (peer-memory kernel + SGD) | "nccl" (baseline)
    seed: int = 1
    data_mode: str = "device"     # "device" | "pinned" (H2D every step) | "jpeg" (ImageFolder)
    extra: Dict[str, str] = field(default_factory=dict)

    def __post_init__(self):
        if self.framework not in FRAMEWORKS:
            raise ValueError(f"framework must be one of {FRAMEWORKS}")
        if self.benchmark not in BENCHMARKS:
            raise ValueError(f"benchmark must be one of {BENCHMARKS}")
        if self.batch_size is None:
            self.batch_size = BATCH_SIZES[self.framework][self.benchmark]
        if self.microbatches is None:
            self.microbatches = MICROBATCHES[self.benchmark] if self.framework == "gpipe" else 1
        lr, mom, wd = HYPER[self.framework][self.benchmark]
        if self.lr is None:
            self.lr = lr
        if self.momentum is None:
            self.momentum = mom
        if self.weight_decay is None:
            self.weight_decay = wd
        # reference rule: pytorch & gpipe are single-node (run/run/run.sh:51-54)
        if self.nodes > 1 and self.framework in ("pytorch", "gpipe"):
            self.nodes = 1

    # ------------------------------------------------------------------ env
    @classmethod
    def from_env(cls, **overrides) -> "RunConfig":
        """Build from the reference's environment contract (missing vars fall back
        to the run_template.sh defaults instead of raising KeyError)."""
        kw = dict(
            epochs=_env_int("EPOCHS", 3),
            batch_size=_env_int("BATCH_SIZE", None),
            log_interval=_env_int("LOGINTER", 25),
            cores_gpu=_env_int("CORES_GPU", 4),
            microbatches=_env_int("MICROBATCHES", None),
            datadir=os.environ.get("DATADIR", ""),
            # extensions (not in the reference): cap the iterations of an epoch / of validation
            steps_per_epoch=_env_int("STEPS", None),
            eval_steps=_env_int("EVAL_STEPS", None),
        )
        kw.update(overrides)
        return cls(**kw)

    def export_env(self) -> Dict[str, str]:
        env = {"EPOCHS": str(self.epochs), "BATCH_SIZE": str(self.batch_size),
               "LOGINTER": str(self.log_interval), "CORES_GPU": str(self.cores_gpu),
               "DATADIR": self.datadir}
        if self.framework == "gpipe":
            env["MICROBATCHES"] = str(self.microbatches)
        return env

    @property
    def world_size(self) -> int:
        return self.gpus * self.nodes if self.framework != "pytorch" else 1

    @property
    def global_batch(self) -> int:
        if self.framework == "gpipe":
            return self.batch_size * self.microbatches
        if self.framework == "horovod":
            return self.batch_size * self.world_size
        return self.batch_size

    def header(self) -> str:
        """``<fw> - <dataset> - <model> - batch=<B>`` (run/run/run_template.sh:187)."""
        return "%s - %s - %s - batch=%d" % (self.framework, self.benchmark, self.arch,
                                            self.batch_size)

    def to_dict(self):
        return asdict(self)
