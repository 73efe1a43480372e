"""Stage / rank configuration of a PipeDream run.

File format = the reference's ``{dp,mp,hybrid}_conf.json``
(``pipedream-fork/optimizer/templates/conf.json.template``; consumed at
``runtime/image_classification/main_with_runtime.py:194-205``):

    {"module_to_stage_map": [0, 1, 2, 3, 3],               # one entry per module incl. criterion
     "stage_to_rank_map": {"0": [0], "1": [1], ...},        # keys are strings
     "stage_to_depth_map": {"0": 3, ...},                   # optional warm-up override
     "parallel": {"tp": 1, "cp": 1}}                        # optional, RESERVED (see below)

``parallel`` reserves the two axes the reference does not have (its models
are CNNs and an LSTM, there is no tensor- or context-parallel code anywhere in the tree): ``tp``
(tensor parallel width inside a stage) and ``cp`` (context / sequence parallel width).  The schema
carries them so that a later runtime can add either without a file-format change; today any value
other than 1 is rejected loudly when the file is loaded.

Topology rules (``runtime/runtime.py:133-176``): a rank's stage, its index among the stage's
replicas, the replica lists of the neighbouring stages, and the number of warm-up minibatches

    warmup(stage) = (num_ranks - 1 - sum(replicas of earlier stages)) // replicas(stage)

(or the depth map's value).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from math import gcd
from typing import Dict, List, Optional, Sequence


@dataclass
class PipelineConfig:
    module_to_stage_map: List[int]
    stage_to_rank_map: Dict[int, List[int]]
    stage_to_depth_map: Optional[Dict[int, int]] = None
    tp: int = 1      # reserved: tensor-parallel width inside a stage (only 1 is implemented)
    cp: int = 1      # reserved: context-parallel width (only 1 is implemented)

    def __post_init__(self):
        if self.tp != 1 or self.cp != 1:
            raise NotImplementedError(
                "parallel.tp=%d / parallel.cp=%d: tensor and context parallelism are reserved fields of "
                "the configuration schema; this runtime implements pipeline x data parallelism only "
                "(tp = cp = 1)" % (self.tp, self.cp))

    # ------------------------------------------------------------------ io
    @classmethod
    def from_json(cls, path_or_dict) -> "PipelineConfig":
        d = path_or_dict
        if not isinstance(d, dict):
            with open(path_or_dict) as f:
                d = json.load(f)
        s2r = {int(k): list(v) for k, v in d["stage_to_rank_map"].items()}
        depth = d.get("stage_to_depth_map")
        depth = {int(k): int(v) for k, v in depth.items()} if depth else None
        par = d.get("parallel") or {}
        return cls(list(d["module_to_stage_map"]), s2r, depth, int(par.get("tp", 1)), int(par.get("cp", 1)))

    def to_dict(self) -> dict:
        d = {"module_to_stage_map": list(self.module_to_stage_map),
             "stage_to_rank_map": {str(k): list(v) for k, v in sorted(self.stage_to_rank_map.items())}}
        if self.stage_to_depth_map:
            d["stage_to_depth_map"] = {str(k): v for k, v in sorted(self.stage_to_depth_map.items())}
        if (self.tp, self.cp) != (1, 1):
            d["parallel"] = {"tp": self.tp, "cp": self.cp}
        return d

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=4)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def data_parallel(cls, num_modules: int, num_ranks: int) -> "PipelineConfig":
        """dp_conf: every module in stage 0, replicated on all ranks."""
        return cls([0] * num_modules, {0: list(range(num_ranks))})

    @classmethod
    def straight(cls, num_stages: int) -> "PipelineConfig":
        """mp_conf: stage i on rank i (criterion shares the last stage)."""
        return cls(list(range(num_stages)) + [num_stages - 1],
                   {i: [i] for i in range(num_stages)})

    @classmethod
    def hybrid(cls, stage_to_num_ranks: Sequence[int]) -> "PipelineConfig":
        """hybrid_conf from per-stage replica counts (``--stage_to_num_ranks_map "0:3,1:1"``)."""
        s2r, r = {}, 0
        for s, n in enumerate(stage_to_num_ranks):
            s2r[s] = list(range(r, r + n))
            r += n
        ns = len(stage_to_num_ranks)
        return cls(list(range(ns)) + [ns - 1], s2r)

    @staticmethod
    def parse_stage_to_num_ranks(text: str) -> List[int]:
        pairs = sorted((int(a), int(b)) for a, b in (kv.split(":") for kv in text.split(",")))
        return [n for _, n in pairs]

    # ------------------------------------------------------------------ topology
    @property
    def num_stages(self) -> int:
        return len(self.stage_to_rank_map)

    @property
    def num_ranks(self) -> int:
        return sum(len(v) for v in self.stage_to_rank_map.values())

    def stage_of_rank(self, rank: int) -> int:
        for s, ranks in self.stage_to_rank_map.items():
            if rank in ranks:
                return s
        raise ValueError(f"rank {rank} is not in the configuration")

    def replicas(self, stage: int) -> int:
        return len(self.stage_to_rank_map[stage])

    def warmup(self, stage: int) -> int:
        if self.stage_to_depth_map is not None:
            return self.stage_to_depth_map[stage]
        w = self.num_ranks - 1
        for s in range(stage):
            w -= self.replicas(s)
        return w // self.replicas(stage)

    def modules_of_stage(self, stage: int) -> List[int]:
        return [i for i, s in enumerate(self.module_to_stage_map) if s == stage]

    def lr_scale(self, stage: int) -> float:
        """Per-stage LR scaling (``runtime/runtime.py:695-701``)."""
        return 1.0 if stage == 0 else self.replicas(stage) / self.replicas(0)

    def num_iterations(self, stage: int, loader_size: int) -> int:
        """Iterations a replica of ``stage`` runs per epoch when a first-stage replica draws
        ``loader_size`` minibatches: rounded down so every downstream replica count divides
        the total (DDLBench addition, ``runtime/runtime.py:663-690``)."""
        counts = [self.replicas(s) for s in sorted(self.stage_to_rank_map)]
        if len(counts) >= 2:
            # The reference rounds to lcm(replica counts of stages 1..) — enough for the
            # configurations DDLBench generates but not in general (a stage replica deals its
            # forwards round-robin to the next stage's replicas, so its own iteration count
            # must also be divisible by that stage's replica count; `schedule.validate` finds
            # e.g. [1,1,2,4] deadlocking under the reference rule).  Round so that
            # loader_size * n_0 is divisible by n_s * n_{s+1} for every boundary.
            lcm = 1
            n0 = counts[0]
            for a, b in zip(counts[:-1], counts[1:]):
                need = (a * b) // gcd(a * b, n0)
                lcm = lcm * need // gcd(lcm, need)
            for c in counts[1:]:
                need = c // gcd(c, n0)
                lcm = lcm * need // gcd(lcm, need)
            loader_size = (loader_size // lcm) * lcm
        if stage == 0:
            return loader_size
        total = loader_size * self.replicas(0)
        assert total % self.replicas(stage) == 0
        return total // self.replicas(stage)


@dataclass
class RankTopology:
    rank: int
    stage: int
    rank_in_stage: int
    num_ranks_in_stage: int
    prev_ranks: List[int]
    next_ranks: List[int]
    warmup: int
    is_first: bool
    is_last: bool
    stage_ranks: List[int]

    @classmethod
    def of(cls, cfg: PipelineConfig, rank: int) -> "RankTopology":
        s = cfg.stage_of_rank(rank)
        ranks = cfg.stage_to_rank_map[s]
        return cls(rank, s, ranks.index(rank), len(ranks),
                   list(cfg.stage_to_rank_map[s - 1]) if s > 0 else [],
                   list(cfg.stage_to_rank_map[s + 1]) if s < cfg.num_stages - 1 else [],
                   cfg.warmup(s), s == 0, s == cfg.num_stages - 1, list(ranks))
