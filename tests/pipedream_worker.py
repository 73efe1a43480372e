"""Worker for the multi-process PipeDream / GPipe tests (gloo on CPU, nccl on GPU)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200.models import build_model  # noqa: E402
from ddlbench_b200.parallel.pipedream.config import PipelineConfig  # noqa: E402
from ddlbench_b200.parallel.pipedream.runtime import StageRuntime  # noqa: E402


def even_cuts(model, n):
    b = model.block_boundaries()
    return [b[len(b) * k // n] + 1 for k in range(1, n)]


def main():
    reps = [int(x) for x in os.environ["PD_REPS"].split(",")]
    arch = os.environ.get("PD_ARCH", "resnet18")
    nit = int(os.environ.get("PD_ITERS", "6"))
    out = os.environ["PD_OUT"]
    gpu = torch.cuda.is_available() and os.environ.get("PD_CPU", "0") != "1"
    rank = int(os.environ["RANK"])
    if gpu:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl" if gpu else "gloo")
    torch.manual_seed(0)
    model = build_model(arch, "cifar10")
    cfg = PipelineConfig.hybrid(reps)
    cuts = even_cuts(model, len(reps))
    dev = f"cuda:{torch.cuda.current_device()}" if gpu else "cpu"
    kw = dict(no_input_pipelining=os.environ.get("PD_NOPIPE", "0") == "1",
              macrobatch=os.environ.get("PD_MACRO", "0") == "1",
              fused_replica_allreduce=os.environ.get("PD_FUSED", "1") == "1")
    rt = StageRuntime(model, cuts, cfg, batch_size=8, device=dev,
                      dtype=torch.bfloat16 if gpu else torch.float32, lr=0.05, momentum=0.9,
                      weight_decay=1e-4, **kw)
    topo = rt.topo
    n = cfg.num_iterations(topo.stage, nit)
    g = torch.Generator().manual_seed(100 + topo.rank_in_stage)
    data = [(torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g))
            for _ in range(n)]
    losses = []
    rt.run_epoch(iter(data) if topo.is_first else None, n,
                 on_forward=lambda i, s: losses.append(float(s[0])) if s is not None else None)
    ne = cfg.num_iterations(topo.stage, 2)
    rt.run_eval(iter(data) if topo.is_first else None, ne)
    sd = rt.ex.state_dict()
    torch.save({"rank": rank, "stage": topo.stage, "losses": losses, "state": sd,
                "warmup": rt.warmup, "versions": rt.ex.store.versions,
                "bytes_sent": rt.p2p.bytes_sent, "graph_replays": rt.graph_replays,
                "graph_slots": len(rt._slots), "fused_replicas": rt.fdp is not None, "grad_scale": float(rt.ex.store.hyper[3]),
                "update_interval": rt.update_interval}, os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
