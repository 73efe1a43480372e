import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200.models import build_model  # noqa: E402
from ddlbench_b200.parallel.gpipe import GPipe, balance_by_size  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist")
    pipe = GPipe(m, balance_by_size(3, m), chunks=4, devices=["cpu"] * 3, dtype=torch.float32,
                 lr=0.1, momentum=0.9, weight_decay=1e-4)
    pipe.set_micro_batch(4)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(16, 1, 28, 28, generator=g)
    t = torch.randint(0, 10, (16,), generator=g)
    losses = []
    for _ in range(3):
        s = pipe.train_step(x if rank == 0 else None, t if rank == 2 else None)
        if s is not None:
            losses.append(float(s[0]))
    pipe.set_micro_batch(16)
    pipe.eval_step(x if rank == 0 else None, t if rank == 2 else None)
    torch.save({"rank": rank, "losses": losses, "state": pipe.state_dict()},
               os.path.join(os.environ["GP_OUT"], f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
