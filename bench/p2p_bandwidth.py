"""Point-to-point bandwidth sweep between rank 0 and rank 1 — the micro-benchmark of
``pipedream-fork/runtime/tests/communication/point_to_point.py:13-98`` (there: ``dist.send/recv``
of CPU copies or ``dist.broadcast`` in a 2-rank group over Gloo, host wall clock).  Here:
NCCL p2p (``DistP2P``), our peer-store kernel (``p2p_put`` over symmetric memory) and the
2-rank broadcast, device-timed, message sizes 40 B ... 3.2 GB.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench/p2p_bandwidth.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402
from ddlbench_b200.dist.symm import SymmContext  # noqa: E402
from ddlbench_b200.parallel import dp  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dp.init()
    rank = dist.get_rank()
    gpu = torch.cuda.is_available()
    dev = torch.device("cuda", torch.cuda.current_device()) if gpu else torch.device("cpu")
    sizes = [10, 1000, 100_000, 10_000_000, 100_000_000] + ([800_000_000] if gpu else [])
    rows = []
    ctx = SymmContext(dev) if gpu else None
    for n in sizes:
        t = torch.ones(n, device=dev)
        iters = 20 if n <= 10_000_000 else 5

        def sendrecv():
            if rank == 0:
                dist.send(t, dst=1)
            elif rank == 1:
                dist.recv(t, src=0)

        def bcast():
            dist.broadcast(t, src=0)

        res = {"floats": n, "bytes": n * 4}
        for name, fn in (("send_recv", sendrecv), ("broadcast", bcast)):
            ms = timed(fn, iters) if gpu else None
            if gpu:
                res[name + "_ms"] = round(ms, 4)
                res[name + "_GBps"] = round(n * 4 / ms / 1e6, 2)
        if gpu and n * 4 % 16 == 0 and n * 4 <= (1 << 30):
            buf = ctx.alloc(n, torch.float32)
            flag = ctx.alloc(4, torch.int32)
            seq = torch.zeros(1, dtype=torch.int32, device=dev)
            tick = torch.zeros(1, dtype=torch.int32, device=dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)

            def put():
                if rank == 0:
                    ops.ext().p2p_put(t, buf.ptrs[1], flag.ptrs[1], seq, tick, 32)
                elif rank == 1:
                    ops.ext().p2p_wait(flag.ptrs[1], seq, status)

            ms = timed(put, iters)
            res["p2p_put_ms"] = round(ms, 4)
            res["p2p_put_GBps"] = round(n * 4 / ms / 1e6, 2)
        rows.append(res)
        if rank == 0:
            print(json.dumps(res), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open("gpurun_out/p2p_bandwidth.json", "w"), indent=1)
    dist.barrier()
    dp.shutdown()


if __name__ == "__main__":
    main()
