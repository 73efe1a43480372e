"""Turn a PipeDream-style run log into a per-epoch table.

Reads the lines ``Epoch <e>: <s> seconds`` and `` * Prec@1 <v> Prec@5 <v>`` (and, for logs
of this framework, ``device-timed | <x> samples/sec``) and prints
``Epoch# / Runtime[h] (cumulative) / Top-1 / Top-5 / Epoch[h]`` — the table the reference's
``pipedream-fork/runtime/scripts/process_output.py:43-65`` builds; a log that starts with
``=> loading checkpoint`` is treated as resumed (a validation precedes the first epoch)."""
from __future__ import annotations

import re
import sys
from typing import Dict, List


def parse(lines: List[str]) -> Dict[str, list]:
    epoch_time, top1, top5, dev = [], [], [], []
    resumed = len(lines) > 1 and "=> loading checkpoint" in lines[0]
    start_epoch = 0
    if resumed:
        m = re.search(r"epoch (\d+)\)", lines[1])
        start_epoch = int(m.group(1)) if m else 0
    seen_epochs = set()
    for ln in lines:
        m = re.search(r"Epoch (\d+): ([\d.]+) seconds", ln)
        if m and int(m.group(1)) not in seen_epochs:
            seen_epochs.add(int(m.group(1)))
            epoch_time.append(float(m.group(2)) / 3600.0)
            continue
        m = re.search(r"\* Prec@1 ([\d.]+) Prec@5 ([\d.]+)", ln)
        if m:
            top1.append(float(m.group(1)))
            top5.append(float(m.group(2)))
            continue
        m = re.search(r"device-timed \| ([\d.]+) samples/sec", ln)
        if m:
            dev.append(float(m.group(1)))
    return {"epoch_hours": epoch_time, "top1": top1, "top5": top5, "device_throughput": dev,
            "resumed": resumed, "start_epoch": start_epoch}


def table(parsed: Dict[str, list]) -> str:
    out = ["Epoch#\tRuntime\tTop-1\tTop-5\tEpoch"]
    run = 0.0
    for i, t in enumerate(parsed["epoch_hours"]):
        run += t
        a1 = parsed["top1"][i] if i < len(parsed["top1"]) else float("nan")
        a5 = parsed["top5"][i] if i < len(parsed["top5"]) else float("nan")
        out.append("%d\t%.3f\t%.3f\t%.3f\t%.3f" % (i + parsed["start_epoch"], run, a1, a5, t))
    return "\n".join(out)


def parse_gnmt(lines: List[str]) -> Dict[str, list]:
    """GNMT logs (``runtime/scripts/process_output_gnmt.py``): epoch times plus the BLEU score
    reported after each epoch (``Epoch <e>: BLEU <v>``) and the last training loss."""
    epoch_time, bleu, loss = [], [], []
    last_loss = float("nan")
    for ln in lines:
        m = re.search(r"Epoch (\d+): ([\d.]+) seconds", ln)
        if m:
            epoch_time.append(float(m.group(2)) / 3600.0)
            loss.append(last_loss)
            continue
        m = re.search(r"BLEU\s*=?\s*([\d.]+)", ln)
        if m:
            bleu.append(float(m.group(1)))
            continue
        m = re.search(r"Loss ([\d.]+)", ln)
        if m:
            last_loss = float(m.group(1))
    return {"epoch_hours": epoch_time, "bleu": bleu, "loss": loss}


def table_gnmt(parsed: Dict[str, list]) -> str:
    out = ["Epoch#\tRuntime\tBLEU\tLoss\tEpoch"]
    run = 0.0
    for i, t in enumerate(parsed["epoch_hours"]):
        run += t
        b = parsed["bleu"][i] if i < len(parsed["bleu"]) else float("nan")
        out.append("%d\t%.3f\t%.2f\t%.4f\t%.3f" % (i, run, b, parsed["loss"][i], t))
    return "\n".join(out)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    gnmt = "--gnmt" in argv
    argv = [x for x in argv if x != "--gnmt"]
    if len(argv) != 1:
        print("python -m ddlbench_b200.utils.process_output [--gnmt] <output.log>")
        return 1
    with open(argv[0]) as f:
        lines = [ln.rstrip("\n") for ln in f]
    print(table_gnmt(parse_gnmt(lines)) if gnmt else table(parse(lines)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
