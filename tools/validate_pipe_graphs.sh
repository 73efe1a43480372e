#!/bin/bash
# A/B of the graph-replayed PipeDream stages (DDLB_PIPE_GRAPHS=1) against the eager pipeline, plus
# a numerics check (the same seeds must give the same losses up to bf16 atomics order).
# Usage:  gpurun --gpus 4 --timeout 900 -- 'tools/validate_pipe_graphs.sh'      (charged 4x)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/pipe_graphs.log
: > $O
N=${N:-4}
for g in 0 1; do
  echo "=== worker test, DDLB_PIPE_GRAPHS=$g" | tee -a $O
  rm -rf /tmp/pg$g && mkdir -p /tmp/pg$g
  PD_REPS=$(python -c "print(','.join(['1']*$N))") PD_OUT=/tmp/pg$g PD_ITERS=24 DDLB_PIPE_GRAPHS=$g \
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
    --master-port $((29810 + g)) tests/pipedream_worker.py < /dev/null 2>&1 | tail -4 | tee -a $O
done
timeout 120 python - <<'PY' 2>&1 | tee -a $O
import glob, torch
for r in range(len(glob.glob("/tmp/pg0/rank*.pt"))):
    a, b = torch.load("/tmp/pg0/rank%d.pt" % r), torch.load("/tmp/pg1/rank%d.pt" % r)
    if a["losses"]:
        d = max(abs(x - y) for x, y in zip(a["losses"], b["losses"]))
        print("stage", a["stage"], "max |loss diff|", d, "replays", b["graph_replays"], "slots", b["graph_slots"])
        print("numerics", "OK" if d < 0.05 and b["graph_replays"] > 0 else "FAIL")
PY
echo "=== peer-store transport (DDLB_P2P=symm): worker test" | tee -a $O
rm -rf /tmp/pgs && mkdir -p /tmp/pgs
PD_REPS=$(python -c "print(','.join(['1']*$N))") PD_OUT=/tmp/pgs PD_ITERS=24 DDLB_P2P=symm \
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
  --master-port 29812 tests/pipedream_worker.py < /dev/null 2>&1 | tail -4 | tee -a $O
timeout 120 python - <<'PY' 2>&1 | tee -a $O
import glob, torch
for r in range(len(glob.glob("/tmp/pg0/rank*.pt"))):
    a, b = torch.load("/tmp/pg0/rank%d.pt" % r), torch.load("/tmp/pgs/rank%d.pt" % r)
    if a["losses"]:
        d = max(abs(x - y) for x, y in zip(a["losses"], b["losses"]))
        print("symm transport: stage", a["stage"], "max |loss diff| vs NCCL", d, "OK" if d < 0.05 else "FAIL")
PY
for g in 0 1; do
  echo "=== ResNet-50 ImageNet, ${N} stages, DDLB_PIPE_GRAPHS=$g" | tee -a $O
  DDLB_PIPE_GRAPHS=$g STEPS=120 EVAL_STEPS=2 EPOCHS=2 timeout 400 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29820 + g)) -m ddlbench_b200.cli \
    -b imagenet -f pipedream -g $N -m resnet50 --stages ${N}x1 --worker < /dev/null 2>&1 \
    | grep -E "device-timed|capture|Traceback|Error" | tail -6 | tee -a $O
done
echo "=== ResNet-50 ImageNet, ${N} stages, DDLB_P2P=symm" | tee -a $O
DDLB_P2P=symm STEPS=120 EVAL_STEPS=2 EPOCHS=2 timeout 400 python -m torch.distributed.run --nnodes=1 \
  --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29823 -m ddlbench_b200.cli \
  -b imagenet -f pipedream -g $N -m resnet50 --stages ${N}x1 --worker < /dev/null 2>&1 \
  | grep -E "device-timed|timed out|Traceback|Error" | tail -6 | tee -a $O
