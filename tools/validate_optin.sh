#!/bin/bash
# One-GPU validation pass for the opt-in paths.  Every step runs under its own
# timeout (a kernel protocol bug traps after ~2 s; the timeouts are a second fence) and writes to
# gpurun_out/.  Usage:  gpurun --timeout 1500 -- 'tools/validate_optin.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/optin
: > $O.log
step() { echo "=== $*" | tee -a $O.log; }

step "numerics: tcgen05 conv3x3 / stem im2col / wgrad (descriptor variant 0)"
DDLB_TEST_CONV_TCGEN05=1 timeout 400 python -m pytest tests/test_kernels_gpu.py -q -k "im2col or dwconv3x3 or conv3x3" < /dev/null 2>&1 | tail -3 | tee -a $O.log
BENCH=0 CHECK_WGRAD=0 CHECK_BNFOLD=0 CHECK_BNBWD=0 timeout 300 python tests/conv_check.py < /dev/null 2>&1 | tail -14 | tee -a $O.log
for v in 0 1; do
  step "wgrad GEMM, DDLB_WGRAD_DESC_VARIANT=$v"
  DDLB_WGRAD_DESC_VARIANT=$v timeout 200 python - <<'PY' 2>&1 | tail -9 | tee -a $O.log
import sys
sys.path.insert(0, "tests")
import conv_check as c
ok = all(c.check_wgrad(*a) for a in [(64, 128, 64), (128, 128, 128), (6272, 256, 1024), (1000, 72, 200)])
print("wgrad", "OK" if ok else "FAIL")
PY
done
step "BN-fold GEMM numerics"
BENCH=0 CHECK_WGRAD=0 timeout 300 python - <<'PY' 2>&1 | tail -8 | tee -a $O.log
import sys
sys.path.insert(0, "tests")
import conv_check as c
ok = all(c.check_bnfold(*a) for a in [(32, 64, 256, 56, 1, False), (32, 256, 64, 56, 1, True), (32, 512, 2048, 7, 1, False), (5, 128, 64, 9, 2, True)])
print("bnfold", "OK" if ok else "FAIL")
PY
step "BN-backward-fold GEMM numerics"
BENCH=0 timeout 300 python - <<'PY' 2>&1 | tail -8 | tee -a $O.log
import sys
sys.path.insert(0, "tests")
import conv_check as c
ok = all(c.check_bnbwd(*a) for a in [(32, 64, 256, 56, 1, True, True), (32, 256, 64, 56, 1, False, False), (32, 2048, 512, 7, 1, False, False), (5, 64, 128, 9, 2, True, False)])
print("bnbwd", "OK" if ok else "FAIL")
PY
step "timing vs cuDNN"
CHECK_WGRAD=0 CHECK_BNFOLD=0 CHECK_BNBWD=0 timeout 300 python tests/conv_check.py < /dev/null 2>&1 | grep "^{" | tee -a $O.log

bench() {  # name, env...
  local name=$1; shift
  step "bench $name"
  env "$@" timeout 240 python bench.py --steps 40 --warmup 5 < /dev/null 2>&1 | tail -1 > gpurun_out/bench_optin_$name.json
  python - "$name" <<'PY' | tee -a $O.log
import json, sys
try:
    d = json.load(open("gpurun_out/bench_optin_%s.json" % sys.argv[1]))
    print(sys.argv[1], d["value"], "img/s", d["ms_per_step"], "ms  loss", d["e2e"]["last_loss"], " launches/step", d["launches_per_step_in_graph"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
bench default DDLB_NONE=1
bench stem DDLB_STEM_IM2COL=1
bench conv3x3 DDLB_CONV3X3_TCGEN05=1
bench conv3x3_dgrad DDLB_CONV3X3_TCGEN05=1 DDLB_CONV3X3_DGRAD=1
bench wgrad DDLB_WGRAD_TCGEN05=1
bench sgd_overlap DDLB_SGD_OVERLAP=1
bench bn_fold DDLB_BN_FOLD=1
bench bn_cluster DDLB_BN_CLUSTER=1
bench bwd_apply_3blocks DDLB_BN_BWD_APPLY_BLOCKS=3
bench dgrad1x1 DDLB_DGRAD1X1_TCGEN05=1
bench bn_bwd_fold DDLB_BN_BWD_FOLD=1
bench bn_both_folds DDLB_BN_FOLD=1 DDLB_BN_BWD_FOLD=1
bench all_validated DDLB_STEM_IM2COL=1 DDLB_CONV3X3_TCGEN05=1 DDLB_SGD_OVERLAP=1
bench all_tcgen05 DDLB_STEM_IM2COL=1 DDLB_CONV3X3_TCGEN05=1 DDLB_CONV3X3_DGRAD=1 DDLB_WGRAD_TCGEN05=1
bench default_again DDLB_NONE=1
step "MobileNetV2 (CIFAR-10 shape) with and without the depthwise kernels"
for f in 0 1; do
  DDLB_DWCONV=$f timeout 200 python - <<'PY' 2>&1 | tail -1 | tee -a $O.log
import os, time, torch
from ddlbench_b200.harness import Trainer
tr = Trainer("mobilenetv2", "cifar10", batch_size=64, device="cuda", dtype=torch.bfloat16, cuda_graph=True, tcgen05=True)
x = torch.randn(64, 3, 32, 32, device="cuda"); t = torch.randint(0, 10, (64,), device="cuda")
for _ in range(5): tr.step(x, t)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): s = tr.step(x, t)
e1.record(); torch.cuda.synchronize()
print("mobilenetv2 DDLB_DWCONV=%s: %.3f ms/step loss %.4f" % (os.environ.get("DDLB_DWCONV"), e0.elapsed_time(e1) / 30, float(s[0])))
PY
done
step done
