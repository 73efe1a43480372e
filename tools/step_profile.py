"""One eager training step of ResNet-50/ImageNet-shape under cudaProfilerStart/Stop — wrap with
   ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ... to get the launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddlbench_b200.harness import Trainer
from ddlbench_b200.data.synthetic import SyntheticImages
tc = os.environ.get("TCGEN05", "1") == "1"
bs = int(os.environ.get("BS", "32"))
tr = Trainer("resnet50", "imagenet", batch_size=bs, device="cuda", dtype=torch.bfloat16, cuda_graph=False, tcgen05=tc)
src = SyntheticImages("imagenet", bs, num_batches=1, device="cuda")
x, t = src.batch(0)
for _ in range(3):
    tr.step_u8(x, t, src.mean, src.std)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step_u8(x, t, src.mean, src.std)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
