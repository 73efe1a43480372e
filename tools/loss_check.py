import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddlbench_b200.harness import Trainer
from ddlbench_b200.data.synthetic import SyntheticImages
for graph in (False, True):
    tr = Trainer("resnet50", "imagenet", batch_size=8, device="cuda", dtype=torch.bfloat16, cuda_graph=graph, tcgen05=os.environ.get("TC","1")=="1")
    src = SyntheticImages("imagenet", 8, num_batches=2, device="cuda")
    ls=[]
    for i in range(8):
        x, t = src.batch(0)
        ls.append(round(float(tr.step_u8(x, t, src.mean, src.std)[0]),3))
    print("overlap", os.environ.get("DDLB_WGRAD_OVERLAP","1"), "fusefin", os.environ.get("DDLB_FUSE_FINALIZE","0"), "tc", os.environ.get("TC","1"), "graph", graph, ls, flush=True)
