"""Per-parameter gradient comparison: GPU engine (bf16 kernels) vs CPU fp32 engine vs torch bf16 autograd."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ddlbench_b200.engine import StageExecutor
from ddlbench_b200.models import build_model
torch.manual_seed(0)
arch, ds = sys.argv[1] if len(sys.argv) > 1 else "resnet18", "cifar10"
m = build_model(arch, ds)
cpu = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
gpu = StageExecutor.from_model(m, device="cuda", dtype=torch.bfloat16, tcgen05=False)
x = torch.randn(16, *m.input_shape); t = torch.randint(0, 10, (16,))
oc, cc = cpu.forward([cpu.prepare_input(x)]); sc, dc = cpu.loss(oc[0], t); cpu.backward(cc, [dc])
og, cg = gpu.forward([gpu.prepare_input(x)]); sg, dg = gpu.loss(og[0], t.cuda()); gpu.backward(cg, [dg])
# torch autograd bf16 channels_last
mb = build_model(arch, ds); mb.load_state_dict(m.state_dict()); mb = mb.cuda().to(torch.bfloat16).to(memory_format=torch.channels_last).train()
yb = mb(x.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)); lb = F.cross_entropy(yb.float(), t.cuda()); lb.backward()
print("loss cpu %.4f gpu-engine %.4f torch-bf16 %.4f" % (sc[0].item(), sg[0].item(), lb.item()))
print("logits err engine %.4f torch %.4f" % ((og[0].float().cpu()-oc[0]).abs().max().item(), (yb.float().cpu()-oc[0]).abs().max().item()))
tb = dict(mb.named_parameters())
for name in cpu.store.layout.order:
    r = cpu.store.g_views[name]; a = gpu.store.g_views[name].float().cpu(); b = tb[name].grad.float().cpu()
    den = r.norm().item() + 1e-12
    print("%-40s ref|g|=%.3e  engine_err=%.3f  torchbf16_err=%.3f" % (name, den, (a-r).norm().item()/den, (b-r).norm().item()/den))
