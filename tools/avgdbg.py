import torch, sys
sys.path.insert(0, '.')
from ddlbench_b200 import ops
torch.manual_seed(0)
n,c,hw,k,s,p,cip = 4,64,14,3,1,1,True
x = torch.randn(n,c,hw,hw,device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
xr = x.float().requires_grad_(True)
y_ref = torch.nn.functional.avg_pool2d(xr,k,s,p,False,cip)
dy = torch.randn_like(y_ref).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y_ref.backward(dy.float())
dx = ops.avgpool_bwd(dy, x.shape, k,s,p,False,cip)
xc = x.float().cpu().requires_grad_(True)
yc = torch.nn.functional.avg_pool2d(xc,k,s,p,False,cip)
yc.backward(dy.float().cpu())
print("torch gpu grad vs cpu grad", (xr.grad.cpu()-xc.grad).abs().max().item())
print("ours vs cpu grad", (dx.float().cpu()-xc.grad).abs().max().item())
print("dy strides", dy.stride(), "dx strides", dx.stride(), "xr.grad strides", xr.grad.stride())
print(dx[0,0,:3,:3].float().cpu(), xc.grad[0,0,:3,:3])
