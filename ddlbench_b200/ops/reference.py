"""Plain PyTorch fp32 reference implementations of every native op.

Used (a) as the CPU execution path so the whole framework is testable on a GPU-less box
and (b) as the numerical oracle the CUDA kernels are compared against in ``tests/`` (the
task's rule: a kernel test compares against a plain PyTorch fp32 reference of the same op).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2


def _up(t: torch.Tensor) -> torch.Tensor:
    """Compute dtype: at least fp32 (fp64 inputs stay fp64 so tests can separate bugs from rounding)."""
    return t if t.dtype in (torch.float32, torch.float64) else t.to(torch.float32)


def act_fwd(x: torch.Tensor, act: int) -> torch.Tensor:
    if act == ACT_RELU:
        return torch.relu(x)
    if act == ACT_RELU6:
        return torch.clamp(x, 0.0, 6.0)
    return x


def act_mask(y: torch.Tensor, act: int) -> torch.Tensor:
    if act == ACT_RELU:
        return (y > 0).to(y.dtype)
    if act == ACT_RELU6:
        return ((y > 0) & (y < 6)).to(y.dtype)
    return torch.ones_like(y)


def sgd_flat(master, mom, grad, hyper, w_outs):
    """hyper = [lr, momentum, weight_decay, grad_scale]; updates master/mom in place and
    writes the new weights (cast) into every tensor of ``w_outs``."""
    lr, mu, wd, gs = [float(v) for v in hyper.tolist()[:4]]
    g = _up(grad) * gs + wd * master
    mom.mul_(mu).add_(g)
    master.add_(mom, alpha=-lr)
    for w in w_outs:
        if w is not None:
            w.copy_(master.to(w.dtype))


def xent(logits: torch.Tensor, target: torch.Tensor, smoothing: float = 0.0,
         grad_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (stats[3] = mean loss, #top1, #top5; dlogits = d(mean loss)/dlogits * grad_scale)."""
    x = _up(logits)
    n, c = x.shape
    logp = F.log_softmax(x, dim=1)
    nll = -logp.gather(1, target.view(-1, 1)).squeeze(1)
    smooth = -logp.mean(dim=1)
    loss = ((1.0 - smoothing) * nll + smoothing * smooth).mean()
    p = logp.exp()
    onehot = torch.zeros_like(p).scatter_(1, target.view(-1, 1), 1.0)
    d = (p - (1.0 - smoothing) * onehot - smoothing / c) * (grad_scale / 1.0)
    xt = x.gather(1, target.view(-1, 1))
    greater = (x > xt).sum(dim=1)
    top1 = (greater == 0).sum().float()
    top5 = (greater < 5).sum().float()
    stats = torch.stack([loss, top1, top5]).to(torch.float32)
    return stats, d.to(logits.dtype)


def _cview(t: torch.Tensor) -> Tuple[int, ...]:
    return (1, -1) + (1,) * (t.dim() - 2)


def bn_stat(x: torch.Tensor, eps: float) -> torch.Tensor:
    """[2, C] = (mean, invstd) over all dims but channel (dim 1)."""
    xf = _up(x)
    dims = [0] + list(range(2, x.dim()))
    mean = xf.mean(dim=dims)
    var = xf.var(dim=dims, unbiased=False)
    return torch.stack([mean, torch.rsqrt(var + eps)])


def bn_apply(x, res, stat, gamma, beta, act: int):
    v = _cview(x)
    y = (_up(x) - stat[0].view(v)) * (stat[1] * _up(gamma)).view(v) + _up(beta).view(v)
    if res is not None:
        y = y + _up(res)
    return act_fwd(y, act).to(x.dtype)


def bn_update_running(stat, x_numel_per_channel: int, eps: float, momentum: float, rmean, rvar):
    mean = stat[0]
    var = 1.0 / (stat[1] * stat[1]) - eps
    m = x_numel_per_channel
    unbiased = var * (m / max(m - 1, 1))
    rmean.add_(momentum * (mean - rmean))
    rvar.add_(momentum * (unbiased - rvar))


def bn_backward(dy, dy2, y, x, stat, gamma, act: int, want_dz: bool):
    """Returns (dx, dz or None, dgamma, dbeta) — all fp32 except dx/dz in x.dtype."""
    v = _cview(x)
    d = _up(dy)
    if dy2 is not None:
        d = d + _up(dy2)
    if act != ACT_NONE:
        d = d * act_mask(_up(y), act)
    dims = [0] + list(range(2, x.dim()))
    m = x.numel() // x.shape[1]
    xhat = (_up(x) - stat[0].view(v)) * stat[1].view(v)
    dbeta = d.sum(dim=dims)
    dgamma = (d * xhat).sum(dim=dims)
    dx = (_up(gamma) * stat[1]).view(v) * (d - dbeta.view(v) / m - xhat * dgamma.view(v) / m)
    return dx.to(x.dtype), (d.to(x.dtype) if want_dz else None), dgamma, dbeta


def bias_act(x, bias, act: int):
    y = _up(x)
    if bias is not None:
        y = y + _up(bias).view(_cview(x))
    return act_fwd(y, act).to(x.dtype)


def act_bwd(dy, dy2, y, act: int):
    d = _up(dy)
    if dy2 is not None:
        d = d + _up(dy2)
    if act != ACT_NONE:
        d = d * act_mask(_up(y), act)
    return d.to(dy.dtype)


def u8_to_norm(x_u8: torch.Tensor, mean, std, dtype) -> torch.Tensor:
    """x_u8: [N, H, W, C] uint8 -> logical [N, C, H, W] tensor (channels-last memory)."""
    c = x_u8.shape[-1]
    m = torch.tensor(mean[:c], dtype=torch.float32, device=x_u8.device)
    s = torch.tensor(std[:c], dtype=torch.float32, device=x_u8.device)
    y = (_up(x_u8) / 255.0 - m) / s
    return y.to(dtype).permute(0, 3, 1, 2)


def revert_varlen(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """out[t, b] = x[len_b - 1 - t, b] for t < len_b else 0 (x: [T, B, ...])."""
    T, B = x.shape[0], x.shape[1]
    t = torch.arange(T, device=x.device).view(T, 1)
    ln = lengths.view(1, B).to(x.device)
    src = (ln - 1 - t).clamp(min=0)
    valid = t < ln
    idx = src.view(T, B, *([1] * (x.dim() - 2))).expand_as(x)
    out = torch.gather(x, 0, idx)
    return out * valid.view(T, B, *([1] * (x.dim() - 2))).to(x.dtype)


def set_mask(lengths: torch.Tensor, T: int) -> torch.Tensor:
    t = torch.arange(T, device=lengths.device).view(T, 1)
    return (t < lengths.view(1, -1)).to(torch.uint8)


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    return _up(x).mean(dim=(2, 3)).to(x.dtype)


def im2col_rows(x, kh: int, kw: int, stride: int, pad: int):
    """Patch matrix with K ordered (kh, kw, c), zero-padded to a multiple of 8 columns."""
    import torch.nn.functional as F
    n, c, h, w = x.shape
    cols = F.unfold(_up(x), (kh, kw), padding=pad, stride=stride)            # [N, c*kh*kw, L]
    L = cols.shape[-1]
    cols = cols.view(n, c, kh * kw, L).permute(0, 3, 2, 1).reshape(n * L, kh * kw * c)
    k = kh * kw * c
    kpad = (k + 7) // 8 * 8
    out = cols.new_zeros((n * L, kpad))
    out[:, :k] = cols
    return out.to(x.dtype)
