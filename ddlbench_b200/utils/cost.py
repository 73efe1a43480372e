"""Analytic cost of one training step of a layer-graph model (for roofline statements): forward
FLOPs of the convolutions / linear layers from a meta-device dry run (2 FLOPs per MAC; backward
= 2x forward: data gradient + weight gradient), and the size of every layer output."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


def model_cost(model: nn.Module, batch: int) -> Dict[str, float]:
    macs = 0
    out_elems = 0
    gemm_out_elems = 0
    hooks = []

    def hook(m, _inp, out):
        nonlocal macs, out_elems, gemm_out_elems
        if not torch.is_tensor(out):
            return
        out_elems += out.numel()
        if isinstance(m, nn.Conv2d):
            kh, kw = m.kernel_size
            macs += out.numel() * (m.in_channels // m.groups) * kh * kw
            gemm_out_elems += out.numel()
        elif isinstance(m, nn.Linear):
            macs += out.numel() * m.in_features
            gemm_out_elems += out.numel()

    for m in model.modules():
        if not list(m.children()):
            hooks.append(m.register_forward_hook(hook))
    try:
        from torch._subclasses.fake_tensor import FakeTensorMode
        with FakeTensorMode(allow_non_fake_inputs=True), torch.no_grad():
            x = torch.empty((batch,) + tuple(model.input_shape))
            was = model.training
            model.eval()
            model(x)
            model.train(was)
    finally:
        for h in hooks:
            h.remove()
    params = sum(p.numel() for p in model.parameters())
    return {"forward_flops": 2.0 * macs, "step_flops": 6.0 * macs, "layer_output_elements": float(out_elems),
            "conv_output_elements": float(gemm_out_elems), "parameters": float(params)}


def measured_peaks(path: Optional[str] = None) -> Dict[str, object]:
    path = path or os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured"}
    except Exception:
        return dict(FALLBACK_PEAKS)


def roofline(model: nn.Module, batch: int, ms_per_step: float) -> Dict[str, object]:
    """Achieved FLOP/s and (estimated) algorithmic HBM bytes/s of one training step against the
    pool's measured peaks.  Byte model for the fused execution (bf16): per convolution output
    tensor, forward = write c, read c, write y, read y (4 passes); backward = BN reduce reads
    dy + c, BN apply reads dy + c and writes dc, dgrad reads dc and writes dx, wgrad reads dc + x
    (~9 passes, 1-bit masks ignored) => 13 passes; plus 18 bytes per parameter for the optimizer
    (bf16 grad, fp32 master + momentum read and written, bf16 weight written).  An estimate: L2
    hits make the real DRAM traffic smaller, re-reads make it larger."""
    c = model_cost(model, batch)
    peaks = measured_peaks()
    tflops = c["step_flops"] / (ms_per_step * 1e-3) / 1e12
    gbytes = (c["conv_output_elements"] * 2.0 * 13.0 + c["parameters"] * 18.0) / 1e9
    gbs = gbytes / (ms_per_step * 1e-3)
    tag = "of measured" if peaks["source"] == "measured" else "of fallback"
    return {"step_gflop": round(c["step_flops"] / 1e9, 1), "achieved_tflops": round(tflops, 1),
            "tensor_frac": "%.3f %s" % (tflops / peaks["bf16_tflops"], tag),
            "algorithmic_gbytes_est": round(gbytes, 2), "achieved_gbs_est": round(gbs, 1),
            "hbm_frac": "%.3f %s" % (gbs / peaks["hbm_gbs"], tag),
            "peaks": {k: peaks[k] for k in ("hbm_gbs", "bf16_tflops", "source")}}
