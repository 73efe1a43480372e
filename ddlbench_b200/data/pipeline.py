"""Host -> device input pipeline for the end-to-end measurement.

Every step the batch starts in PINNED HOST memory as raw uint8 HWC pixels (1 byte/value; the
reference ships normalised fp32 NCHW, 4 bytes/value, with a synchronous ``data.to(device)``
inside the step — ``benchmark/imagenet/imagenet_pytorch.py:152``).  The copy of batch i+1 is
issued on a dedicated copy stream into the other half of a double buffer while batch i
trains; a fused kernel converts uint8 -> normalised bf16 channels-last straight into the
captured graph's static input.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .synthetic import SyntheticImages


class DevicePrefetcher:
    def __init__(self, source: SyntheticImages, device, depth: int = 2):
        self.src = source
        self.device = torch.device(device)
        self.depth = depth
        self.on_gpu = self.device.type == "cuda"
        img, lab = source.batch(0)
        self.stage_x = [torch.empty(img.shape, dtype=torch.uint8, device=self.device)
                        for _ in range(depth)]
        self.stage_t = [torch.empty(lab.shape, dtype=torch.int64, device=self.device)
                        for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.on_gpu else None
        self.ready = [None] * depth      # event: copy into slot finished
        self.consumed = [None] * depth   # event: consumer finished reading slot
        self.i_issue = 0
        self.h2d_bytes = source.bytes_per_step()

    def issue(self) -> None:
        """Start the H2D copy of the next batch into the next slot."""
        k = self.i_issue % self.depth
        img, lab = self.src.batch(self.i_issue)
        if self.on_gpu:
            with torch.cuda.stream(self.copy_stream):
                if self.consumed[k] is not None:
                    self.copy_stream.wait_event(self.consumed[k])
                self.stage_x[k].copy_(img, non_blocking=True)
                self.stage_t[k].copy_(lab, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                self.ready[k] = ev
        else:
            self.stage_x[k].copy_(img)
            self.stage_t[k].copy_(lab)
        self.i_issue += 1

    def get(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device tensors of batch ``i`` (the compute stream waits for its copy)."""
        k = i % self.depth
        if self.on_gpu and self.ready[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        return self.stage_x[k], self.stage_t[k]

    def release(self, i: int) -> None:
        """Mark slot of batch ``i`` as consumed (recorded on the compute stream)."""
        if self.on_gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.consumed[i % self.depth] = ev
