"""Host-resident batches through a transform with copy / compute overlap.

A chain applied to a CPU tensor works (the ops stage it to the GPU and back) but serialises
host->device copy, kernels and device->host copy.  `HostPipeline` splits the batch of clips into
micro-batches and runs three CUDA streams so that the PCIe transfers of neighbouring micro-batches hide
behind each other and behind the kernels; the result lands in pinned host memory.  This is the path
`bench.py` times as `e2e`.
"""
from typing import Optional

import torch


class HostPipeline:
    def __init__(self, transform, chunk_clips: int = 128, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device; acids_transforms_b200 has no CPU fallback")
        self.transform = transform
        self.chunk = int(chunk_clips)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, x_host: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_host [B, ...] (ideally pinned) -> pinned host tensor with the transform's output for every clip."""
        B = x_host.shape[0]
        n = (B + self.chunk - 1) // self.chunk
        staged = [None, None]
        free_in = [None, None]          # event: compute finished reading staged[i]
        outs_pending = []
        self.h2d_bytes = self.d2h_bytes = 0
        caller = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(caller)
        for i in range(n):
            lo, hi = i * self.chunk, min(B, (i + 1) * self.chunk)
            slot = i & 1
            with torch.cuda.stream(self.s_in):
                if free_in[slot] is not None:
                    self.s_in.wait_event(free_in[slot])
                src = x_host[lo:hi]
                staged[slot] = src.to(self.device, non_blocking=True)
                self.h2d_bytes += src.numel() * src.element_size()
                ready = self.s_in.record_event()
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ready)
                y = self.transform(staged[slot])
                free_in[slot] = self.s_run.record_event()
                done = self.s_run.record_event()
            if out is None:
                out = torch.empty((B,) + tuple(y.shape[1:]), dtype=y.dtype, pin_memory=True)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                out[lo:hi].copy_(y, non_blocking=True)
                y.record_stream(self.s_out)
                self.d2h_bytes += y.numel() * y.element_size()
            outs_pending.append(y)
        caller.wait_stream(self.s_out)
        caller.wait_stream(self.s_run)
        return out
