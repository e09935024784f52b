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
    """`depth` bounds how many micro-batches the host may run ahead of the device->host copies.  Without the bound
    the host enqueues every micro-batch of every call at once, the caching allocator cannot recycle the per-chunk
    device tensors (they are still referenced by queued work) and falls back to `cudaMalloc`, which synchronises the
    device: steps that take 15 ms in steady state then take 20-55 ms at random."""

    def __init__(self, transform, chunk_clips: int = 128, device: Optional[torch.device] = None, depth: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device; acids_transforms_b200 has no CPU fallback")
        self.transform = transform
        self.chunk = int(chunk_clips)
        self.depth = max(1, int(depth))
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._staged = [None, None]      # persistent device staging buffers for the inputs (double buffered)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staging(self, slot: int, like: torch.Tensor) -> torch.Tensor:
        buf = self._staged[slot]
        shape = (self.chunk,) + tuple(like.shape[1:])
        if buf is None or tuple(buf.shape) != shape or buf.dtype != like.dtype:
            buf = torch.empty(shape, dtype=like.dtype, device=self.device)
            self._staged[slot] = buf
        return buf

    def __call__(self, x_host: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_host [B, ...] (ideally pinned) -> pinned host tensor with the transform's output for every clip."""
        B = x_host.shape[0]
        n = (B + self.chunk - 1) // self.chunk
        free_in = [None, None]          # event: compute finished reading staging slot i
        copied = []                     # event per micro-batch: its result is in host memory
        self.h2d_bytes = self.d2h_bytes = 0
        caller = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(caller)
        for i in range(n):
            lo, hi = i * self.chunk, min(B, (i + 1) * self.chunk)
            slot = i & 1
            if i >= self.depth:
                copied[i - self.depth].synchronize()        # bound the host's run-ahead (see class docstring)
            with torch.cuda.stream(self.s_in):
                if free_in[slot] is not None:
                    self.s_in.wait_event(free_in[slot])
                src = x_host[lo:hi]
                staged = self._staging(slot, src)[:hi - lo]
                staged.copy_(src, non_blocking=True)
                self.h2d_bytes += src.numel() * src.element_size()
                ready = self.s_in.record_event()
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ready)
                y = self.transform(staged)
                free_in[slot] = self.s_run.record_event()
                done = free_in[slot]
            if out is None:
                out = torch.empty((B,) + tuple(y.shape[1:]), dtype=y.dtype, pin_memory=True)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                out[lo:hi].copy_(y, non_blocking=True)
                y.record_stream(self.s_out)
                self.d2h_bytes += y.numel() * y.element_size()
                copied.append(self.s_out.record_event())
        caller.wait_stream(self.s_out)
        caller.wait_stream(self.s_run)
        return out
