"""One streaming step of a stateful chain as a CUDA graph.

The reference's purpose is block-by-block processing (`README.md:4`): `OverlapAdd.forward -> RealtimeSTFT / RealtimeDGT
.forward -> ... -> invert -> OverlapAdd.invert` on a few hundred samples per call (`oadd.py:70-104`, `stft.py:248-266`,
`dgt.py:284-302`).  At that size every kernel runs for a few microseconds and the step is bound by launch and
dispatcher overhead, not by the GPU.  `GraphedStep` captures the whole step once and replays it with a single
`cudaGraphLaunch`.

The stateful modules replace their carry tensors on every call (`self.input_buffer = ...`), which a replay cannot
follow: it would keep reading the tensor that was current at capture time.  The capture therefore ends with
`old.copy_(new)` for every tensor attribute the step replaced and puts `old` back, so the state lives at fixed
addresses and advances inside the graph.  Limits: the step must not synchronise with the host (PGHI inversion does),
must not draw host-side random numbers, and its state shapes must not change after the warm-up calls.
"""
from typing import Callable, Dict, Tuple

import torch


def _tensor_attrs(root: torch.nn.Module) -> Dict[Tuple[torch.nn.Module, str, bool], torch.Tensor]:
    """Every tensor a module of the tree holds as a buffer or as a plain attribute (the realtime transforms use both)."""
    found = {}
    for m in root.modules():
        for name, buf in m._buffers.items():
            if isinstance(buf, torch.Tensor):
                found[(m, name, True)] = buf
        for name, val in vars(m).items():
            if isinstance(val, torch.Tensor) and not isinstance(val, torch.nn.Parameter):
                found[(m, name, False)] = val
    return found


def _assign(key, value: torch.Tensor) -> None:
    m, name, is_buffer = key
    if is_buffer:
        m._buffers[name] = value
    else:
        object.__setattr__(m, name, value)


class GraphedStep:
    """step(x) -> y for a fixed input shape, replayed from a CUDA graph.

    transform: the module tree whose state `step` advances (usually the chain itself).
    step:      callable(Tensor) -> Tensor built from the chain, e.g. `lambda x: chain.invert(chain(x))`.
    example:   a CUDA tensor with the shape / dtype of every later input.
    The `warmup` eager calls settle the state shapes (the carry buffers start as `[keep]` and take the batch shape on
    first use); the state is put back afterwards (`reset`), i.e. the first replay starts where the module stood.
    The returned tensor is the graph's static output: consume or copy it before the next call.
    """

    def __init__(self, transform: torch.nn.Module, step: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor,
                 warmup: int = 3):
        if not example.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors; acids_transforms_b200 has no CPU fallback")
        self.transform = transform
        self.step = step
        self.static_in = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        initial = _tensor_attrs(transform)
        self._initial = {key: val.clone() for key, val in initial.items()}
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                step(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        settled = _tensor_attrs(transform)
        # the state is whatever the warm-up calls replaced
        self._state = [key for key, val in settled.items() if initial.get(key) is not val]
        self.reset()
        before = {key: settled[key] for key in self._state}
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = step(self.static_in)
            after = _tensor_attrs(transform)
            for key, old in before.items():
                new = after[key]
                if new is old:
                    continue
                if new.shape != old.shape or new.dtype != old.dtype:
                    raise RuntimeError("GraphedStep: state '%s' changed shape during capture (%s -> %s)"
                                       % (key[1], tuple(old.shape), tuple(new.shape)))
                old.copy_(new)
                _assign(key, old)
            extra = [key[1] for key, val in after.items() if key not in before and settled.get(key) is not val]
            if extra:
                raise RuntimeError("GraphedStep: state %s appeared during capture; raise `warmup`" % extra)

    def reset(self) -> None:
        """Back to the state of the module as it was handed over: its own values where the shape is unchanged (e.g. the
        random initial phase of the sinebank inversion), zeros for the carry buffers that took the batch shape."""
        cur = _tensor_attrs(self.transform)
        for key in self._state:
            first = self._initial.get(key)
            if first is not None and first.shape == cur[key].shape:
                cur[key].copy_(first)
            else:
                cur[key].zero_()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise RuntimeError("GraphedStep was captured for %s %s, got %s %s"
                               % (tuple(self.static_in.shape), self.static_in.dtype, tuple(x.shape), x.dtype))
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
