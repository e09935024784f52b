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


class StreamStep:
    """The block-by-block step of an (OverlapAdd, RealtimeSTFT | RealtimeDGT) pair as ONE kernel launch per direction —
    or one launch for the whole round trip — instead of the 7-9 kernels of the eager modules (csrc/stream.cu).

        step = StreamStep(oadd, rt, batch_shape=(16,), device="cuda")
        X = step.analysis(block)        # OverlapAdd.forward -> rt.forward          [..., n hop] -> complex [..., n, F]
        y = step.synthesis(X)           # rt.invert -> OverlapAdd.invert            complex [..., n, F] -> [..., n hop]
        y = step.roundtrip(block)       # both, the spectrum never leaves the registers

    The carried state (`OverlapAdd.input_buffer` / `.output_buffer`, oadd.py:70-104) lives in two tensors of this object
    at fixed addresses and is advanced in place by the kernels, so a call is capturable in a CUDA graph as is
    (`graph=True` replays the round trip from one: the launch is then a single cudaGraphLaunch without any per-call
    Python argument marshalling).  The synthesis half reproduces the eager kernels bit for bit, the analysis half to rounding
    (<= 2e-6 of the peak: ptxas fuses the window into the first butterfly level differently in the two kernels).  `pull()` copies
    the modules' current state in, `push()` hands the state back to the modules (so eager calls can continue the stream).
    """

    def __init__(self, oadd, rt, batch_shape=(), device="cuda", graph: bool = False, block: int = 0):
        from . import ops
        self._ops = ops
        self.oadd, self.rt = oadd, rt
        self.n_fft, self.hop = int(rt._n_fft), int(oadd._hop)
        if int(oadd._n_fft) != self.n_fft:
            raise RuntimeError("StreamStep: OverlapAdd(n_fft=%d) feeds a transform of n_fft=%d" % (oadd._n_fft, self.n_fft))
        if self.n_fft % self.hop:
            raise RuntimeError("StreamStep: hop=%d must divide n_fft=%d" % (self.hop, self.n_fft))
        self.keep = self.n_fft - self.hop
        self.batch_shape = tuple(int(b) for b in batch_shape)
        dev = torch.device(device)
        self.tail = torch.zeros(self.batch_shape + (self.keep,), dtype=torch.float32, device=dev)
        self.carry = torch.zeros(self.batch_shape + (self.keep,), dtype=torch.float32, device=dev)
        self.window = rt.window.detach()[:self.n_fft].to(dev, torch.float32).contiguous()
        self.inv_window = rt.inv_window.detach()[:self.n_fft].to(dev, torch.float32).contiguous()
        self.gain = float(oadd._gain)
        self._graph = None
        if graph:
            if block <= 0 or block % self.hop:
                raise RuntimeError("StreamStep(graph=True) needs the block length (a multiple of hop)")
            self._static_in = torch.zeros(self.batch_shape + (block,), dtype=torch.float32, device=dev)
            self._static_out = torch.empty_like(self._static_in)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._roundtrip(self._static_in, self._static_out)       # opt-in of the kernel's shared memory happens here
            torch.cuda.current_stream(dev).wait_stream(side)
            self.reset()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._roundtrip(self._static_in, self._static_out)

    def reset(self) -> None:
        self.tail.zero_()
        self.carry.zero_()

    def pull(self) -> None:
        """Take over the stream where the eager modules stand."""
        if self.oadd.input_buffer.shape == self.tail.shape:
            self.tail.copy_(self.oadd.input_buffer)
        if self.oadd.output_buffer.shape == self.carry.shape:
            self.carry.copy_(self.oadd.output_buffer)

    def push(self) -> None:
        """Hand the stream back to the eager modules."""
        self.oadd.input_buffer = self.tail.clone()
        self.oadd.output_buffer = self.carry.clone()

    def analysis(self, x: torch.Tensor) -> torch.Tensor:
        return self._ops.stream_analysis(x, self.window, self.n_fft, self.hop, self.tail)

    def synthesis(self, X: torch.Tensor) -> torch.Tensor:
        return self._ops.stream_synthesis(X, self.inv_window, self.n_fft, self.hop, self.gain, self.carry)

    def _roundtrip(self, x: torch.Tensor, out=None) -> torch.Tensor:
        return self._ops.stream_roundtrip(x, self.window, self.inv_window, self.n_fft, self.hop, self.gain, self.tail, self.carry, out=out)

    def roundtrip(self, x: torch.Tensor) -> torch.Tensor:
        if self._graph is not None:
            if x.shape != self._static_in.shape:
                raise RuntimeError("StreamStep was captured for %s, got %s" % (tuple(self._static_in.shape), tuple(x.shape)))
            self._static_in.copy_(x, non_blocking=True)
            self._graph.replay()
            return self._static_out
        return self._roundtrip(x)
