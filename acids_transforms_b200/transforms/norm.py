"""Normalize — global offset/scale (acids_transforms/transforms/norm.py:13-99).

`scale_data` is one fused reduction on the GPU (min, max, mean, unbiased std in a single pass,
`acids_stats`); the statistics stay on the device, so fitting never synchronises the host.  The
spectral representations read `offset` / `scale` straight from these buffers inside their kernels.
"""
from typing import Optional

import torch

from .base import AudioTransform
from .. import _torch_ops  # noqa: F401  (registers torch.ops.acids_b200)

__all__ = ["Normalize"]


def fit_offset_scale(mode: str, st: torch.Tensor):
    """(offset, scale) from the float64 [min, max, mean, std] statistics, in float32 like norm.py:26-38."""
    mn, mx = st[0].to(torch.float32), st[1].to(torch.float32)
    if mode == "unipolar":
        return mn, mx - mn
    if mode == "bipolar":
        off = (mx + mn) / 2
        return off, mx - off
    return st[2].to(torch.float32), st[3].to(torch.float32)


class Normalize(AudioTransform):
    scriptable = True

    def __repr__(self):
        return "Normalize(mode=%s)" % self.mode

    def __init__(self, mode: Optional[str] = "gaussian"):
        super().__init__()
        self.mode = mode
        self.needs_scaling = True
        self.register_buffer("offset", torch.zeros(0))
        self.register_buffer("scale", torch.ones(1))

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        if self.mode == "unipolar" or self.mode == "bipolar" or self.mode == "gaussian":
            self.set_stats(torch.ops.acids_b200.stats(x, 0, 0.0))
        self.needs_scaling = False

    @torch.jit.export
    def set_stats(self, st: torch.Tensor) -> None:
        """Fit from precomputed [min, max, mean, std] (float64[4]) — used by the fused scale_data paths."""
        mn, mx = st[0].to(torch.float32), st[1].to(torch.float32)
        if self.mode == "unipolar":
            self.offset = mn
            self.scale = mx - mn
        elif self.mode == "bipolar":
            off = (mx + mn) / 2
            self.offset = off
            self.scale = mx - off
        elif self.mode == "gaussian":
            self.offset = st[2].to(torch.float32)
            self.scale = st[3].to(torch.float32)
        self.needs_scaling = False

    @torch.jit.export
    def get_offset(self) -> Optional[torch.Tensor]:
        return self.offset

    @torch.jit.export
    def get_scale(self) -> Optional[torch.Tensor]:
        return self.scale

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return (x - self.offset) / self.scale

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        return x * self.scale + self.offset

    def get_normalization_modes(self):
        return ["unipolar", "bipolar", "gaussian"]

    # ---- reference test hooks (norm.py:49-99) ----
    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        from ..utils.misc import frame
        x = frame(x, min(256, x.shape[-1]), min(64, x.shape[-1]), -1)
        tol = torch.finfo(x.dtype).eps
        x_norm = x
        for mode in self.get_normalization_modes():
            self.mode = mode
            self.scale_data(x)
            x_norm = self(x)
            if mode == "unipolar":
                assert x_norm.min() == 0. and x_norm.max() == 1.
            elif mode == "bipolar":
                assert abs(float(x_norm.min()) + 1.) < 1e-6 and x_norm.max() == 1.
            else:
                assert (x_norm.mean().abs() < 16 * tol).item()
                assert ((x_norm.std() - 1).pow(2) < tol).item()
        return x_norm if time is None else (x_norm, time)

    def test_inversion(self, x: torch.Tensor, tolerance: Optional[float] = None):
        from ..utils.misc import frame
        x = frame(x, min(256, x.shape[-1]), min(64, x.shape[-1]), -1)
        tol = torch.finfo(x.dtype).eps if tolerance is None else tolerance
        for mode in self.get_normalization_modes():
            self.mode = mode
            self.scale_data(x)
            x_den = self.invert(self(x))
            assert ((x.min() - x_den.min()).pow(2) < tol).item()
            assert ((x.max() - x_den.max()).pow(2) < tol).item()
        return {}

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True):
        x = torch.rand((5, 256))
        transform.scale_data(x)
        y = transform(x)
        if invert:
            transform.invert(y)
