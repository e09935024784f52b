"""OverlapAdd — streaming framing / overlap-add with carry buffers (acids_transforms/transforms/oadd.py).

`forward` is metadata only (prepend the saved tail, strided frame view).  `invert` is one gather
kernel (`acids_ola_stream`) that adds the carried tail and every frame touching an output sample in a
fixed order — the reference loops over frames in Python with slice `+=` (oadd.py:100-101) — and writes
the next carry.  Buffers follow the data's device (the reference allocates them on the CPU, oadd.py:38).
"""
from typing import Optional

import torch

from .base import AudioTransform, frame_times
from ..utils.misc import frame
from .. import _torch_ops  # noqa: F401

__all__ = ["OverlapAdd"]


def _oadd_after_load(module, incompatible_keys):
    """A loaded checkpoint may carry another n_fft / hop_length / gain: refresh their Python mirrors."""
    module._n_fft = int(module.n_fft.item())
    module._hop = int(module.hop_length.item())
    module.frames_out = module._n_fft // module._hop - 1
    module.keep = module.frames_out * module._hop
    module._gain = float(module.gain_compensation)


class OverlapAdd(AudioTransform):
    @property
    def invertible(self):
        return True

    @property
    def scriptable(self):
        return True

    @property
    def needs_scaling(self):
        return False

    def __repr__(self):
        return "OverlapAdd(n_fft=%s, hop_length=%s)" % (self._n_fft, self._hop)

    def __init__(self, n_fft: int = 1024, hop_length: int = 128, dim: int = -1) -> None:
        super().__init__()
        self._n_fft = int(n_fft)
        self._hop = int(hop_length)
        self.register_buffer("n_fft", torch.tensor(n_fft))
        self.register_buffer("hop_length", torch.tensor(hop_length))
        self.frames_out = self._n_fft // self._hop - 1                      # oadd.py:27
        self.keep = self.frames_out * self._hop
        self.register_buffer("input_buffer", torch.zeros(self.keep))
        self.register_buffer("output_buffer", torch.zeros(self.keep))
        # oadd.py:31: the peak of a rectangular overlap-add of ones, scaled by 2/overlap — 2.0 for every integer N/H
        ones = torch.ones(12, (self.frames_out + 1) * self._n_fft)
        fr = frame(ones, self._n_fft, self._hop, -1)
        rec = torch.zeros(12, fr.size(-2) * self._hop + self._n_fft)
        for i in range(fr.size(-2)):
            rec[..., i * self._hop:i * self._hop + self._n_fft] += fr[..., i, :] / (int(self._n_fft / self._hop) / 2)
        self.register_buffer("gain_compensation", rec.max())
        self._gain = float(rec.max())
        self.register_load_state_dict_post_hook(_oadd_after_load)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[..., C] -> [..., C/hop, n_fft] strided view over (saved tail ++ x)  (oadd.py:70-74)."""
        if self.input_buffer.shape[:-1] != x.shape[:-1] or self.input_buffer.device != x.device:
            tail = torch.zeros(list(x.shape[:-1]) + [self.keep], dtype=x.dtype, device=x.device)
        else:
            tail = self.input_buffer
        self.input_buffer = x[..., x.size(-1) - self.keep:].clone() if self.keep > 0 else x[..., :0]
        return frame(torch.cat([tail, x], dim=-1), self._n_fft, self._hop, -1)

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-2), self._hop, self.sr, time)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: Optional[float] = None) -> torch.Tensor:
        """[..., n, n_fft] -> [..., (n-1) hop + n_fft - keep], carrying `keep` samples to the next call (oadd.py:91-104)."""
        carry: Optional[torch.Tensor] = None
        if self.output_buffer.shape[:-1] == x.shape[:-2] and self.keep > 0:
            carry = self.output_buffer
        out, carry_out = torch.ops.acids_b200.ola_stream(x, self._hop, self.keep, carry, self._gain)
        self.output_buffer = carry_out
        return out
