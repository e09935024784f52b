"""Fused stages that `ComposeAudioTransform` substitutes for adjacent children at construction.

(STFT | DGT) + Magnitude  -> one kernel, wave -> normalised (log-)mel, the spectrum never reaches HBM;
(STFT | DGT) + Polar*     -> spectrum written once, both representation kernels fill the stacked output.

A fused stage holds the very same child modules as the chain (so buffers, `scale_data` results and
`state_dict` entries are shared); results are identical to running the children one after the other.
"""
from typing import Optional

import torch

from .base import AudioTransform, frame_times
from .spectral_repr import Magnitude, Polar, PolarIF, _contrast_id
from .stft import STFT
from .dgt import DGT
from .. import _torch_ops  # noqa: F401


class FusedSTFTMagnitude(AudioTransform):
    scriptable = True
    invertible = True
    needs_scaling = True

    def __init__(self, stft: STFT, mag: Magnitude):
        super().__init__(sr=stft.sr)
        self.stft = stft
        self.mag = mag

    def __repr__(self):
        return "Fused(%r -> %r)" % (self.stft, self.mag)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.stft.track_phase:                      # phase_buffer wanted: the spectrum must be materialised
            return self.mag(self.stft(x))
        return torch.ops.acids_b200.stft_mag_fwd(x, self.stft.window, self.stft._n_fft, self.stft._hop,
                                                 self.mag.band_meta(), self.mag.band_coef(),
                                                 _contrast_id(self.mag.contrast_mode), self.mag._eps,
                                                 self.mag.norm.get_offset(), self.mag.norm.get_scale(),
                                                 not self.mag.keep_nyquist)

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-2), self.stft._hop, self.sr, time)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.mag.scale_data(self.stft(x))

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        return self.stft.invert(self.mag.invert(x, inversion_mode), inversion_mode)


class FusedSTFTPolar(AudioTransform):
    scriptable = True
    invertible = True
    needs_scaling = True

    def __init__(self, stft: STFT, polar):
        super().__init__(sr=stft.sr)
        self.stft = stft
        self.polar = polar

    def __repr__(self):
        return "Fused(%r -> %r)" % (self.stft, self.polar)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.polar(self.stft(x))

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-3), self.stft._hop, self.sr, time)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.polar.scale_data(self.stft(x))

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        return self.stft.invert(self.polar.invert(x, inversion_mode), inversion_mode)


def build_plan(transforms):
    """The execution plan of a chain: children in order, with fusable neighbours replaced by a fused stage."""
    plan, i = [], 0
    items = list(transforms)
    while i < len(items):
        t = items[i]
        nxt = items[i + 1] if i + 1 < len(items) else None
        if type(t) in (STFT, DGT) and type(nxt) is Magnitude and (
                not nxt.mel or nxt.mel_bank.shape[-2] == t._n_fft // 2 + 1):
            plan.append(FusedSTFTMagnitude(t, nxt))
            i += 2
        else:
            plan.append(t)
            i += 1
    return plan
