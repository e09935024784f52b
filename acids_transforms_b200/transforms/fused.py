"""Fused stages that `ComposeAudioTransform` substitutes for adjacent children at construction.

(STFT | DGT) + Magnitude  -> one kernel, wave -> normalised (log-)mel, the spectrum never reaches HBM;
(STFT | DGT) + Polar*     -> spectrum written once, both representation kernels fill the stacked output.

A fused stage holds the very same child modules as the chain (so buffers, `scale_data` results and
`state_dict` entries are shared); results are identical to running the children one after the other.
"""
import os
from typing import Optional

import torch

from .base import AudioTransform, frame_times
from .spectral_repr import Magnitude, Polar, PolarIF, _contrast_id
from .stft import STFT
from .dgt import DGT
from .raw import MidSide
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
        """Magnitude.scale_data(STFT(x)) (base.py:144-148, spectral_repr.py:242-245) as ONE pass of the forward kernel in
        its statistics mode: min / max / sum / sum of squares of contrast(|X|) per CTA, one tiny merge, no spectrum in HBM
        and no host synchronisation."""
        if self.stft.track_phase:                      # the reference leaves phase_buffer behind: materialise
            self.mag.scale_data(self.stft(x))
            return
        st = torch.ops.acids_b200.stft_stats(x, self.stft.window, self.stft._n_fft, self.stft._hop,
                                             _contrast_id(self.mag.contrast_mode), self.mag._eps)
        self.mag.norm.set_stats(st)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        return self.stft.invert(self.mag.invert(x, inversion_mode), inversion_mode)


class FusedSTFTPolar(AudioTransform):
    """[MidSide +] (STFT | DGT) + (Polar | PolarIF): one kernel, wave -> stacked [..., T, 2, F'] (raw.py:145-162,
    stft.py:101-102, spectral_repr.py:431-440) when the phase half needs no scan over the frames (raw phase or the
    forward-difference IF); any other configuration runs the children in turn."""
    scriptable = True
    invertible = True
    needs_scaling = True

    def __init__(self, stft: STFT, polar, midside: Optional[MidSide] = None, one_kernel: Optional[bool] = None):
        super().__init__(sr=stft.sr)
        self.stft = stft
        self.polar = polar
        self.has_midside = midside is not None
        self.midside = midside if midside is not None else MidSide(sr=stft.sr)
        # Measured on B200 (tools/polar_debug.py): the one-kernel path is issue bound on the arctangents
        # (25 instructions per bin on the FFT's own warps) — 1.37 ms against 1.32 ms for the three-kernel chain at n_fft 1024
        # (512 clips) and 2.8 ms against 1.5 ms at n_fft 4096, where a CTA holds one frame and the banded projection cannot
        # amortise a column's coefficients over several rows.  It therefore runs only on request (`one_kernel=True`, or
        # ACIDS_B200_FUSE_POLAR=1 at construction); the default is the chain, whose results it reproduces.
        if one_kernel is None:
            one_kernel = os.environ.get("ACIDS_B200_FUSE_POLAR", "0") == "1"
        self.one_kernel = bool(one_kernel)

    def __repr__(self):
        return "Fused(%s%r -> %r)" % ("%r -> " % self.midside if self.has_midside else "", self.stft, self.polar)

    def _prologue(self, x: torch.Tensor) -> torch.Tensor:
        if self.has_midside:
            x = self.midside(x)
        return x

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        stack = self.polar.stack
        mode = self.polar._phase_mode()
        method = self.polar._phase_method()
        fusable = self.one_kernel and (mode == 0 or (mode == 2 and method == 0)) and stack is not None and stack == -2 and not self.stft.track_phase
        stereo = x.ndim >= 2 and x.size(-2) == 2
        if self.has_midside and (self.midside.normalize or not stereo):
            fusable = False
        if not fusable:
            # the chain: spectrum once, then the representation kernels.  MidSide (un-normalised, on stereo input) is folded
            # into the STFT's sample loads, so its waveform is never written (raw.py:145-161 + stft.py:101-102 in one kernel)
            if self.has_midside and stereo and not self.midside.normalize and not self.stft.track_phase:
                X = torch.ops.acids_b200.midside_stft_fwd(x, self.stft.window, self.stft._n_fft, self.stft._hop,
                                                          2 if self.midside.pad_mid else 1)
                return self.polar(X)
            return self.polar(self.stft(self._prologue(x)))
        ms = 0
        if self.has_midside:
            ms = 2 if self.midside.pad_mid else 1
        mag = self.polar.magnitude
        return torch.ops.acids_b200.stft_polar_fwd(x, self.stft.window, self.stft._n_fft, self.stft._hop, mag.band_meta(),
                                                   mag.band_coef(), _contrast_id(mag.contrast_mode), mag._eps,
                                                   mag.norm.get_offset(), mag.norm.get_scale(), mode, method,
                                                   self.polar._phase_weighted(), self.polar.phase.norm.get_offset(),
                                                   self.polar.phase.norm.get_scale(), not self.polar.keep_nyquist, ms)

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-3), self.stft._hop, self.sr, time)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.polar.scale_data(self.stft(self._prologue(x)))

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        y = self.stft.invert(self.polar.invert(x, inversion_mode), inversion_mode)
        if self.has_midside:
            y = self.midside.invert(y, inversion_mode)
        return y


def _bank_fits(mag: Magnitude, stft: STFT) -> bool:
    return (not mag.mel) or mag.mel_bank.shape[-2] == stft._n_fft // 2 + 1


def build_plan(transforms):
    """The execution plan of a chain: children in order, with fusable neighbours replaced by a fused stage."""
    plan, i = [], 0
    items = list(transforms)
    while i < len(items):
        t = items[i]
        nxt = items[i + 1] if i + 1 < len(items) else None
        nxt2 = items[i + 2] if i + 2 < len(items) else None
        if type(t) in (STFT, DGT) and type(nxt) is Magnitude and _bank_fits(nxt, t):
            plan.append(FusedSTFTMagnitude(t, nxt))
            i += 2
        elif type(t) in (STFT, DGT) and type(nxt) in (Polar, PolarIF) and _bank_fits(nxt.magnitude, t):
            plan.append(FusedSTFTPolar(t, nxt))
            i += 2
        elif type(t) is MidSide and type(nxt) in (STFT, DGT) and type(nxt2) in (Polar, PolarIF) and _bank_fits(nxt2.magnitude, nxt):
            plan.append(FusedSTFTPolar(nxt, nxt2, t))
            i += 3
        else:
            plan.append(t)
            i += 1
    return plan
