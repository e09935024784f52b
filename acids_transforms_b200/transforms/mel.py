"""MFCC — host-side mirror of acids_transforms/transforms/mel.py.

In the reference this class is a torchaudio `MelSpectrogram` (power mel spectrogram, frequency-major
output [..., n_mels, T], not invertible, no log, no DCT — mel.py:38-44, :68-77); that is the default
behaviour here too, computed by ONE kernel: periodic-Hann STFT, |X|^power, banded HTK mel projection
and optional Normalize, without materialising the spectrum.

BASELINE.json's MFCC configuration asks for "40 coefficients", which the reference class cannot
produce; the opt-in `n_mfcc=` follows torchaudio.transforms.MFCC (dB with top_db=80, ortho DCT-II,
_transforms.py:701-718) and yields [..., n_mfcc, T].
"""
import math
from typing import Optional

import torch

from .base import AudioTransform, NotInvertibleError, frame_times
from .norm import Normalize
from .spectral_repr import Dummy, melscale_fbanks
from .. import ops as _ops
from .. import _torch_ops  # noqa: F401

__all__ = ["MFCC"]


def create_dct(n_mfcc: int, n_mels: int) -> torch.Tensor:
    """Ortho DCT-II matrix [n_mels, n_mfcc] (torchaudio functional.py:636-667)."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()


def _mfcc_reference_keys(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """A reference checkpoint holds the torchaudio sub-module's buffers (`transform.spectrogram.window`,
    `transform.mel_scale.fb`, mel.py:43-44).  Both are functions of (n_fft, n_mels, sr), which this module re-derives
    itself: check that they describe the same transform, then consume the keys so that strict loading succeeds."""
    w = state_dict.pop(prefix + "transform.spectrogram.window", None)
    fb = state_dict.pop(prefix + "transform.mel_scale.fb", None)
    if w is not None and w.numel() != module.n_fft:
        error_msgs.append("MFCC: checkpoint window has %d taps, this module n_fft=%d" % (w.numel(), module.n_fft))
    if fb is not None and tuple(fb.shape) != (module.n_fft // 2 + 1, module.n_mels):
        error_msgs.append("MFCC: checkpoint mel bank is %s, this module needs (%d, %d)" % (tuple(fb.shape), module.n_fft // 2 + 1, module.n_mels))


class MFCC(AudioTransform):
    @property
    def invertible(self):
        return False

    @property
    def scriptable(self):
        return True

    @property
    def needs_scaling(self):
        return self.has_norm

    def __repr__(self):
        s = "MFCC(n_fft=%s, hop_length=%spower=%s, n_mels=%s" % (self.n_fft, self.hop_length, self.power, self.n_mels)
        if self.has_norm:
            s += ", f%s" % self.norm
        return s + ")"

    def __init__(self, n_fft: int = 1024, hop_length: int = 256, power: float = 2., n_mels: int = 128, sr: int = 44100,
                 norm_mode: Optional[str] = None, n_mfcc: Optional[int] = None, top_db: float = 80.0):
        super().__init__(sr=sr)
        self.has_norm = norm_mode is not None
        self.norm = Normalize(mode=norm_mode) if norm_mode is not None else Dummy()
        self.n_mfcc = 0 if n_mfcc is None else int(n_mfcc)
        self.top_db = float(top_db)
        self.n_fft = 0
        self.hop_length = 0
        self.power = 2.0
        self.n_mels = 0
        self.register_buffer("window", torch.zeros(0), persistent=False)
        self.register_buffer("mel_meta", torch.zeros(0, dtype=torch.int32), persistent=False)
        self.register_buffer("mel_coef", torch.zeros(0), persistent=False)
        self.register_buffer("dct_mat", torch.zeros(0, 0), persistent=False)
        self.set_transform(n_fft, n_mels, hop_length, power)
        self._register_load_state_dict_pre_hook(_mfcc_reference_keys, with_module=True)

    @torch.jit.unused
    def set_transform(self, n_fft: int, n_mels: int, hop_length: int, power: float) -> None:
        """mel.py:38-44: MelSpectrogram(sr, n_fft, hop_length=hop, n_mels=n_mels, power=power) — periodic Hann,
        f_min 0, f_max sr/2, HTK scale, no filter normalisation."""
        self.n_fft, self.hop_length, self.power, self.n_mels = int(n_fft), int(hop_length), float(power), int(n_mels)
        dev = self.window.device
        self.window = torch.hann_window(self.n_fft).to(dev)
        fb = melscale_fbanks(self.n_fft // 2 + 1, 0.0, float(self.sr // 2), self.n_mels, self.sr)
        meta, coef = _ops.BandedMatrix(fb).tensors()
        self.mel_meta, self.mel_coef = meta.to(dev), coef.to(dev)
        self.dct_mat = create_dct(self.n_mfcc, self.n_mels).to(dev) if self.n_mfcc > 0 else torch.zeros(0, 0, device=dev)

    @torch.jit.unused
    def dense_bank(self) -> torch.Tensor:
        """The [n_fft/2+1, n_mels] HTK filterbank torchaudio's MelSpectrogram would hold (mel.py:43-44)."""
        return melscale_fbanks(self.n_fft // 2 + 1, 0.0, float(self.sr // 2), self.n_mels, self.sr)

    @torch.jit.export
    def mel_power(self, x: torch.Tensor) -> torch.Tensor:
        """Un-normalised mel spectrogram [..., n_mels, T]."""
        return torch.ops.acids_b200.melspec_fwd(x, self.window, self.n_fft, self.hop_length, self.mel_meta, self.mel_coef,
                                                self.power, None, None)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.n_mfcc > 0:
            y = torch.ops.acids_b200.mfcc_dct(self.mel_power(x), self.dct_mat, self.top_db)
            if self.has_norm:
                y = self.norm(y)
            return y
        return torch.ops.acids_b200.melspec_fwd(x, self.window, self.n_fft, self.hop_length, self.mel_meta, self.mel_coef,
                                                self.power, self.norm.get_offset(), self.norm.get_scale())

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        # the reference counts "chunks" on dim -2, i.e. the MEL axis of the [.., n_mels, T] output (mel.py:47-57)
        return y, frame_times(y.size(-2), self.hop_length, self.sr, time)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        if self.has_norm:
            self.norm.scale_data(x)          # mel.py:60-62: fitted on whatever it is given

    @property
    def ratio(self):
        return self.hop_length

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        raise NotInvertibleError
