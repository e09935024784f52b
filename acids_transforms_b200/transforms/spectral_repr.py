"""Spectral representations — host-side mirror of acids_transforms/transforms/spectral_repr.py.

Real / Imaginary / Magnitude / Phase / IF and the stacked Cartesian / Polar / PolarIF.  Each
`forward` / `invert` is one kernel launch that fuses what the reference runs as 4-6 eager passes
(abs -> 513x513 matmul -> log -> sub -> div; angle -> unwrap -> diff -> scale -> normalise), reading
the Normalize buffers on the device.  Reference quirks are reproduced by default and listed in
DESIGN.md (keep_nyquist=False drops bin 0; scale_data ignores the mel bank; the inverse mel bank is
a row-normalised transpose; the asymmetric /pi of the IF).  Deliberate deviations:

* `IF(weighted=True)` works on every call (the reference raises IndexError from its second call on,
  spectral_repr.py:339);
* `IF.invert` never modifies its argument (the reference scales it in place when un-normalised,
  spectral_repr.py:361-368).
"""
import math
from typing import Dict, Optional, Tuple, Union

import torch

from .base import AudioTransform
from .norm import Normalize
from .stft import STFT
from ..utils.misc import reshape_batches
from .. import ops as _ops
from .. import _torch_ops  # noqa: F401

__all__ = ["Real", "Imaginary", "Magnitude", "Phase", "IF", "Cartesian", "Polar", "PolarIF"]


class Dummy(AudioTransform):
    """Identity normalisation (spectral_repr.py:17-18)."""
    scriptable = True
    mode: Optional[str] = None

    def __init__(self):
        super().__init__()
        self.mode = None

    @torch.jit.export
    def get_offset(self) -> Optional[torch.Tensor]:
        return None

    @torch.jit.export
    def get_scale(self) -> Optional[torch.Tensor]:
        return None

    @torch.jit.export
    def set_stats(self, st: torch.Tensor) -> None:
        pass


def _contrast_id(mode: Optional[str]) -> int:
    if mode is None or mode == "none":
        return 0
    if mode == "log1p":
        return 1
    if mode == "log":
        return 2
    if mode == "log10":
        return 3
    raise TypeError("unknown contrast type %s" % mode)      # spectral_repr.py:201


def _method_id(method: Optional[str]) -> int:
    if method is None:
        raise AttributeError("method None not known")
    if method == "forward":
        return 0
    if method == "backward":
        return 1
    if method == "central":
        return 2
    raise AttributeError("method %s not known" % method)    # spectral_repr.py:331


class _Representation(AudioTransform):
    def __init__(self, sr: int = 44100, mode: Optional[str] = None, keep_nyquist: bool = True):
        super().__init__(sr=sr)
        if mode is None or mode == "none":
            self.norm = Dummy()
        else:
            self.norm = Normalize(mode)
        self.keep_nyquist = keep_nyquist

    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return True

    @property
    def needs_scaling(self):
        return True

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.norm.scale_data(x)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        # de-normalise (+ the zero bin appended when keep_nyquist=False), spectral_repr.py:46-53
        return torch.ops.acids_b200.phase_inv(x, 0, 0, self.norm.get_offset(), self.norm.get_scale(), not self.keep_nyquist)

    @torch.jit.export
    def invert_polar(self, x: torch.Tensor, mag: torch.Tensor) -> torch.Tensor:
        """mag * exp(i invert(x)) in one kernel: the tail of SpectralRepresentation.invert (spectral_repr.py:447-452)."""
        return torch.ops.acids_b200.phase_inv_polar(x, mag, 0, 0, self.norm.get_offset(), self.norm.get_scale(), not self.keep_nyquist)

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True):
        shape = (2, 10, 513)
        z = torch.randn(*shape) * torch.exp(2j * math.pi * torch.rand(*shape))
        transform.scale_data(z)
        y = transform(z)
        if invert:
            transform.invert(y)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        stft = STFT()
        if time is None:
            X = stft(x)
            self.scale_data(X)
            return self(X)
        X, time = stft.forward_with_time(x, time)
        self.scale_data(X)
        return self.forward_with_time(X, time)


class Real(_Representation):
    """Real part, normalised (spectral_repr.py:78-105): a strided view of the interleaved spectrum plus the
    Normalize elementwise — no kernel of its own."""

    def __repr__(self):
        return "Real(norm=%s)" % self.norm.mode

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.keep_nyquist:
            x = x[..., 1:]
        return self.norm(x.real)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.norm.scale_data(x.real.contiguous())

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT(n_fft=512, hop_length=128)
        X = stft(flat)
        self.scale_data(X)
        rec = torch.complex(self.invert(self(X)), X.imag)
        return {"direct": stft.invert(rec).reshape(batch + [-1])}


class Imaginary(_Representation):
    def __repr__(self):
        return "Imaginary(norm=%s)" % self.norm.mode

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_complex(x):
            y = self.norm(x.imag)
        else:
            y = torch.zeros_like(x)
        if not self.keep_nyquist:
            y = y[..., 1:]
        return y

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.norm.scale_data(x.imag.contiguous())

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        self.scale_data(X)
        rec = torch.complex(X.real, self.invert(self(X)))
        return {"direct": stft.invert(rec).reshape(batch + [-1])}


def build_mel_banks(sr: int, n_fft: int, keep_nyquist: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Square HTK mel banks of Magnitude (spectral_repr.py:173-189): forward = columns / column sums,
    inverse = (rows / row sums)^T.  Same float32 torch operations as torchaudio's melscale_fbanks
    (functional.py:518-588), so the buffers are bit-identical to the reference's on the same host."""
    n_bins = n_fft // 2 + 1
    fft_scale = torch.arange(n_bins) / n_fft * sr
    if not keep_nyquist:
        fft_scale = fft_scale[..., 1:]
    fb = melscale_fbanks(n_bins, float(fft_scale[0]), float(fft_scale[-1]), n_bins, sr)
    col = fb.sum(0)
    fwd = fb / torch.where(col != 0, col, torch.ones_like(col)).unsqueeze(0)
    row = fb.sum(1)
    inv = fb / torch.where(row != 0, row, torch.ones_like(row)).unsqueeze(1)
    return fwd.unsqueeze(0), inv.transpose(-2, -1).unsqueeze(0)


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK triangular filterbank [n_freqs, n_mels], no area normalisation (torchaudio functional.py:518-588)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _magnitude_after_load(module, incompatible_keys):
    """load_state_dict may replace `mel_bank` / `inverse_mel_bank` / `eps`: re-derive the banded views the kernels
    consume and the Python mirror of eps."""
    module.refresh_bands()
    module._eps = float(module.eps)


class Magnitude(_Representation):
    def __repr__(self):
        if self.mel:
            return "Magnitude(mel=%s, n_fft=%s, norm=%s)" % (self.mel, self.n_fft, self.norm.mode)
        return "Magnitude(norm=%s)" % self.norm.mode

    def __init__(self, sr: int = 44100, mode: Optional[str] = "unipolar", contrast: Optional[str] = "log1p",
                 mel: bool = True, n_fft: int = 1024, dtype: Optional[torch.dtype] = None, eps: Optional[float] = None,
                 keep_nyquist: bool = True, norm: Optional[str] = None):
        # `norm=` is accepted as an alias of `mode=`: the reference's README passes it (README.md:54)
        super().__init__(sr=sr, mode=mode if norm is None else norm)
        self.contrast_mode = contrast
        self.mel = mel
        self.n_fft = n_fft
        if dtype is None:
            dtype = torch.get_default_dtype()
        if eps is None:
            eps = torch.finfo(dtype).eps
        self.register_buffer("eps", torch.tensor(eps))
        self._eps = float(eps)
        self.keep_nyquist = keep_nyquist
        assert sr is not None and n_fft is not None
        fwd, inv = build_mel_banks(sr, n_fft, keep_nyquist)
        self.register_buffer("mel_bank", fwd)
        self.register_buffer("inverse_mel_bank", inv)
        # banded views of the (99.6 % zero) banks, what the kernels consume; derived, hence not persistent
        self.register_buffer("mel_meta", torch.zeros(0, dtype=torch.int32), persistent=False)
        self.register_buffer("mel_coef", torch.zeros(0), persistent=False)
        self.register_buffer("inv_meta", torch.zeros(0, dtype=torch.int32), persistent=False)
        self.register_buffer("inv_coef", torch.zeros(0), persistent=False)
        self.refresh_bands()
        self.register_load_state_dict_post_hook(_magnitude_after_load)

    @torch.jit.unused
    def refresh_bands(self) -> None:
        """Re-derive the banded tensors from `mel_bank` / `inverse_mel_bank` (after construction or load_state_dict)."""
        dev = self.mel_bank.device
        meta, coef = _ops.BandedMatrix(self.mel_bank).tensors()
        self.mel_meta, self.mel_coef = meta.to(dev), coef.to(dev)
        meta, coef = _ops.BandedMatrix(self.inverse_mel_bank).tensors()
        self.inv_meta, self.inv_coef = meta.to(dev), coef.to(dev)

    def contrast(self, mag: torch.Tensor) -> torch.Tensor:
        """spectral_repr.py:191-201 — kept for API compatibility; `forward` applies it inside the kernel."""
        if self.contrast_mode == "log1p":
            return torch.log(1 + mag)
        if self.contrast_mode == "log":
            return torch.log(torch.clamp(mag, self._eps, None))
        if self.contrast_mode == "log10":
            return torch.log10(torch.clamp(mag, self._eps, None))
        if self.contrast_mode is None or self.contrast_mode == "none":
            return mag
        raise TypeError("unknown contrast type %s" % self.contrast_mode)

    def invert_contrast(self, mag: torch.Tensor) -> torch.Tensor:
        """spectral_repr.py:203-213."""
        if self.contrast_mode == "log1p":
            return torch.exp(mag) - 1
        if self.contrast_mode == "log":
            return torch.exp(mag) - self._eps
        if self.contrast_mode == "log10":
            return torch.pow(10.0, mag)
        if self.contrast_mode is None or self.contrast_mode == "none":
            return mag
        raise TypeError("unknown contrast type %s" % self.contrast_mode)

    @torch.jit.export
    def band_meta(self) -> Optional[torch.Tensor]:
        if self.mel:
            return self.mel_meta
        return None

    @torch.jit.export
    def band_coef(self) -> Optional[torch.Tensor]:
        if self.mel:
            return self.mel_coef
        return None

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(contrast(|x| @ mel_bank) - offset) / scale in one kernel (spectral_repr.py:215-226)."""
        return torch.ops.acids_b200.mag_epilogue(x, self.band_meta(), self.band_coef(), _contrast_id(self.contrast_mode),
                                                 self._eps, self.norm.get_offset(), self.norm.get_scale(),
                                                 not self.keep_nyquist)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        """contrast^-1(x * scale + offset) @ inverse_mel_bank (spectral_repr.py:228-240)."""
        meta: Optional[torch.Tensor] = None
        coef: Optional[torch.Tensor] = None
        if self.mel:
            meta = self.inv_meta
            coef = self.inv_coef
        return torch.ops.acids_b200.mag_invert(x, meta, coef, _contrast_id(self.contrast_mode), self._eps,
                                               self.norm.get_offset(), self.norm.get_scale(), not self.keep_nyquist)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        # statistics of contrast(|x|) WITHOUT the mel projection, like the reference (spectral_repr.py:242-245)
        self.norm.set_stats(torch.ops.acids_b200.stats(x, _contrast_id(self.contrast_mode), self._eps, True))

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        self.scale_data(X)
        mag = self.invert(self(X))
        rec = torch.ops.acids_b200.polar_to_complex(mag, X.angle())
        return {"direct": stft.invert(rec).reshape(batch + [-1])}


class Phase(_Representation):
    def __init__(self, sr: int = 44100, mode: Optional[str] = None, keep_nyquist: bool = True, unwrap: bool = False):
        super().__init__(sr=sr, mode=mode, keep_nyquist=keep_nyquist)
        self.unwrap = unwrap

    def __repr__(self):
        return "Phase(norm=%s, unwrap=%s)" % (self.norm.mode, self.unwrap)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """angle -> (unwrap over frames) -> normalise (spectral_repr.py:270-278)."""
        return torch.ops.acids_b200.phase_fwd(x, 1 if self.unwrap else 0, 0, False, self.norm.get_offset(),
                                              self.norm.get_scale(), not self.keep_nyquist)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        raw = torch.ops.acids_b200.phase_fwd(x, 1 if self.unwrap else 0, 0, False, None, None, False)
        self.norm.scale_data(raw)

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        self.scale_data(X)
        rec = torch.ops.acids_b200.polar_to_complex(X.abs(), self.invert(self(X)))
        return {"direct": stft.invert(rec).reshape(batch + [-1])}


class IF(_Representation):
    def __repr__(self):
        return "IF(method=%s, norm=%s)" % (self.method, self.norm.mode)

    def __init__(self, sr: int = 44100, mode: Optional[str] = "gaussian", method: Optional[str] = "forward",
                 weighted: bool = False, keep_nyquist: bool = True):
        super().__init__(sr=sr, mode=mode)
        self.method = method
        self.weighted = weighted
        self.keep_nyquist = keep_nyquist
        self.register_buffer("eps", torch.tensor(torch.finfo(torch.float32).eps))

    def get_if_methods(self):
        return ["backward", "forward", "central"]

    @torch.jit.export
    def get_if(self, data: torch.Tensor) -> torch.Tensor:
        """Un-normalised instantaneous frequency (spectral_repr.py:319-335)."""
        return torch.ops.acids_b200.phase_fwd(data, 2, _method_id(self.method), self.weighted, None, None, False)

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.norm.scale_data(self.get_if(x))

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.acids_b200.phase_fwd(x, 2, _method_id(self.method), self.weighted, self.norm.get_offset(),
                                              self.norm.get_scale(), not self.keep_nyquist)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        """De-normalise, undo the +-pi scaling, integrate over frames (spectral_repr.py:359-375)."""
        return torch.ops.acids_b200.phase_inv(x, 2, _method_id(self.method), self.norm.get_offset(), self.norm.get_scale(),
                                              not self.keep_nyquist)

    @torch.jit.export
    def invert_polar(self, x: torch.Tensor, mag: torch.Tensor) -> torch.Tensor:
        return torch.ops.acids_b200.phase_inv_polar(x, mag, 2, _method_id(self.method), self.norm.get_offset(),
                                                    self.norm.get_scale(), not self.keep_nyquist)

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        outs = {}
        for method in self.get_if_methods():
            self.method = method
            self.scale_data(X)
            rec = torch.ops.acids_b200.polar_to_complex(X.abs(), self.invert(self(X)))
            outs[method] = stft.invert(rec).reshape(batch + [-1])
        return outs


SpectralRepresentationType = Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]


class SpectralRepresentation(AudioTransform):
    """Two representations of the same spectrum stacked on dim `stack` (spectral_repr.py:399-484)."""

    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return True

    @property
    def needs_scaling(self):
        return True

    def __init__(self, sr: int = 44100, magnitude_transform=None, phase_transform=None, magnitude_args={}, phase_args={},
                 stack: Optional[int] = -2, keep_nyquist: bool = True):
        super().__init__(sr=sr)
        if type(self) == SpectralRepresentation:
            raise RuntimeError("SpectralRepresentation should not be called directly.")
        self.keep_nyquist = keep_nyquist
        self.magnitude = magnitude_transform(sr=sr, **magnitude_args, keep_nyquist=keep_nyquist)
        self.phase = phase_transform(sr=sr, **phase_args, keep_nyquist=keep_nyquist)
        self.stack = stack

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.magnitude.scale_data(x)
        self.phase.scale_data(x)

    def _stacked_forward(self, x: torch.Tensor) -> torch.Tensor:
        first = self.magnitude(x)
        second = self.phase(x)
        stack = self.stack
        if stack is None:
            raise RuntimeError("stack=None returns a tuple: call forward_unstacked()")
        return torch.stack([first, second], dim=stack)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._stacked_forward(x)

    def forward_unstacked(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.magnitude(x), self.phase(x)

    def _split(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        stack = self.stack
        if stack is None:
            raise RuntimeError("stack=None: pass the two tensors to the children's invert()")
        return x.select(stack, 0), x.select(stack, 1)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        first, second = self._split(x)
        mag = self.magnitude.invert(first)
        # mag * exp(i phase.invert(second)), spectral_repr.py:450-452; the phase transform does both in one pass
        return self.phase.invert_polar(second, mag)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        stft = STFT()
        if time is None:
            X = stft(x)
            self.scale_data(X)
            return self(X)
        X, time = stft.forward_with_time(x, time)
        self.scale_data(X)
        return self.forward_with_time(X, time)

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        self.scale_data(X)
        return {"direct": stft.invert(self.invert(self(X))).reshape(batch + [-1])}

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True):
        z = torch.randn(2, 10, 513) * torch.exp(2j * math.pi * torch.rand(2, 10, 513))
        transform.scale_data(z)
        y = transform(z)
        if invert:
            transform.invert(y)


class Cartesian(SpectralRepresentation):
    def __repr__(self):
        return "Cartesian(real_norm=%s, imag_norm=%s)" % (self.magnitude.norm.mode, self.phase.norm.mode)

    def __init__(self, sr: int = 44100, real_args={"mode": "gaussian"}, imag_args={"mode": "gaussian"}, stack: Optional[int] = -2,
                 keep_nyquist: bool = True):
        super().__init__(sr, Real, Imaginary, real_args, imag_args, stack=stack, keep_nyquist=keep_nyquist)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        first, second = self._split(x)
        return torch.complex(self.magnitude.invert(first), self.phase.invert(second))     # spectral_repr.py:494-508


class _PolarBase(SpectralRepresentation):
    """Magnitude + (Phase | IF): both kernels write straight into their slot of the stacked output."""

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        stack = self.stack
        if stack is None or (stack != -2 and stack != x.ndim - 1):
            return self._stacked_forward(x)
        return torch.ops.acids_b200.polar_fwd(x, self.magnitude.band_meta(), self.magnitude.band_coef(),
                                              _contrast_id(self.magnitude.contrast_mode), self.magnitude._eps,
                                              self.magnitude.norm.get_offset(), self.magnitude.norm.get_scale(),
                                              self._phase_mode(), self._phase_method(), self._phase_weighted(),
                                              self.phase.norm.get_offset(), self.phase.norm.get_scale(), not self.keep_nyquist)

    def _phase_mode(self) -> int:
        return 0

    def _phase_method(self) -> int:
        return 0

    def _phase_weighted(self) -> bool:
        return False


class Polar(_PolarBase):
    def __repr__(self):
        return "Polar(real_norm=%s, imag_norm=%s)" % (self.magnitude.norm.mode, self.phase.norm.mode)

    def __init__(self, sr: int = 44100, magnitude_args={"mode": "bipolar"}, phase_args={"mode": "bipolar"},
                 stack: Optional[int] = -2, keep_nyquist: bool = True):
        super().__init__(sr, Magnitude, Phase, magnitude_args, phase_args, stack=stack, keep_nyquist=keep_nyquist)

    def _phase_mode(self) -> int:
        return 1 if self.phase.unwrap else 0


class PolarIF(_PolarBase):
    def __repr__(self):
        return "PolarIF(real_norm=%s, imag_norm=%s)" % (self.magnitude.norm.mode, self.phase.norm.mode)

    def __init__(self, sr: int = 44100, magnitude_args={"mode": "bipolar"}, phase_args={"mode": "bipolar"},
                 stack: Optional[int] = -2, keep_nyquist: bool = True):
        super().__init__(sr, Magnitude, IF, magnitude_args, phase_args, stack=stack, keep_nyquist=keep_nyquist)

    def _phase_mode(self) -> int:
        return 2

    def _phase_method(self) -> int:
        return _method_id(self.phase.method)

    def _phase_weighted(self) -> bool:
        return self.phase.weighted

    def test_inversion(self, x: torch.Tensor):
        flat, batch = reshape_batches(x, -1)
        stft = STFT()
        X = stft(flat)
        outs = {}
        for method in self.phase.get_if_methods():
            self.phase.method = method
            self.scale_data(X)
            outs[method] = stft.invert(self.invert(self(X))).reshape(batch + [-1])
        return outs
