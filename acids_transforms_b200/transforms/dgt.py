"""DGT / RealtimeDGT — discrete Gabor transform: the STFT with a Gaussian analysis window and its
canonical dual as synthesis window (acids_transforms/transforms/dgt.py:24-123, :238-302).

The analysis / complex-inverse path is the same pair of kernels as STFT.  Phase-gradient heap
integration (PGHI, dgt.py:156-236, :338-466) is a sequential priority-queue
flood fill per clip: device tensors go through csrc/pghi.cu (one CTA per clip or stream, the
reference's visiting order and float32 operation order), host tensors through the numpy + heapq
restatement (transforms/pghi.py) that also pins the kernels in the tests; the phase is recombined
with the magnitude and inverted by the CUDA kernels.  Like the reference's, it is eager-only (the reference scripts a TorchScript heap; here scripted modules
support every other inversion mode and raise for "pghi").
"""
import math
from typing import Dict, List, Optional, Union

import torch

from .stft import STFT, RealtimeSTFT, MAX_NFFT, realtime_sinebank
from .base import AudioTransform
from ..utils.misc import frame
from .. import _torch_ops  # noqa: F401
from . import pghi as _pghi

__all__ = ["DGT", "RealtimeDGT"]


def gaussian_window(n_fft: int) -> torch.Tensor:
    """w[k] = exp(-(2k + 1 - N)^2 / (2 (2 lambda)^2)), lambda = sqrt(-N^2 / (8 ln 0.01))  (dgt.py:108-112)."""
    lam = (-torch.tensor([n_fft]).long() ** 2 / (8 * math.log(0.01))) ** .5
    n = torch.arange(0, 2 * n_fft + 1) - (2 * n_fft) / 2
    w = torch.exp(-n ** 2 / (2 * (lam * 2) ** 2))
    return w[1:2 * n_fft + 1:2]


def canonical_dual(window: torch.Tensor, hop: int) -> torch.Tensor:
    """g[l] = w[l] / sum_k w[l - k H]^2  (dgt.py:114-123), vectorised: the denominator only depends on
    l mod H, so it is one reshape-and-sum instead of the reference's O(N * N/H) Python double loop."""
    n = window.numel()
    w2 = window.double() ** 2
    den = torch.zeros(n, dtype=torch.float64)
    for k in range(-(n // hop), n // hop + 1):
        lo, hi = max(0, k * hop), min(n, n + k * hop)
        if hi > lo:
            den[lo:hi] += w2[lo - k * hop:hi - k * hop]
    return (window.double() / den).to(window.dtype)


class DGT(STFT):
    def __repr__(self):
        return "DGT(n_fft=%d, hop_length=%d, inversion_mode = %s)" % (self._n_fft, self._hop, self.inversion_mode)

    def __init__(self, sr: int = 44100, n_fft: int = 1024, hop_length: int = 256, dtype: Optional[torch.dtype] = None,
                 inversion_mode: Optional[str] = "pghi", tolerance: float = 1.e-2, track_phase: Optional[bool] = None):
        AudioTransform.__init__(self, sr)
        self._init_buffers(dtype)
        self.window_name = "gaussian"
        self.register_buffer("tolerance", torch.tensor(tolerance))
        self._finish_init(n_fft, hop_length, inversion_mode, track_phase)

    @staticmethod
    def get_inversion_modes() -> List[str]:
        return ["griffin_lim", "keep_input", "random", "sinebank", "pghi"]

    def _get_window(self) -> torch.Tensor:
        return gaussian_window(self._n_fft)

    def _get_dual_window(self) -> torch.Tensor:
        return canonical_dual(self._get_window(), self._hop)

    def realtime(self):
        mode = self.inversion_mode if self.inversion_mode in RealtimeDGT.get_inversion_modes() else "pghi"
        return RealtimeDGT(sr=self.sr, n_fft=self._n_fft, hop_length=self._hop, inversion_mode=mode)

    def invert_without_phase(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        mode = self.inversion_mode if inversion_mode is None else inversion_mode
        if mode == "pghi":
            return self._pghi_istft(x)
        return self._phaseless(x, mode)

    @torch.jit.unused
    def pghi(self, mag: torch.Tensor, tolerance: float = 1.e-4) -> torch.Tensor:
        """Phase of a [..., T, F] magnitude by heap integration (dgt.py:156-162): the CUDA kernel (csrc/pghi.cu, one CTA per
        clip, the reference's visiting and operation order) for device tensors; host tensors keep the numpy + heapq
        restatement (transforms/pghi.py), which is also what pins the kernel in the tests."""
        if mag.is_cuda:
            from .. import ops
            return ops.pghi(mag, float(self.gamma), self._n_fft, self._hop, float(tolerance), float(self.eps))
        if mag.dim() > 2:
            return torch.stack([self.pghi(m, tolerance) for m in mag])
        return _pghi.pghi(mag, float(self.gamma), self._n_fft, self._hop, float(tolerance), float(self.eps))

    @torch.jit.unused
    def _pghi_istft(self, x: torch.Tensor) -> torch.Tensor:
        # dgt.py:136-143: one flood fill per clip (all clips of a device batch in one launch), then mag * exp(i phase) -> ISTFT
        # with the dual window
        phase = self.pghi(x, float(self.tolerance))
        return self._istft(torch.ops.acids_b200.polar_to_complex(x.contiguous(), phase.to(x.device)))

    def test_inversion(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        outs = {}
        x_dgt = self.forward(x)
        outs["direct"] = self.invert(x_dgt)
        for mode in ["griffin_lim", "keep_input", "random", "sinebank", "pghi"]:
            outs[mode] = self.invert(x_dgt.abs(), inversion_mode=mode)
        return outs


class RealtimeDGT(DGT):
    """Per-frame DGT of pre-framed input (dgt.py:238-302)."""

    def __init__(self, sr: int = 44100, n_fft: int = 1024, hop_length: int = 256, dtype: Optional[torch.dtype] = None,
                 batch_size: Union[int, List[int]] = 2, inversion_mode: Optional[str] = "pghi"):
        super().__init__(sr=sr, n_fft=n_fft, hop_length=hop_length, dtype=dtype, inversion_mode=inversion_mode)
        self._batch = [batch_size] if isinstance(batch_size, int) else list(batch_size)
        self.register_buffer("hgi_mag_buffer", torch.zeros(*self._batch, 2, n_fft // 2 + 1))
        self.register_buffer("hgi_phase_buffer", torch.zeros(*self._batch, n_fft // 2 + 1))
        self.register_buffer("random_phase", 2 * math.pi * torch.rand(n_fft // 2 + 1))
        self.register_buffer("time_index", torch.tensor(0.))

    def __repr__(self):
        return "RealtimeDGT(n_fft=%d, hop_length=%d, inversion_mode = %s)" % (self._n_fft, self._hop, self.inversion_mode)

    @staticmethod
    def get_inversion_modes() -> List[str]:
        return ["random", "pghi", "keep_input", "sinebank"]

    @torch.jit.export
    def get_batch_size(self) -> List[int]:
        return [int(b) for b in self._batch]

    @torch.jit.export
    def reset(self, batch_size: List[int]) -> None:
        self._batch = [int(b) for b in batch_size]
        dev = self.hgi_mag_buffer.device
        self.hgi_mag_buffer = torch.zeros(self._batch + [2, self._n_fft // 2 + 1], device=dev)
        self.hgi_phase_buffer = torch.zeros(self._batch + [self._n_fft // 2 + 1], device=dev)

    @torch.jit.export
    def set_batch_size(self, batch_size: List[int]) -> None:
        self.reset(batch_size)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = torch.ops.acids_b200.stft_fwd(x, self.window, self._n_fft, self._n_fft, False)     # dgt.py:284-289
        if self.track_phase:
            self.phase_buffer = torch.ops.acids_b200.phase_fwd(y.reshape([-1, 1, y.size(-1)]), 0, 0, False, None, None,
                                                               False).reshape(y.shape)
        return y

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        return self.forward(x), time

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        if torch.is_complex(x):
            return torch.ops.acids_b200.irfft_frames(x, self.inv_window, self._n_fft)       # dgt.py:297-302
        return self.invert_without_phase(x, inversion_mode)

    def _gamma_value(self, n_fft: int) -> float:
        # dgt.py:373-374: the real-time class overrides gamma with lambda itself (not 2 pi lambda^2)
        return math.sqrt(-float(n_fft) ** 2 / (8 * math.log(0.01)))

    def invert_without_phase(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        """dgt.py:304-327: every mode but "sinebank" also refreshes the two-frame PGHI history."""
        mode = self.inversion_mode if inversion_mode is None else inversion_mode
        if list(x.shape[:-2]) != self._batch:
            self.reset(list(x.shape[:-2]))
        if mode == "keep_input":
            phase = self._get_phase_buffer(x)
            if phase.size(0) == 0 or phase.shape != x.shape:
                phase = 2 * math.pi * torch.rand_like(x)
        elif mode == "pghi":
            phase = self._rt_pghi(x)
        elif mode == "random":
            phase = 2 * math.pi * torch.rand_like(x)
        elif mode == "sinebank":
            return self._rt_sinebank(x) * self.inv_window[:self._n_fft].to(x.device)
        else:
            raise ValueError("inversion mode %s not valid." % mode)
        phase = phase.to(x.device)
        self._update_hgi_buffers(x, phase)
        z = torch.ops.acids_b200.polar_to_complex(x.contiguous(), phase.contiguous())
        return torch.ops.acids_b200.irfft_frames(z, self.inv_window, self._n_fft)

    @torch.jit.unused
    def _rt_pghi(self, x: torch.Tensor) -> torch.Tensor:
        # dgt.py:338-357: the new frames are appended to the two remembered frames, one flood fill per clip
        n_bins = x.size(-1)
        flat = x.reshape(-1, x.size(-2), n_bins)
        hist_mag = self.hgi_mag_buffer.reshape(-1, 2, n_bins)
        hist_phase = self.hgi_phase_buffer.reshape(-1, n_bins)
        if flat.is_cuda:
            # csrc/pghi.cu rt_pghi_kernel: one CTA per stream, no trip to the host (the three scalar buffers are read back
            # once per value, not once per block); quiet bins take a normal random phase
            from .. import ops
            gamma, tol, eps = self._pghi_scalars()
            ph = ops.rt_pghi(flat, hist_mag, hist_phase, gamma, self._n_fft, self._hop, tol, eps, noise=torch.randn_like(flat))
        else:
            ph = _pghi.rt_pghi(flat, hist_mag, hist_phase, float(self.gamma), self._n_fft, self._hop, float(self.tolerance),
                               float(self.eps))
        return ph.reshape(x.shape)

    def _pghi_scalars(self):
        """(gamma, tolerance, eps) as Python floats, re-read from the device buffers only when one of them changed."""
        bufs = (self.gamma, self.tolerance, self.eps)
        key = tuple((b.data_ptr(), b._version) for b in bufs)
        if getattr(self, "_pghi_scalar_key", None) != key:
            self._pghi_scalar_val = tuple(float(v) for v in torch.stack([b.reshape(()).float() for b in bufs]).cpu())
            self._pghi_scalar_key = key
        return self._pghi_scalar_val

    def _update_hgi_buffers(self, mag: torch.Tensor, phase: torch.Tensor) -> None:
        # dgt.py:329-336, on (|x|, angle(x)) of x = mag * exp(i phase): the angle is the phase wrapped to (-pi, pi]
        if mag.size(-2) > 1:
            self.hgi_mag_buffer = mag[..., -2:, :].abs()
        else:
            self.hgi_mag_buffer = torch.stack([self.hgi_mag_buffer[..., 1, :].to(mag.device), mag[..., -1, :].abs()], -2)
        last = phase[..., -1, :]
        self.hgi_phase_buffer = torch.atan2(torch.sin(last), torch.cos(last))

    def _rt_sinebank(self, x_fft: torch.Tensor) -> torch.Tensor:
        y, self.random_phase, self.time_index = realtime_sinebank(x_fft, self.random_phase, self.time_index, self._n_fft,
                                                                  self._hop, self.sr)
        return y

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        y = self.forward(frame(x, self._n_fft, self._hop, -1))
        return y if time is None else (y, None)

    def test_inversion(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .oadd import OverlapAdd
        self.reset(list(x.shape[:-1]))
        chunk = self._n_fft * 4
        outs = {}
        for mode in ["direct", "random", "keep_input", "sinebank", "pghi"]:
            oadd = OverlapAdd(self._n_fft, self._hop)
            pieces = []
            for part in x.split(chunk, -1):
                spec = self.forward(oadd(part))
                frames = self.invert(spec) if mode == "direct" else self.invert(spec.abs(), inversion_mode=mode)
                pieces.append(oadd.invert(frames))
            outs[mode] = torch.cat(pieces, -1)
        return outs

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True):
        x = torch.zeros(2, 1, int(transform.n_fft.item()))
        transform.reset(list(x.shape[:-1]))
        x_t = transform(x)
        if invert:
            transform.invert(x_t)
