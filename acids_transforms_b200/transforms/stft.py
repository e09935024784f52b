"""STFT / RealtimeSTFT — host-side mirror of acids_transforms/transforms/stft.py.

Same constructor, buffers (`n_fft, hop_length, window[16384], inv_window[16384], gamma, eps,
phase_buffer` — stft.py:41-48) and methods as the reference; the arithmetic is the fused
framing + window + real-FFT kernel (`acids_stft_fwd`) and the inverse rFFT + overlap-add kernel
(`acids_istft_ola`).  Differences that are deliberate and observable:

* `n_fft` / `hop_length` are mirrored in Python ints, so `forward` performs no device->host sync
  (the reference calls `.item()` on device buffers three times per call, stft.py:100-101);
* `phase_buffer` (stft.py:134-135: a full atan2 pass on every forward) is only filled when it can be
  consumed, i.e. when the inversion mode is "keep_input".  `STFT(..., track_phase=True)` restores the
  reference's behaviour (the buffer is rewritten by every forward); asking for "keep_input" on a module
  that has not been tracking warns, falls back to a random phase like the reference does without a buffer
  (stft.py:155-158), and turns tracking on for the following forwards;
* n_fft must be a power of two in [32, 16384];
* the CPU index tensors that crash the reference on CUDA (stft.py:110) are built on the data's device.
"""
import math
import warnings
from typing import Dict, List, Optional

import torch

from .base import AudioTransform, frame_times
from ..utils.misc import frame, reshape_batches
from .. import _torch_ops  # noqa: F401

__all__ = ["STFT", "RealtimeSTFT"]

MAX_NFFT = 16384


def make_window(name: str, n: int) -> torch.Tensor:
    """torch.<name>_window(n), periodic (stft.py:51-52, :80-81)."""
    if name == "hann":
        return torch.hann_window(n)
    if name == "hamming":
        return torch.hamming_window(n)
    if name == "blackman":
        return torch.blackman_window(n)
    if name == "bartlett":
        return torch.bartlett_window(n)
    if name == "kaiser":
        return torch.kaiser_window(n)
    raise ValueError("Window %s is not known" % name)


def pghi_gamma(n_fft: int) -> float:
    # stft.py:77-78 (consumed by PGHI only; kept for state_dict parity)
    return 2 * math.pi * (math.sqrt(-float(n_fft) ** 2 / (8 * math.log(0.01)))) ** 2


class STFT(AudioTransform):
    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return True

    @property
    def needs_scaling(self):
        return False

    def __repr__(self):
        return "STFT(n_fft=%d, hop_length=%d, inversion_mode = %s)" % (self._n_fft, self._hop, self.inversion_mode)

    def __init__(self, sr: int = 44100, n_fft: int = 1024, hop_length: int = 256, dtype: Optional[torch.dtype] = None,
                 inversion_mode: str = "griffin_lim", window: str = "hann", track_phase: Optional[bool] = None):
        super().__init__(sr=sr)
        self._init_buffers(dtype)
        make_window(window, 8)                       # raises ValueError for an unknown window (stft.py:54)
        self.window_name = window
        self._finish_init(n_fft, hop_length, inversion_mode, track_phase)

    # ---- construction helpers shared with DGT ----
    def _init_buffers(self, dtype: Optional[torch.dtype]):
        dtype = dtype or torch.get_default_dtype()
        self.register_buffer("n_fft", torch.zeros(1).long())
        self.register_buffer("hop_length", torch.zeros(1).long())
        self.register_buffer("window", torch.zeros(MAX_NFFT))
        self.register_buffer("inv_window", torch.zeros(MAX_NFFT))
        self.register_buffer("gamma", torch.zeros(1))
        self.register_buffer("eps", torch.tensor(torch.finfo(dtype).eps, dtype=dtype))
        self.register_buffer("phase_buffer", torch.zeros(0))
        self._n_fft = 0
        self._hop = 0
        self.track_phase = False
        self.inversion_mode = ""
        self.register_load_state_dict_post_hook(_sync_ints_after_load)

    def _finish_init(self, n_fft, hop_length, inversion_mode, track_phase: Optional[bool] = None):
        if n_fft is not None:
            assert hop_length is not None, "n_fft and hop_length must be given together"
        if hop_length is not None:
            assert n_fft is not None, "n_fft and hop_length must be given together"
        if n_fft is not None and hop_length is not None:
            self.set_params(int(n_fft), int(hop_length))
        if inversion_mode in type(self).get_inversion_modes():
            self.inversion_mode = inversion_mode
        else:
            raise ValueError("Inversion mode %s not known" % inversion_mode)
        # None: fill `phase_buffer` only when this module's own mode can consume it; True: on every forward, like the
        # reference (stft.py:103, :134-135); False: never
        self.track_phase = (inversion_mode == "keep_input") if track_phase is None else bool(track_phase)

    @torch.jit.export
    def set_params(self, n_fft: int, hop_length: int) -> None:
        if n_fft > 16384:          # MAX_NFFT, stft.py:10
            raise ValueError("n_fft above 16384")
        self._n_fft = n_fft
        self._hop = hop_length
        self.n_fft.fill_(n_fft)
        self.hop_length.fill_(hop_length)
        self.window.zero_()
        self.inv_window.zero_()
        self.window[:n_fft] = self._get_window().to(self.window.device)
        self.inv_window[:n_fft] = self._get_dual_window().to(self.window.device)
        self.gamma.fill_(self._gamma_value(n_fft))

    def _gamma_value(self, n_fft: int) -> float:
        return pghi_gamma(n_fft)

    def _get_window(self) -> torch.Tensor:
        return make_window(self.window_name, self._n_fft)

    def _get_dual_window(self) -> torch.Tensor:
        return self._get_window()          # stft.py:83-84

    @property
    def ratio(self):
        return self._hop

    @torch.jit.export
    def set_inversion_mode(self, inversion_mode: str) -> None:
        if inversion_mode in self.get_inversion_modes():
            self.inversion_mode = inversion_mode
            if inversion_mode == "keep_input":
                self.track_phase = True
        else:
            raise AttributeError("inversion mode %s not valid" % inversion_mode)

    @staticmethod
    def get_inversion_modes() -> List[str]:
        return ["griffin_lim", "keep_input", "random", "sinebank"]

    # ---- forward ----
    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[..., L] -> complex64 [..., 1 + L // hop, n_fft/2 + 1]  (stft.py:97-104)."""
        y = torch.ops.acids_b200.stft_fwd(x, self.window, self._n_fft, self._hop, True)
        if self.track_phase:
            self._replace_phase_buffer(y)
        return y

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-2), self._hop, self.sr, time)

    def _replace_phase_buffer(self, y: torch.Tensor) -> None:
        # flattened [B, T, F] like the reference's x_fft.angle() taken before the batch reshape (stft.py:103)
        flat = y.reshape([-1, y.size(-2), y.size(-1)])
        self.phase_buffer = torch.ops.acids_b200.phase_fwd(flat, 0, 0, False, None, None, False)

    def _get_phase_buffer(self, mag: torch.Tensor) -> torch.Tensor:
        if mag.shape[:-2] != self.phase_buffer.shape[:-2]:
            self.phase_buffer = torch.zeros(0, device=mag.device)
        return self.phase_buffer

    # ---- inverse ----
    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        """complex [..., T, F] -> [..., hop (T-1)]  (stft.py:119-128); real input: phaseless modes."""
        if torch.is_complex(x):
            return torch.ops.acids_b200.istft_ola(x, self.inv_window, self._n_fft, self._hop)
        flat, batch = reshape_batches(x, -2)
        y = self.invert_without_phase(flat, inversion_mode)
        return y.reshape(batch + [y.size(-1)])

    def _istft(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.acids_b200.istft_ola(x, self.inv_window, self._n_fft, self._hop)

    def _random_phase_istft(self, mag: torch.Tensor) -> torch.Tensor:
        phase = 2 * math.pi * torch.rand_like(mag)
        return self._istft(torch.ops.acids_b200.polar_to_complex(mag, phase))

    def invert_without_phase(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        """stft.py:149-172."""
        mode = self.inversion_mode if inversion_mode is None else inversion_mode
        return self._phaseless(x, mode)

    def _phaseless(self, x: torch.Tensor, mode: str) -> torch.Tensor:
        if mode == "keep_input":
            phase = self._get_phase_buffer(x)
            if phase.size(0) == 0:
                if not self.track_phase:
                    warnings.warn("keep_input: this module has not been recording the phase of its input (track_phase "
                                  "is off unless inversion_mode='keep_input' or track_phase=True is given at construction); "
                                  "falling back to a random phase and recording from the next forward on")
                    self.track_phase = True
                return self._random_phase_istft(x)
            return self._istft(torch.ops.acids_b200.polar_to_complex(x, phase.to(x.device)))
        if mode == "griffin_lim":
            return self.griffin_lim(x)
        if mode == "random":
            return self._random_phase_istft(x)
        if mode == "sinebank":
            return self.get_sinebank_inversion(x)
        raise ValueError("inversion mode %s not valid." % mode)

    def griffin_lim(self, mag: torch.Tensor, n_iter: int = 30, momentum: float = 0.99) -> torch.Tensor:
        """Fast Griffin-Lim (torchaudio griffinlim, functional.py:297-353, as called at stft.py:174-178:
        power 1, 30 iterations, momentum 0.99, random init) on the new ISTFT / STFT kernels."""
        x = torch.ops.acids_b200.polar_to_complex(mag, 2 * math.pi * torch.rand_like(mag))     # mag * exp(i random phase)
        tprev = torch.zeros_like(x)
        for _ in range(n_iter):
            wave = self._istft(x)
            rebuilt = torch.ops.acids_b200.stft_fwd(wave, self.inv_window, self._n_fft, self._hop, True)
            # the trimmed ISTFT output is one hop shorter than the frames it came from
            if rebuilt.size(-2) < mag.size(-2):
                rebuilt = torch.nn.functional.pad(rebuilt, [0, 0, 0, mag.size(-2) - rebuilt.size(-2)])
            # angles = rebuilt - mom * tprev; x = mag * angles / (|angles| + 1e-16): one kernel
            x = torch.ops.acids_b200.griffinlim_update(rebuilt, tprev, mag, momentum)
            tprev = rebuilt
        return self._istft(x)

    def get_sinebank_inversion(self, x_fft: torch.Tensor) -> torch.Tensor:
        """Additive resynthesis with one sinusoid per bin (stft.py:180-191); plain torch glue, off the hot path."""
        dev = x_fft.device
        lead = [1] * (x_fft.ndim - 2)
        n_bins = self._n_fft // 2 + 1
        freqs = torch.linspace(0, self.sr / 2, n_bins, device=dev).view(lead + [-1, 1])
        phase0 = 2 * math.pi * torch.rand(n_bins, 1, device=dev)
        x_fft = x_fft / x_fft.abs().max()
        length = self._hop * x_fft.size(-2) + self._n_fft
        t = torch.linspace(0, length / self.sr, length, device=dev).view(lead + [1, -1])
        env = torch.nn.functional.interpolate(x_fft.transpose(-2, -1), length, mode="linear") / (2 * math.pi)
        y = (env * torch.sin(2 * math.pi * freqs * t + phase0)).sum(-2)
        return y / y.max()

    def realtime(self):
        mode = self.inversion_mode if self.inversion_mode in RealtimeSTFT.get_inversion_modes() else "random"
        return RealtimeSTFT(sr=self.sr, n_fft=self._n_fft, hop_length=self._hop, inversion_mode=mode,
                            window=self.window_name)

    # ---- reference test hooks (stft.py:194-212) ----
    def test_inversion(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        outs = {}
        x_stft = self.forward(x)
        outs["direct"] = self.invert(x_stft)
        for mode in self.get_inversion_modes():
            outs[mode] = self.invert(x_stft.abs(), inversion_mode=mode)
        return outs

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True, batch_size=(2, 2)):
        x = torch.zeros(*batch_size, 44100)
        time = torch.zeros(*batch_size)
        transform.forward(x)
        x_t, _ = transform.forward_with_time(x, time)
        if invert:
            transform.invert(x_t)


def _sync_ints_after_load(module, incompatible_keys):
    """load_state_dict may change the n_fft / hop_length buffers: refresh their Python mirrors."""
    module._n_fft = int(module.n_fft.item())
    module._hop = int(module.hop_length.item())


def realtime_sinebank(x_fft: torch.Tensor, random_phase: torch.Tensor, time_index: torch.Tensor, n_fft: int, hop: int,
                      sr: int):
    """One sinusoid per bin with a running time origin (stft.py:276-291, dgt.py reuses it); returns
    (frames [..., n, n_fft], random_phase, time_index)."""
    dev = x_fft.device
    batch = list(x_fft.shape[:-2])
    lead = [1] * len(batch)
    n_bins = x_fft.size(-1)
    if random_phase.ndim < 2 or list(random_phase.shape[:-2]) != batch:
        random_phase = 2 * math.pi * torch.rand(batch + [1, n_bins], device=dev)
    t = torch.arange(n_fft, device=dev).unsqueeze(0) + torch.arange(x_fft.size(-2), device=dev).unsqueeze(1) * hop
    t = (t / sr + time_index.to(dev)).view(lead + [t.size(0), 1, t.size(1)])
    freqs = torch.linspace(0, sr / 2, n_bins, device=dev).view(lead + [1, -1, 1])
    sines = torch.sin(2 * math.pi * freqs * t + random_phase.to(dev).view(batch + [1, -1, 1]))
    y = (x_fft.unsqueeze(-1) * sines).sum(-2) / n_bins
    time_index = time_index + (x_fft.size(-2) * hop + n_fft) / sr
    return y, random_phase, time_index


class RealtimeSTFT(STFT):
    """Per-frame transform of pre-framed input (stft.py:215-362): [..., n, n_fft] <-> [..., n, F]."""

    def __init__(self, sr: int = 44100, n_fft: int = 1024, hop_length: int = 256, dtype: Optional[torch.dtype] = None,
                 inversion_mode: Optional[str] = "random", window: str = "hann", batch_size: int = 2):
        super().__init__(sr=sr, n_fft=n_fft, hop_length=hop_length, dtype=dtype, inversion_mode=inversion_mode, window=window)
        self.batch_size = batch_size
        self.register_buffer("random_phase", 2 * math.pi * torch.rand(self._n_fft // 2 + 1))
        self.register_buffer("time_index", torch.tensor(0.))

    def __repr__(self):
        return "RealtimeSTFT(n_fft=%d, hop_length=%d, inversion_mode = %s)" % (self._n_fft, self._hop, self.inversion_mode)

    @staticmethod
    def get_inversion_modes() -> List[str]:
        return ["keep_input", "random", "sinebank"]

    def reset(self, x=None):
        self.time_index = torch.tensor(0., device=self.time_index.device)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = torch.ops.acids_b200.stft_fwd(x, self.window, self._n_fft, self._n_fft, False)     # rfft(x * window), stft.py:248-253
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
            return torch.ops.acids_b200.irfft_frames(x, self.inv_window, self._n_fft)       # irfft(x) * inv_window, stft.py:259-266
        return self.invert_without_phase(x, inversion_mode)

    @torch.jit.export
    def get_batch_size(self, batch_size: int):
        return batch_size

    @torch.jit.export
    def set_batch_size(self, batch_size: int):
        self.batch_size = batch_size

    def invert_without_phase(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        """stft.py:293-310."""
        mode = self.inversion_mode if inversion_mode is None else inversion_mode
        if mode == "keep_input":
            phase = self._get_phase_buffer(x)
            if phase.size(0) == 0 or phase.shape != x.shape:
                phase = 2 * math.pi * torch.rand_like(x)
        elif mode == "random":
            phase = 2 * math.pi * torch.rand_like(x)
        elif mode == "sinebank":
            return self.get_sinebank_inversion(x) * self.inv_window[:self._n_fft].to(x.device)
        else:
            raise ValueError("inversion mode %s not valid." % mode)
        z = torch.ops.acids_b200.polar_to_complex(x, phase.to(x.device))
        return torch.ops.acids_b200.irfft_frames(z, self.inv_window, self._n_fft)

    def get_sinebank_inversion(self, x_fft: torch.Tensor) -> torch.Tensor:
        """stft.py:276-291 — torch glue with the index tensors on the data's device."""
        y, self.random_phase, self.time_index = realtime_sinebank(x_fft, self.random_phase, self.time_index, self._n_fft,
                                                                  self._hop, self.sr)
        return y

    # ---- reference test hooks (stft.py:313-362) ----
    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        fr = frame(x, self._n_fft, self._hop, -1)
        y = self.forward(fr)          # all frames in one launch (the reference loops over them in Python)
        return y if time is None else (y, None)

    def test_inversion(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .oadd import OverlapAdd
        self.reset()
        chunk = self._n_fft * 4
        outs = {}
        for mode in ["direct", "sinebank"]:
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
        x = torch.zeros(2, int(transform.n_fft.item()))
        x_t = transform(x)
        if invert:
            transform.invert(x_t)
