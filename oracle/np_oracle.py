"""CPU oracle for the acids_transforms spectral hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``acids_transforms_b200/`` may import
this module: it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` leg of ``bench.py``; it is never the thing shipped.

Every function restates, in plain numpy, what the reference computes for one row
of SURVEY.md §8(a) and cites the reference ``file:line`` it follows (paths are
relative to ``/root/reference``; ``site-packages/...`` are the pinned third-party
sources the reference delegates to: torch 2.11.0, torchaudio 2.11.0).

Parity status: PINNED.  The reference holds no golden vectors of its own (its
tests are smoke tests), so this oracle is pinned against outputs
of the reference itself, imported in the build container from /root/reference
by ``tests/golden/make_golden.py``; the resulting fixtures live in
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks every function
here against them.

Arithmetic is float32 wherever the reference's is (``dtype=np.float32``); most
functions accept ``dtype=np.float64`` to produce a higher-precision "truth"
for error budgeting.
"""
from __future__ import annotations

import math
import numpy as np

PI32 = np.float32(math.pi)
TWO_PI32 = np.float32(2.0 * math.pi)
MAX_NFFT = 16384  # stft.py:10


# --------------------------------------------------------------------------
# windows  (A1, A2)
# --------------------------------------------------------------------------
def periodic_window(name: str, n: int, dtype=np.float32) -> np.ndarray:
    """``torch.<name>_window(n)`` with its default ``periodic=True``.

    stft.py:51-52, :80-81 — STFT stores ``getattr(torch, f"{window}_window")(n_fft)``.
    Generalised-cosine windows evaluated on the periodic grid 2*pi*k/n.
    """
    k = np.arange(n, dtype=np.float64)
    if name == "hann":
        w = 0.5 - 0.5 * np.cos(2 * np.pi * k / n)
    elif name == "hamming":
        w = 0.54 - 0.46 * np.cos(2 * np.pi * k / n)
    elif name == "blackman":
        w = 0.42 - 0.5 * np.cos(2 * np.pi * k / n) + 0.08 * np.cos(4 * np.pi * k / n)
    elif name == "bartlett":
        w = 1.0 - np.abs(2.0 * k / n - 1.0)
    else:
        raise ValueError("Window %s is not known" % name)  # stft.py:54
    return w.astype(dtype)


def dgt_window(n_fft: int, dtype=np.float32) -> np.ndarray:
    """Gaussian analysis window of the DGT.  dgt.py:108-112.

    lambda = sqrt(-N^2 / (8 ln 0.01)); w = exp(-n^2 / (2 (2 lambda)^2)) sampled
    at the half-integer grid n = (2k + 1 - N) (the reference samples a 2N+1 grid
    and keeps the odd entries).
    """
    f = np.dtype(dtype).type
    lam = f(math.sqrt(-float(n_fft) ** 2 / (8.0 * math.log(0.01))))
    n = (np.arange(0, 2 * n_fft + 1, dtype=np.float64) - n_fft).astype(dtype)
    den = f(2.0) * (lam * f(2.0)) ** 2
    w = np.exp(-(n * n) / den).astype(dtype)
    return w[1:2 * n_fft + 1:2]


def dgt_dual_window(window: np.ndarray, hop: int) -> np.ndarray:
    """Canonical dual (synthesis) window g[l] = w[l] / sum_k w[l - k H]^2.  dgt.py:114-123."""
    n = window.shape[0]
    w2 = window.astype(window.dtype) ** 2
    g = np.zeros_like(window)
    reach = n // hop
    for l in range(n):
        den = window.dtype.type(0)
        for k in range(-reach, reach + 1):
            dl = l - k * hop
            if 0 <= dl < n:
                den = den + w2[dl]
        g[l] = window[l] / den
    return g


def dgt_gamma(n_fft: int) -> float:
    """stft.py:77-78 / dgt.py:105-106 (only consumed by PGHI, kept for state_dict parity)."""
    return 2 * math.pi * (math.sqrt(-n_fft ** 2 / (8 * math.log(0.01)))) ** 2


# --------------------------------------------------------------------------
# STFT / ISTFT  (A3, A15, A17)
# --------------------------------------------------------------------------
def reflect_pad(x: np.ndarray, p: int) -> np.ndarray:
    """torch.stft(center=True, pad_mode='reflect'): site-packages/torch/functional.py stft wrapper."""
    if p == 0:
        return x
    if x.shape[-1] <= p:
        raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension")
    return np.concatenate([x[..., p:0:-1], x, x[..., -2:-p - 2:-1]], axis=-1)


def stft(x: np.ndarray, n_fft: int, hop: int, window: np.ndarray, dtype=np.float32) -> np.ndarray:
    """STFT.forward ≡ DGT.forward.  stft.py:97-104, dgt.py:63-70.

    x: [..., L] real → [..., T, F] complex, T = 1 + L // hop, F = n_fft/2 + 1,
    unnormalised, one-sided, frames centred (reflect padding by n_fft/2).
    """
    x = np.asarray(x, dtype=dtype)
    L = x.shape[-1]
    if not (0 < n_fft <= L + 2 * (n_fft // 2)):
        raise RuntimeError("stft: expected 0 < n_fft <= padded length")
    xp = reflect_pad(x, n_fft // 2)
    T = 1 + (xp.shape[-1] - n_fft) // hop
    idx = (np.arange(T) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[..., idx] * np.asarray(window, dtype=dtype)
    cdt = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    return np.fft.rfft(frames, axis=-1).astype(cdt)


def istft_envelope(T: int, n_fft: int, hop: int, window: np.ndarray) -> np.ndarray:
    """OLA of window^2 over T frames (untrimmed).  site-packages/torch/_refs/__init__.py:3780-3790."""
    env = np.zeros(n_fft + hop * (T - 1), dtype=window.dtype)
    w2 = window * window
    for t in range(T):
        env[t * hop:t * hop + n_fft] += w2
    return env


def istft(X: np.ndarray, n_fft: int, hop: int, window: np.ndarray, dtype=np.float32) -> np.ndarray:
    """STFT.invert / DGT.invert complex branch.  stft.py:119-128, dgt.py:85-93.

    X: [..., T, F] complex → [..., hop (T-1)] real:
    y = OLA(irfft(X_t) g) / OLA(g^2), trimmed by n_fft/2 on both sides
    (torch.istft defaults center=True, length=None; _refs/__init__.py:3760-3804).
    """
    X = np.asarray(X)
    T = X.shape[-2]
    window = np.asarray(window, dtype=dtype)
    fr = np.fft.irfft(X, n=n_fft, axis=-1).astype(dtype) * window
    out_len = n_fft + hop * (T - 1)
    y = np.zeros(X.shape[:-2] + (out_len,), dtype=dtype)
    for t in range(T):
        y[..., t * hop:t * hop + n_fft] += fr[..., t, :]
    env = istft_envelope(T, n_fft, hop, window)
    start, end = n_fft // 2, out_len - n_fft // 2
    env_t = env[start:end]
    if np.abs(env_t).min() <= 1e-11:
        raise RuntimeError("istft: window overlap add min: 1")  # same check as torch.istft
    return y[..., start:end] / env_t


def rt_stft_forward(frames: np.ndarray, window: np.ndarray, dtype=np.float32) -> np.ndarray:
    """RealtimeSTFT/RealtimeDGT.forward on pre-framed input.  stft.py:248-253, dgt.py:284-289."""
    cdt = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    return np.fft.rfft(np.asarray(frames, dtype) * np.asarray(window, dtype), axis=-1).astype(cdt)


def rt_stft_invert(X: np.ndarray, inv_window: np.ndarray, dtype=np.float32) -> np.ndarray:
    """RealtimeSTFT.invert complex branch: irfft(X) * inv_window.  stft.py:259-266, dgt.py:297-302."""
    n = inv_window.shape[0]
    return np.fft.irfft(X, n=n, axis=-1).astype(dtype) * np.asarray(inv_window, dtype)


def frame_times(T: int, hop: int, sr: int, time: np.ndarray) -> np.ndarray:
    """forward_with_time: t*hop/sr + time[..., None].  stft.py:106-117."""
    shifts = (np.arange(T, dtype=np.float32) * np.float32(hop)) / np.float32(sr)
    return shifts + np.asarray(time, np.float32)[..., None]


# --------------------------------------------------------------------------
# mel filterbanks  (A5, A9)
# --------------------------------------------------------------------------
def _linspace32(start: float, end: float, steps: int) -> np.ndarray:
    """float32 torch.linspace (symmetric evaluation: from start for the first half, from end after)."""
    f = np.float32
    if steps == 1:
        return np.array([start], dtype=f)
    s, e = f(start), f(end)
    step = (e - s) / f(steps - 1)
    i = np.arange(steps)
    lo = (s + step * i.astype(f)).astype(f)
    hi = (e - step * (steps - 1 - i).astype(f)).astype(f)
    return np.where(i < steps // 2, lo, hi).astype(f)


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """HTK triangular filterbank, no area norm: [n_freqs, n_mels] float32.

    site-packages/torchaudio/functional/functional.py:518-588 (+ :425-516 helpers),
    called by spectral_repr.py:177-178 and (via MelSpectrogram) mel.py:43-44.
    """
    f = np.float32
    all_freqs = _linspace32(0.0, float(sample_rate // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = _linspace32(m_min, m_max, n_mels + 2)
    f_pts = (f(700.0) * (np.power(f(10.0), m_pts / f(2595.0)).astype(f) - f(1.0))).astype(f)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(f(0), np.minimum(down, up)).astype(f)


def magnitude_banks(sr: int, n_fft: int, keep_nyquist: bool = True):
    """Forward (column-normalised) and inverse (row-normalised, transposed) square banks.

    spectral_repr.py:173-189.  Returns (mel_bank[F,F], inverse_mel_bank[F,F]).
    """
    f = np.float32
    n_bins = n_fft // 2 + 1
    fft_scale = (np.arange(n_bins, dtype=np.int64).astype(f) / f(n_fft) * f(sr)).astype(f)
    if not keep_nyquist:
        fft_scale = fft_scale[1:]
    fb = melscale_fbanks(n_bins, float(fft_scale[0]), float(fft_scale[-1]), n_bins, sr)
    col = fb.sum(0, dtype=f)
    fwd = fb / np.where(col != 0, col, f(1))[None, :]
    row = fb.sum(1, dtype=f)
    inv = fb / np.where(row != 0, row, f(1))[:, None]
    return fwd.astype(f), np.ascontiguousarray(inv.T).astype(f)


# --------------------------------------------------------------------------
# Normalize  (A7 statistics, forward/invert)
# --------------------------------------------------------------------------
def normalize_stats(x: np.ndarray, mode):
    """Normalize.scale_data → (offset, scale).  norm.py:26-38."""
    f = x.dtype.type
    if mode == "unipolar":
        mn = x.min()
        return f(mn), f((x - mn).max())
    if mode == "bipolar":
        mn, mx = x.min(), x.max()
        off = f((mx + mn) / f(2))
        return off, f(mx - off)
    if mode == "gaussian":
        x64 = x.astype(np.float64)
        return f(x64.mean()), f(x64.std(ddof=1))  # torch.std is unbiased
    return f(0), f(1)


def normalize_forward(x, offset, scale):
    """norm.py:40-41."""
    return (x - offset) / scale


def normalize_invert(x, offset, scale):
    """norm.py:43-44."""
    return x * scale + offset


# --------------------------------------------------------------------------
# Magnitude  (A6, A7, A8)
# --------------------------------------------------------------------------
def contrast(mag: np.ndarray, mode, eps) -> np.ndarray:
    """spectral_repr.py:191-201 (note: log(1 + m), not log1p)."""
    f = mag.dtype.type
    if mode == "log1p":
        return np.log(f(1) + mag)
    if mode == "log":
        return np.log(np.maximum(mag, f(eps)))
    if mode == "log10":
        return np.log10(np.maximum(mag, f(eps)))
    if mode is None or mode == "none":
        return mag
    raise TypeError("unknown contrast type %s" % mode)


def invert_contrast(y: np.ndarray, mode, eps) -> np.ndarray:
    """spectral_repr.py:203-213."""
    f = y.dtype.type
    if mode == "log1p":
        return np.exp(y) - f(1)
    if mode == "log":
        return np.exp(y) - f(eps)
    if mode == "log10":
        return np.power(f(10), y)
    if mode is None or mode == "none":
        return y
    raise TypeError("unknown contrast type %s" % mode)


def magnitude_forward(X, mel_bank, contrast_mode="log1p", eps=np.finfo(np.float32).eps,
                      offset=0.0, scale=1.0, keep_nyquist=True, dtype=np.float32):
    """Magnitude.forward.  spectral_repr.py:215-226.

    y = (c(|X| @ mel_bank) - offset) / scale; ``keep_nyquist=False`` drops bin 0 (sic).
    ``mel_bank=None`` ≡ ``mel=False``.
    """
    f = np.dtype(dtype).type
    mag = np.abs(X).astype(dtype)
    if mel_bank is not None:
        mag = (mag @ np.asarray(mel_bank, dtype)).astype(dtype)
    y = contrast(mag, contrast_mode, eps)
    y = normalize_forward(y, f(offset), f(scale)).astype(dtype)
    if not keep_nyquist:
        y = y[..., 1:]
    return y


def magnitude_stats_input(X, contrast_mode="log1p", eps=np.finfo(np.float32).eps, dtype=np.float32):
    """What Magnitude.scale_data feeds to Normalize: c(|X|) — no mel.  spectral_repr.py:242-245."""
    return contrast(np.abs(X).astype(dtype), contrast_mode, eps)


def magnitude_invert(y, inverse_mel_bank, contrast_mode="log1p", eps=np.finfo(np.float32).eps,
                     offset=0.0, scale=1.0, keep_nyquist=True, dtype=np.float32):
    """Magnitude.invert.  spectral_repr.py:228-240."""
    f = np.dtype(dtype).type
    m = normalize_invert(np.asarray(y, dtype), f(offset), f(scale))
    if not keep_nyquist:
        m = np.concatenate([m, np.zeros(m.shape[:-1] + (1,), dtype)], -1)
    m = invert_contrast(m, contrast_mode, eps)
    if inverse_mel_bank is not None:
        m = (m @ np.asarray(inverse_mel_bank, dtype)).astype(dtype)
    return m


# --------------------------------------------------------------------------
# Phase / unwrap / IF  (A10-A13)
# --------------------------------------------------------------------------
def angle(X, dtype=np.float32):
    return np.angle(X).astype(dtype)


def unwrap(phase: np.ndarray) -> np.ndarray:
    """numpy-style unwrap along axis -2 (frames), float32 running sum.  utils/misc.py:12-26."""
    f = phase.dtype.type
    pi, two_pi = f(math.pi), f(2 * math.pi)
    out = phase.copy()
    if phase.shape[-2] < 2:
        return out
    d = phase[..., 1:, :] - phase[..., :-1, :]
    dd = np.mod(d + pi, two_pi) - pi  # python/torch remainder: sign of the divisor
    dd = np.where((dd == -pi) & (d > 0), pi, dd)
    corr = dd - d
    corr = np.where(np.abs(d) < pi, f(0), corr)
    out[..., 1:, :] = phase[..., 1:, :] + np.cumsum(corr, axis=-2, dtype=phase.dtype)
    return out


def fdiff_forward(x):
    """utils/misc.py:65-68."""
    return np.concatenate([x[..., :1, :], (x[..., 1:, :] - x[..., :-1, :]) / x.dtype.type(2)], axis=-2)


def fdiff_backward(x):
    """utils/misc.py:71-75."""
    return fdiff_forward(x[..., ::-1, :])[..., ::-1, :]


def fdiff_central(x):
    """utils/misc.py:77-80."""
    return np.concatenate([x[..., :1, :], (x[..., 2:, :] - x[..., :-2, :]) / x.dtype.type(4), x[..., -1:, :]], axis=-2)


def fint_forward(x):
    """utils/misc.py:82-86 (rows[1:] doubled, then cumsum over frames)."""
    y = x.copy()
    y[..., 1:, :] = y[..., 1:, :] * x.dtype.type(2)
    return np.cumsum(y, axis=-2, dtype=x.dtype)


def fint_backward(x):
    """utils/misc.py:89-93."""
    return fint_forward(x[..., ::-1, :])[..., ::-1, :]


def fint_central(x):
    """utils/misc.py:96-104 (sequential; does not invert fdiff_central — reproduced as is)."""
    out = np.zeros_like(x)
    f4 = x.dtype.type(4)
    n = x.shape[-2]
    out[..., 0, :] = x[..., 0, :]
    out[..., -1, :] = x[..., -1, :]
    for i in range(2, n, 2):
        out[..., i, :] = out[..., i - 2, :] + f4 * x[..., i - 1, :]
    for i in range(n - 1, 0, -2):
        # i-2 may be negative: python indexing wraps exactly like torch's
        out[..., i - 2, :] = out[..., i, :] - f4 * x[..., i - 1, :]
    return out


def if_weight_window(T: int, dtype=np.float32):
    """Parabolic weighting over frames.  spectral_repr.py:337-345."""
    n = np.arange(T).astype(dtype)
    N = np.dtype(dtype).type(T)
    return ((1.5 * N) / (N ** 2 - 1) * (1 - ((n - (N / 2 - 1)) / (N / 2)) ** 2)).astype(dtype)


def inst_freq(X, method="forward", weighted=False, dtype=np.float32):
    """IF.get_if.  spectral_repr.py:319-335."""
    f = np.dtype(dtype).type
    pi = f(math.pi)
    ph = unwrap(angle(X, dtype))
    if method == "backward":
        y = fdiff_backward(ph).copy()
        y[..., 1:, :] /= -pi
    elif method == "forward":
        y = fdiff_forward(ph).copy()
        y[..., :-1, :] /= pi
    elif method == "central":
        y = fdiff_central(ph).copy()
        y[..., 1:-1, :] /= f(2) * pi
    else:
        raise AttributeError("method %s not known" % method)
    if weighted:
        y = y * if_weight_window(y.shape[-2], dtype)[:, None]
    return y.astype(dtype)


def if_forward(X, method="forward", weighted=False, offset=0.0, scale=1.0, keep_nyquist=True, dtype=np.float32):
    """IF.forward.  spectral_repr.py:351-357."""
    f = np.dtype(dtype).type
    y = normalize_forward(inst_freq(X, method, weighted, dtype), f(offset), f(scale)).astype(dtype)
    return y if keep_nyquist else y[..., 1:]


def if_invert(y, method="forward", offset=0.0, scale=1.0, keep_nyquist=True, dtype=np.float32):
    """IF.invert → unwrapped phase.  spectral_repr.py:359-375."""
    f = np.dtype(dtype).type
    pi = f(math.pi)
    x = (np.asarray(y, dtype) * f(scale) + f(offset)).astype(dtype)
    if method == "backward":
        x[..., 1:, :] *= -pi
        x = fint_backward(x)
    if method == "forward":
        x[..., :-1, :] *= pi
        x = fint_forward(x)
    elif method == "central":
        x[..., 1:-1, :] *= f(2) * pi
        x = fint_central(x)
    if not keep_nyquist:
        x = np.concatenate([x, np.zeros(x.shape[:-1] + (1,), dtype)], -1)
    return x


def phase_forward(X, do_unwrap=False, offset=0.0, scale=1.0, keep_nyquist=True, dtype=np.float32):
    """Phase.forward.  spectral_repr.py:270-278."""
    f = np.dtype(dtype).type
    p = angle(X, dtype)
    if do_unwrap:
        p = unwrap(p)
    p = normalize_forward(p, f(offset), f(scale)).astype(dtype)
    return p if keep_nyquist else p[..., 1:]


def polar_to_complex(mag, phase):
    """SpectralRepresentation.invert tail: mag * exp(i phase).  spectral_repr.py:452."""
    mag = np.asarray(mag)
    cdt = np.complex64 if mag.dtype == np.float32 else np.complex128
    return (mag * (np.cos(phase) + 1j * np.sin(phase))).astype(cdt)


# --------------------------------------------------------------------------
# MFCC (= MelSpectrogram in the reference) and the opt-in DCT variant  (A9)
# --------------------------------------------------------------------------
def mel_spectrogram(x, sr=44100, n_fft=1024, hop=256, n_mels=128, power=2.0, dtype=np.float32, fb=None):
    """MFCC.forward ≡ torchaudio MelSpectrogram: [..., L] → [..., n_mels, T].

    mel.py:38-44, :68-73; site-packages/torchaudio/transforms/_transforms.py:604-631, :417;
    functional.py:54-146.  Periodic Hann, centre/reflect, |X|^power, raw HTK bank.
    """
    w = periodic_window("hann", n_fft, dtype)
    X = stft(x, n_fft, hop, w, dtype)                       # [..., T, F]
    mag = np.abs(X).astype(dtype)
    spec = mag if power == 1.0 else np.power(mag, np.dtype(dtype).type(power))
    if fb is None:
        fb = melscale_fbanks(n_fft // 2 + 1, 0.0, float(sr // 2), n_mels, sr)
    mel = (spec @ np.asarray(fb, dtype)).astype(dtype)       # [..., T, M]
    return np.swapaxes(mel, -1, -2)


def create_dct(n_mfcc: int, n_mels: int) -> np.ndarray:
    """Ortho DCT-II matrix [n_mels, n_mfcc].  functional.py:636-667."""
    n = np.arange(n_mels, dtype=np.float32)
    k = np.arange(n_mfcc, dtype=np.float32)[:, None]
    dct = np.cos(np.float32(math.pi / float(n_mels)) * (n + np.float32(0.5)) * k).astype(np.float32)
    dct[0] *= np.float32(1.0 / math.sqrt(2.0))
    dct *= np.float32(math.sqrt(2.0 / float(n_mels)))
    return np.ascontiguousarray(dct.T)


def amplitude_to_db_power(mel, top_db=80.0):
    """AmplitudeToDB('power', top_db): 10 log10(clamp(x, 1e-10)), floor at per-item max - top_db.

    functional.py:390-404 with multiplier 10, amin 1e-10, db_multiplier log10(max(1e-10, 1.0)) = 0.
    The max is taken per leading-batch item over the packed (channel, mel, time) dims.
    """
    f = mel.dtype.type
    x_db = f(10) * np.log10(np.maximum(mel, f(1e-10)))
    if top_db is not None:
        shape = x_db.shape
        packed = shape[-3] if x_db.ndim > 2 else 1
        v = x_db.reshape(-1, packed, shape[-2], shape[-1])
        v = np.maximum(v, v.max(axis=(-3, -2, -1), keepdims=True) - f(top_db))
        x_db = v.reshape(shape)
    return x_db


def mfcc_dct(x, sr=44100, n_fft=1024, hop=256, n_mels=128, n_mfcc=40, dtype=np.float32, fb=None):
    """Opt-in variant following torchaudio.transforms.MFCC (log_mels=False): [..., L] → [..., n_mfcc, T].

    site-packages/torchaudio/transforms/_transforms.py:701-718.
    """
    mel = mel_spectrogram(x, sr, n_fft, hop, n_mels, 2.0, dtype, fb)
    db = amplitude_to_db_power(mel, 80.0)
    dct = create_dct(n_mfcc, n_mels).astype(dtype)
    return np.swapaxes(np.swapaxes(db, -1, -2) @ dct, -1, -2).astype(dtype)


# --------------------------------------------------------------------------
# mu-law, one-hot  (A18, A19)
# --------------------------------------------------------------------------
def mulaw_encode(x, channels=256, reciprocal_divide=False):
    """torchaudio mu_law_encoding in float32 op-by-op.  functional.py:690-700; raw.py:282-283.

    ``reciprocal_divide=True`` reproduces the CUDA eager chain, where dividing by a
    host scalar is computed as a multiplication by its float32 reciprocal.
    """
    f = np.float32
    x = np.asarray(x, f)
    mu = f(channels - 1.0)
    l1p = np.log1p(mu).astype(f)
    num = (np.sign(x) * np.log1p(mu * np.abs(x)).astype(f)).astype(f)
    q = (num * (f(1) / l1p)).astype(f) if reciprocal_divide else (num / l1p).astype(f)
    q = ((q + f(1)) / f(2)).astype(f)
    q = (q * mu).astype(f)
    q = (q + f(0.5)).astype(f)
    return q.astype(np.int64)  # C-style truncation; values are >= 0 for |x| <= 1


def mulaw_decode(q, channels=256):
    """torchaudio mu_law_decoding.  functional.py:723-729; raw.py:314-316."""
    f = np.float32
    mu = f(channels - 1.0)
    x = (np.asarray(q).astype(f) / mu).astype(f) * f(2) - f(1)
    l1p = np.log1p(mu).astype(f)
    return (np.sign(x) * (np.exp(np.abs(x) * l1p).astype(f) - f(1)) / mu).astype(f)


def one_hot(q, n_classes: int, layout="categorical"):
    """F.one_hot int64; 'channel' puts classes on dim -2.  raw.py:285-292, misc.py:176-179."""
    q = np.asarray(q, np.int64)
    oh = (q[..., None] == np.arange(n_classes, dtype=np.int64)).astype(np.int64)
    return np.ascontiguousarray(np.swapaxes(oh, -1, -2)) if layout == "channel" else oh


# --------------------------------------------------------------------------
# Mono / MidSide  (A20)
# --------------------------------------------------------------------------
def mono_mix(x):
    """Mono(mode='mix', squeeze=True): (L + R) / 2.  raw.py:34-49."""
    return (x.sum(-2) / x.dtype.type(2)).astype(x.dtype)


def midside_forward(x, pad_mid=True):
    """raw.py:145-162 (two-channel branch)."""
    f = x.dtype.type
    mid = (x[..., 0, :] + x[..., 1, :]) / f(2)
    side = (x[..., 0, :] - x[..., 1, :]) / f(2)
    if pad_mid:
        mid = mid / f(math.sqrt(2))
    return np.stack([mid, side], -2).astype(x.dtype)


def midside_invert(x, pad_mid=True):
    """raw.py:164-180."""
    f = x.dtype.type
    mid, side = x[..., 0, :], x[..., 1, :]
    if pad_mid:
        mid = mid * f(math.sqrt(2))
    return np.stack([mid + side, mid - side], -2).astype(x.dtype)


# --------------------------------------------------------------------------
# OverlapAdd (streaming)  (A16)
# --------------------------------------------------------------------------
def frame(x, wsize: int, hsize: int):
    """utils/misc.py:148-165 along the last axis (copying, zero-padded tail)."""
    L = x.shape[-1]
    n = (L - wsize) // hsize
    if L >= n * hsize + wsize:
        n += 1
    need = n * hsize + wsize
    if L <= need:
        x = np.concatenate([x, np.zeros(x.shape[:-1] + (need - L,), x.dtype)], -1)
    idx = (np.arange(n) * hsize)[:, None] + np.arange(wsize)[None, :]
    return x[..., idx]


class OverlapAddState:
    """Carry buffers of one OverlapAdd module.  oadd.py:23-31."""

    def __init__(self, n_fft=1024, hop=128):
        self.n_fft, self.hop = n_fft, hop
        self.frames_out = n_fft // hop - 1
        self.input_buffer = np.zeros(self.frames_out * hop, np.float32)
        self.output_buffer = np.zeros(self.frames_out * hop, np.float32)
        self.gain = np.float32(1.0)
        ones = np.ones((12, (self.frames_out + 1) * n_fft), np.float32)
        self.gain = self._invert_plain(frame(ones, n_fft, hop)).max()

    def _invert_plain(self, fr):
        """oadd.py:57-67."""
        n_fft, hop = self.n_fft, self.hop
        overlap = int(n_fft / hop)
        out = np.zeros(fr.shape[:-2] + (fr.shape[-2] * hop + n_fft,), np.float32)
        for i in range(fr.shape[-2]):
            out[..., i * hop:i * hop + n_fft] += fr[..., i, :] / np.float32(overlap / 2)
        return out / self.gain

    def forward(self, x):
        """oadd.py:33-42, :70-74."""
        keep = self.frames_out * self.hop
        if self.input_buffer.shape[:-1] != x.shape[:-1]:
            buf = np.zeros(x.shape[:-1] + (keep,), np.float32)
        else:
            buf = self.input_buffer.copy()
        self.input_buffer = x[..., -keep:].copy()
        return frame(np.concatenate([buf, x], -1), self.n_fft, self.hop)

    def invert(self, fr):
        """oadd.py:44-53, :91-104."""
        n_fft, hop = self.n_fft, self.hop
        keep = self.frames_out * hop
        if self.output_buffer.shape[:-1] != fr.shape[:-2]:
            buf = np.zeros(fr.shape[:-2] + (keep,), np.float32)
        else:
            buf = self.output_buffer.copy()
        tail = (fr.shape[-2] - 1) * hop + n_fft - keep
        rec = np.concatenate([buf, np.zeros(fr.shape[:-2] + (tail,), np.float32)], -1)
        for i in range(fr.shape[-2]):
            rec[..., i * hop:i * hop + n_fft] += fr[..., i, :]
        self.output_buffer = rec[..., -keep:].copy()
        return rec[..., :-keep] / self.gain
