"""CPU port of the reference's hot path with the reference's own torch calls — the CPU BASELINE.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/np_oracle.py): imported by tests/ and by bench.py's
`cpu_baseline` / `--impl reference` legs, never by the product.

The reference is pure Python whose arithmetic is a sequence of eager torch CPU ops; it cannot travel to
the GPU box (/root/reference does not exist there and its sources may not be copied), so the baseline
re-issues exactly that sequence of library calls, multi-threaded like the reference (ATen/MKL/OpenMP):

  DGT.forward / STFT.forward   dgt.py:63-70, stft.py:97-104   torch.stft -> transpose -> angle() side effect
  Magnitude.forward            spectral_repr.py:215-226       abs -> matmul(mel_bank) -> log(1 + .) -> Normalize
  STFT.invert (complex)        stft.py:119-128                torch.istft

tests/test_oracle_golden.py pins these against the golden vectors.
"""
import math

import torch


def hann(n_fft):
    return torch.hann_window(n_fft)


def gaussian_window(n_fft):
    """dgt.py:108-112."""
    lam = (-torch.tensor([n_fft]).long() ** 2 / (8 * math.log(0.01))) ** .5
    n = torch.arange(0, 2 * n_fft + 1) - (2 * n_fft) / 2
    return torch.exp(-n ** 2 / (2 * (lam * 2) ** 2))[1:2 * n_fft + 1:2]


def stft_forward(x, window, n_fft, hop, phase_buffer=True):
    """stft.py:97-104: the reference also materialises angle(X) on every call (its phase_buffer)."""
    flat = x.reshape(-1, x.shape[-1])
    X = torch.stft(flat, n_fft=n_fft, hop_length=hop, window=window, return_complex=True, onesided=True).transpose(-2, -1)
    keep = X.angle() if phase_buffer else None
    return X.reshape(x.shape[:-1] + X.shape[-2:]), keep


def magnitude_forward(X, mel_bank, offset, scale):
    """spectral_repr.py:215-226 with contrast='log1p' and a fitted Normalize (norm.py:40-41)."""
    mag = X.abs()
    mag = torch.matmul(mag, mel_bank)
    mag = torch.log(1 + mag)
    return (mag - offset) / scale


def cfg2_forward(x, window, mel_bank, offset, scale, n_fft=1024, hop=256):
    X, _ = stft_forward(x, window, n_fft, hop)
    return magnitude_forward(X, mel_bank, offset, scale)


def istft(X, window, n_fft, hop):
    """stft.py:119-128."""
    flat = X.reshape((-1,) + X.shape[-2:])
    y = torch.istft(flat.transpose(-2, -1), n_fft=n_fft, hop_length=hop, window=window, onesided=True)
    return y.reshape(X.shape[:-2] + y.shape[-1:])
