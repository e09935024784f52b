"""Registers the hot-path ops with the PyTorch dispatcher as `torch.ops.acids_b200.*`.

The nn.Module mirror of the reference calls these, which keeps every module TorchScript-scriptable
(`torch.jit.script` resolves `torch.ops.<ns>.<op>` from the registered schema).  The implementations
are the ctypes calls of ops.py — CUDA kernels behind the C ABI; there is no other backend.
"""
import os
from typing import Optional, Tuple

import torch

from . import ops

NS = "acids_b200"
SHIM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libacids_b200_torch.so")
# Two registrations of the same schemas over the same C ABI:
#   * libacids_b200_torch.so (csrc/torch_shim.cpp, TORCH_LIBRARY in C++) — the default whenever it is built: a chain saved
#     with torch.jit.save loads in any process that dlopens it, Python or not;
#   * the Python registration below — used when the shim is absent or ACIDS_B200_PY_OPS=1 (the ctypes twin, ops.py).
USE_SHIM = os.path.exists(SHIM_PATH) and os.environ.get("ACIDS_B200_PY_OPS", "0") != "1" and os.environ.get("ACIDS_B200_LIB") is None
if USE_SHIM:
    torch.ops.load_library(SHIM_PATH)
    _lib = None
else:
    _lib = torch.library.Library(NS, "DEF")

_SCHEMAS = {
    "stft_fwd": "(Tensor x, Tensor window, int n_fft, int hop, bool center) -> Tensor",
    "midside_stft_fwd": "(Tensor x, Tensor window, int n_fft, int hop, int midside) -> Tensor",
    "stft_mag_fwd": "(Tensor x, Tensor window, int n_fft, int hop, Tensor? band_meta, Tensor? band_coef, int contrast, "
                    "float eps, Tensor? offset, Tensor? scale, bool drop_first) -> Tensor",
    "stft_polar_fwd": "(Tensor x, Tensor window, int n_fft, int hop, Tensor? band_meta, Tensor? band_coef, int contrast, "
                      "float eps, Tensor? mag_offset, Tensor? mag_scale, int phase_mode, int method, bool weighted, "
                      "Tensor? ph_offset, Tensor? ph_scale, bool drop_first, int midside=0) -> Tensor",
    "mag_epilogue": "(Tensor X, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? offset, "
                    "Tensor? scale, bool drop_first) -> Tensor",
    "mag_invert": "(Tensor y, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? offset, "
                  "Tensor? scale, bool pad_last) -> Tensor",
    "melspec_fwd": "(Tensor x, Tensor window, int n_fft, int hop, Tensor band_meta, Tensor band_coef, float power, "
                   "Tensor? offset, Tensor? scale) -> Tensor",
    "mfcc_dct": "(Tensor mel, Tensor dct, float top_db) -> Tensor",
    "phase_fwd": "(Tensor X, int mode, int method, bool weighted, Tensor? offset, Tensor? scale, bool drop_first) -> Tensor",
    "phase_inv": "(Tensor y, int mode, int method, Tensor? offset, Tensor? scale, bool pad_last) -> Tensor",
    "polar_fwd": "(Tensor X, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? mag_offset, "
                 "Tensor? mag_scale, int phase_mode, int method, bool weighted, Tensor? ph_offset, Tensor? ph_scale, "
                 "bool drop_first) -> Tensor",
    "phase_inv_polar": "(Tensor y, Tensor mag, int mode, int method, Tensor? offset, Tensor? scale, bool pad_last) -> Tensor",
    "polar_to_complex": "(Tensor mag, Tensor phase) -> Tensor",
    "griffinlim_update": "(Tensor rebuilt, Tensor tprev, Tensor mag, float momentum) -> Tensor",
    "istft_ola": "(Tensor X, Tensor window, int n_fft, int hop) -> Tensor",
    "irfft_frames": "(Tensor X, Tensor window, int n_fft) -> Tensor",
    "ola_stream": "(Tensor frames, int hop, int keep, Tensor? carry_in, float gain) -> (Tensor, Tensor)",
    "mulaw_encode": "(Tensor x, int channels, int one_hot) -> Tensor",
    "mulaw_decode": "(Tensor q, int channels) -> Tensor",
    "one_hot": "(Tensor q, int n_classes) -> Tensor",
    "stats": "(Tensor x, int contrast, float eps, bool abs_contrast=False) -> Tensor",
    "stft_stats": "(Tensor x, Tensor window, int n_fft, int hop, int contrast, float eps) -> Tensor",
    "mono_mix": "(Tensor x) -> Tensor",
    "midside": "(Tensor x, bool pad_mid, bool inverse) -> Tensor",
}


def _stft_fwd(x, window, n_fft: int, hop: int, center: bool):
    return ops.stft_fwd(x, window, n_fft, hop, center)


def _midside_stft_fwd(x, window, n_fft: int, hop: int, midside: int):
    return ops.midside_stft_fwd(x, window, n_fft, hop, midside)


def _stft_mag_fwd(x, window, n_fft: int, hop: int, band_meta, band_coef, contrast: int, eps: float, offset, scale,
                  drop_first: bool):
    return ops.stft_mag_fwd(x, window, n_fft, hop, ops.as_band(band_meta, band_coef), contrast, eps, offset, scale, drop_first)


def _stft_polar_fwd(x, window, n_fft: int, hop: int, band_meta, band_coef, contrast: int, eps: float, mag_offset, mag_scale,
                    phase_mode: int, method: int, weighted: bool, ph_offset, ph_scale, drop_first: bool, midside: int = 0):
    """[MidSide ->] STFT -> Polar / PolarIF, wave -> stacked [..., T, 2, F'].  One kernel when the phase mode needs no scan
    over the frames (raw phase, forward-difference IF); otherwise the spectrum is produced once and consumed by the two
    representation kernels writing straight into their slot of the stacked tensor."""
    if ops.fusable_phase(phase_mode, method):
        return ops.stft_polar_fwd(x, window, n_fft, hop, ops.as_band(band_meta, band_coef), contrast, eps, mag_offset, mag_scale,
                                  phase_mode, method, weighted, ph_offset, ph_scale, drop_first, midside)
    X = ops.midside_stft_fwd(x, window, n_fft, hop, midside) if midside else ops.stft_fwd(x, window, n_fft, hop, True)
    return _polar_fwd(X, band_meta, band_coef, contrast, eps, mag_offset, mag_scale, phase_mode, method, weighted,
                      ph_offset, ph_scale, drop_first)


def _mag_epilogue(X, band_meta, band_coef, contrast: int, eps: float, offset, scale, drop_first: bool):
    return ops.mag_epilogue(X, ops.as_band(band_meta, band_coef), contrast, eps, offset, scale, drop_first)


def _mag_invert(y, band_meta, band_coef, contrast: int, eps: float, offset, scale, pad_last: bool):
    return ops.mag_invert(y, ops.as_band(band_meta, band_coef), contrast, eps, offset, scale, pad_last)


def _melspec_fwd(x, window, n_fft: int, hop: int, band_meta, band_coef, power: float, offset, scale):
    return ops.melspec_fwd(x, window, n_fft, hop, ops.as_band(band_meta, band_coef), power, offset, scale)


def _mfcc_dct(mel, dct, top_db: float):
    return ops.mfcc_dct(mel, dct, None if top_db < 0 else top_db)


def _phase_fwd(X, mode: int, method: int, weighted: bool, offset, scale, drop_first: bool):
    return ops.phase_fwd(X, mode, method, weighted, offset, scale, drop_first)


def _phase_inv(y, mode: int, method: int, offset, scale, pad_last: bool):
    return ops.phase_inv(y, mode, method, offset, scale, pad_last)


def _polar_fwd(X, band_meta, band_coef, contrast: int, eps: float, mag_offset, mag_scale, phase_mode: int, method: int,
               weighted: bool, ph_offset, ph_scale, drop_first: bool):
    Xd = ops._dev(X)
    band = ops.as_band(band_meta, band_coef)
    if band is None and Xd.ndim >= 2:       # no mel bank: both halves from one read of the spectrum
        return ops.polar_fwd(X, contrast, eps, mag_offset, mag_scale, phase_mode, method, weighted, ph_offset, ph_scale, drop_first)
    F = Xd.shape[-1]
    if band is not None and Xd.ndim >= 2 and ops.polar_rows_pays(phase_mode, F):
        # mel bank + raw phase: one row-tile kernel, one read of the spectrum
        return ops.polar_rows_fwd(X, band, contrast, eps, mag_offset, mag_scale, phase_mode, method, weighted, ph_offset, ph_scale, drop_first)
    n_mag = (band.n_out if band is not None else F) - int(drop_first)
    n_ph = F - int(drop_first)
    if n_mag != n_ph:
        raise RuntimeError("stack expects each tensor to be equal size, but got [%d] and [%d] bins" % (n_mag, n_ph))
    T = Xd.shape[-2]
    rows = Xd.numel() // F
    out = torch.empty((rows // T, T, 2, n_mag), dtype=torch.float32, device=Xd.device)
    ops.mag_epilogue(Xd, band, contrast, eps, mag_offset, mag_scale, drop_first, out=out, out_slot=0, out_slots=2)
    ops.phase_fwd(Xd, phase_mode, method, weighted, ph_offset, ph_scale, drop_first, out=out, out_slot=1, out_slots=2)
    out = out.reshape(tuple(Xd.shape[:-2]) + (T, 2, n_mag))
    return ops._ret(out, X)


def _phase_inv_polar(y, mag, mode: int, method: int, offset, scale, pad_last: bool):
    return ops.phase_inv_polar(y, mag, mode, method, offset, scale, pad_last)


def _polar_to_complex(mag, phase):
    return ops.polar_to_complex(mag, phase)


def _griffinlim_update(rebuilt, tprev, mag, momentum: float):
    return ops.griffinlim_update(rebuilt, tprev, mag, momentum)


def _istft_ola(X, window, n_fft: int, hop: int):
    return ops.istft_ola(X, window, n_fft, hop)


def _irfft_frames(X, window, n_fft: int):
    return ops.irfft_frames(X, window, n_fft)


def _ola_stream(frames, hop: int, keep: int, carry_in, gain: float):
    return ops.ola_stream(frames, hop, keep, carry_in, gain)


def _mulaw_encode(x, channels: int, one_hot: int):
    return ops.mulaw_encode(x, channels, one_hot)


def _mulaw_decode(q, channels: int):
    return ops.mulaw_decode(q, channels)


def _one_hot(q, n_classes: int):
    return ops.one_hot(q, n_classes)


def _stats(x, contrast: int, eps: float, abs_contrast: bool = False):
    return ops._ret(ops.stats(x, contrast, eps, abs_contrast), x)


def _stft_stats(x, window, n_fft: int, hop: int, contrast: int, eps: float):
    return ops._ret(ops.stft_stats(x, window, n_fft, hop, contrast, eps), x)


def _mono_mix(x):
    return ops.mono_mix(x)


def _midside(x, pad_mid: bool, inverse: bool):
    return ops.midside(x, pad_mid, inverse)


if _lib is not None:
    for _name, _schema in _SCHEMAS.items():
        _lib.define(_name + _schema)
        _lib.impl(_name, globals()["_" + _name], "CompositeExplicitAutograd")
