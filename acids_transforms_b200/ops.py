"""Tensor-level entry points: torch supplies device memory and the current stream, the work is done
by the CUDA kernels behind the C ABI (include/acids_b200.h).  No op here has a CPU or eager path.

Host (CPU) tensors are accepted for drop-in compatibility with the reference, whose tests feed CPU
tensors: they are copied to the current CUDA device, processed there, and the result is copied
back — the copies are part of the call, which is what `bench.py`'s `e2e` number measures.
"""
import ctypes
import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Band, CONTRAST_IDS, IF_METHOD_IDS, ONEHOT_IDS, PHASE_IF, PHASE_RAW, PHASE_UNWRAP


# ------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------
def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.AcidsError("acids_transforms_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _dev(t: torch.Tensor) -> torch.Tensor:
    """Tensor on the compute device (host tensors are staged to the current CUDA device)."""
    if t.is_cuda:
        return t
    _require_cuda()
    return t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)


def _ret(y: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return y if like.is_cuda else y.to(like.device)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _run(out: torch.Tensor, fn, *args):
    """Launch through the C ABI unless there is nothing to compute (empty batch)."""
    if out.numel() == 0:
        return
    _lib.check(fn(*args))


def _scalar(t, device) -> Optional[torch.Tensor]:
    """Normalize.offset / .scale as a 1-element float32 device tensor (or None)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.tensor(float(t), dtype=torch.float32)
    if t.numel() != 1:      # Normalize.offset starts as torch.zeros(0) before scale_data (norm.py:22)
        raise RuntimeError("normalisation buffers are not set: call scale_data() first")
    return t.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()


class BandedMatrix:
    """Column-banded ("group-ELL") view of a mostly-zero [n_in, n_out] matrix — the mel banks.

    spectral_repr.py:173-189 builds these matrices dense; 99.6 % of the square 513x513 bank is zero
    (SURVEY.md §8a A5).  Layout (include/acids_b200.h, acids_band): columns in groups of 32; group g applies
    cnt[g] = its widest band to every column, column m reading rows start[m] .. start[m]+cnt[g]-1 with zero
    padded coefficients.  meta = (cnt, base)[n_groups] ++ start[n_out] ++ (n_in, coef_len) (the trailing pair is
    host-side bookkeeping).  Device copies are made lazily per device.
    """

    def __init__(self, dense: torch.Tensor):
        m = dense.detach().to("cpu", torch.float32).numpy()
        if m.ndim == 3:
            m = m[0]
        self.n_in, self.n_out = int(m.shape[0]), int(m.shape[1])
        nz = m != 0
        first = np.where(nz.any(0), nz.argmax(0), 0)
        last = np.where(nz.any(0), self.n_in - 1 - nz[::-1].argmax(0), -1)
        width = (last - first + 1).clip(min=0)
        self.nnz_stored = int(width.sum())
        n_groups = (self.n_out + 31) // 32
        start = np.zeros(self.n_out, np.int32)
        ginfo = np.zeros((n_groups, 2), np.int32)
        blocks = []
        base = 0
        for g in range(n_groups):
            cols = np.arange(g * 32, min(self.n_out, g * 32 + 32))
            cnt = int(width[cols].max()) if cols.size else 0
            blk = np.zeros((cnt, 32), np.float32)
            for c in cols:
                s0 = int(min(first[c], self.n_in - cnt)) if cnt else 0      # keep the window inside the input
                s0 = max(s0, 0)
                start[c] = s0
                if width[c]:
                    blk[first[c] - s0:first[c] - s0 + width[c], c - g * 32] = m[first[c]:last[c] + 1, c]
            ginfo[g] = (cnt, base)
            base += cnt
            blocks.append(blk)
        coef = np.concatenate(blocks, 0).reshape(-1) if base else np.zeros(0, np.float32)
        self.coef_len = int(coef.size)
        meta = np.concatenate([ginfo.reshape(-1), start, np.array([self.n_in, self.coef_len], np.int32)])
        self._meta = torch.from_numpy(meta.astype(np.int32))
        self._coef = torch.from_numpy(coef if coef.size else np.zeros(32, np.float32)).contiguous()
        self._dev = {}

    def tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(meta int32, coef float32): what the modules keep as non-persistent buffers."""
        return self._meta.clone(), self._coef.clone()

    @classmethod
    def from_tensors(cls, meta: torch.Tensor, coef: torch.Tensor) -> "BandedMatrix":
        self = cls.__new__(cls)
        # meta = (cnt, base)[ceil(n_out / 32)] ++ start[n_out] ++ (n_in, coef_len): solve for n_out
        total = int(meta.numel()) - 2
        n_out = next(n for n in range(max(total - 2 * ((total + 31) // 32) - 2, 0), total + 1) if n + 2 * ((n + 31) // 32) == total)
        self.n_out = n_out
        # host-side bookkeeping pair at the end of meta: read ONCE per bank (as_band caches the view on the buffers'
        # identity and version), so a CUDA-resident bank costs one 8-byte device->host read at first use and none after
        self.n_in = int(meta[-2].item()) if meta.numel() >= 2 else -1
        self.coef_len = int(coef.numel())
        self.nnz_stored = self.coef_len
        self._meta, self._coef, self._dev = meta, coef, {}
        return self

    def on(self, device) -> Band:
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = (self._meta.to(device).contiguous(), self._coef.to(device).contiguous())
        meta, coef = self._dev[key]
        return Band(meta.data_ptr(), coef.data_ptr(), self.n_out, self.coef_len, self.n_in)


_NO_BAND = Band(None, None, 0, 0, 0)
_BAND_CACHE = {}


def as_band(meta: Optional[torch.Tensor], coef: Optional[torch.Tensor]) -> Optional[BandedMatrix]:
    """BandedMatrix view of a (meta, coef) buffer pair, cached on the buffers' identity/version.  An entry keeps its two
    tensors alive: a freed bank's addresses could otherwise be handed to a different bank with the same key (the metadata of
    a 513 -> 128 and of a 1025 -> 128 bank have the same size)."""
    if meta is None or coef is None:
        return None
    key = (meta.data_ptr(), coef.data_ptr(), meta._version, coef._version, meta.numel(), coef.numel())
    hit = _BAND_CACHE.get(key)
    if hit is None:
        if len(_BAND_CACHE) > 64:
            _BAND_CACHE.clear()
        hit = _BAND_CACHE[key] = (BandedMatrix.from_tensors(meta, coef), meta, coef)
    return hit[0]


def _band(b: Optional[BandedMatrix], device) -> Band:
    return b.on(device) if b is not None else _NO_BAND


def _cid(contrast) -> int:
    return contrast if isinstance(contrast, int) else CONTRAST_IDS[contrast]


def _mid(method) -> int:
    return method if isinstance(method, int) else IF_METHOD_IDS[method]


def _flat_batch(x: torch.Tensor, event_dims: int):
    """reshape_batches (utils/misc.py:168-178): flatten leading dims, keep the last `event_dims`."""
    batch = x.shape[:x.ndim - event_dims]
    return x.reshape((-1,) + tuple(x.shape[x.ndim - event_dims:])).contiguous(), batch


def n_frames_centered(L: int, hop: int) -> int:
    return 1 + L // hop


# ------------------------------------------------------------------------------------------------
# (1) STFT forward
# ------------------------------------------------------------------------------------------------
def _check_stft_input(x, n_fft, center=True):
    if x.dtype != torch.float32:
        raise RuntimeError("acids_b200: expected a float32 waveform, got %s" % x.dtype)
    if center and not (0 < n_fft // 2 < x.shape[-1]):
        # same condition torch's reflect padding enforces for torch.stft(center=True)
        raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension, "
                           "but got: padding (%d, %d) at dimension 2 of input %s" % (n_fft // 2, n_fft // 2, list(x.shape)))


def stft_fwd(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int, center: bool = True) -> torch.Tensor:
    """x [..., L] float32 -> complex64 [..., T, F]  (stft.py:97-104, dgt.py:63-70).
    center=False: x is pre-framed [..., n, n_fft] -> [..., n, F]  (stft.py:248-253)."""
    lib = _lib.load()
    xd = _dev(x)
    if center:
        _check_stft_input(xd, n_fft)
        xf, batch = _flat_batch(xd, 1)
        B, L = xf.shape
        T = n_frames_centered(L, hop)
        hop_k, ldx = hop, L
    else:
        if xd.shape[-1] != n_fft:
            raise RuntimeError("acids_b200: pre-framed input must have n_fft=%d samples per frame, got %d" % (n_fft, xd.shape[-1]))
        if xd.dtype != torch.float32:
            raise RuntimeError("acids_b200: expected float32 frames")
        xf2, batch = _flat_batch(xd, 1)              # [rows, n_fft]
        T = xf2.shape[0]
        xf = xf2.reshape(1, -1)
        B, L, hop_k, ldx = 1, xf.shape[1], n_fft, xf.shape[1]
    F = n_fft // 2 + 1
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((B, T, F), dtype=torch.complex64, device=xf.device)
    with torch.cuda.device(xf.device):
        _run(out, lib.acids_stft_fwd, _ptr(xf), B, L, ldx, _ptr(w), n_fft, hop_k, int(center), T, _ptr(out), _stream(xf.device))
    out = out.reshape(tuple(batch) + ((T, F) if center else (F,)))
    return _ret(out, x)


def midside_stft_fwd(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int, midside: int) -> torch.Tensor:
    """MidSide.forward -> STFT.forward in one kernel: stereo x [..., 2, L] -> complex64 [..., 2, T, F] (raw.py:145-161,
    stft.py:101-102); midside = 1 (pad_mid=False) or 2 (pad_mid=True).  The mid/side waveform is never written."""
    lib = _lib.load()
    xd = _dev(x)
    _check_stft_input(xd, n_fft)
    if xd.ndim < 2 or xd.shape[-2] != 2:
        raise RuntimeError("acids_b200: the fused MidSide prologue needs a stereo [..., 2, L] input")
    xf, batch = _flat_batch(xd, 1)
    B, L = xf.shape
    T, F = n_frames_centered(L, hop), n_fft // 2 + 1
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((B, T, F), dtype=torch.complex64, device=xf.device)
    with torch.cuda.device(xf.device):
        _run(out, lib.acids_midside_stft_fwd, _ptr(xf), B, L, _ptr(w), n_fft, hop, T, int(midside), _ptr(out), _stream(xf.device))
    return _ret(out.reshape(tuple(batch) + (T, F)), x)


def stft_mag_fwd(x, window, n_fft, hop, band: Optional[BandedMatrix], contrast, eps, offset, scale,
                 drop_first: bool = False, out: Optional[torch.Tensor] = None, out_slot: int = 0, out_slots: int = 1):
    """Fused STFT + Magnitude.forward: x [..., L] -> float32 [..., T, n_cols - drop]  (stft.py:101 + spectral_repr.py:215-226).
    With out_slots = 2 the rows are written into slot `out_slot` of a stacked [..., T, 2, n] tensor (Polar*)."""
    lib = _lib.load()
    xd = _dev(x)
    _check_stft_input(xd, n_fft)
    xf, batch = _flat_batch(xd, 1)
    B, L = xf.shape
    T = n_frames_centered(L, hop)
    if band is not None and band.n_in != n_fft // 2 + 1:
        # the reference's matmul raises the same way when Magnitude.n_fft disagrees with the STFT's
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (B * T, n_fft // 2 + 1, band.n_in, band.n_out))
    n_cols = band.n_out if band is not None else n_fft // 2 + 1
    n_keep = n_cols - int(drop_first)
    dev = xf.device
    w = _dev(window).to(torch.float32).contiguous()
    own = out is None
    if own:
        out = torch.empty((B, T, out_slots, n_keep) if out_slots > 1 else (B, T, n_keep), dtype=torch.float32, device=dev)
    row_stride = out_slots * n_keep
    base = out.view(-1)[out_slot * n_keep:] if out_slots > 1 else out
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stft_mag_fwd, _ptr(xf), B, L, L, _ptr(w), n_fft, hop, 1, T, _band(band, dev), _cid(contrast),
                                          float(eps), _ptr(off), _ptr(sc), int(drop_first), _ptr(base), T * row_stride,
                                          row_stride, _stream(dev))
    if not own:
        return out
    out = out.reshape(tuple(batch) + tuple(out.shape[1:]))
    return _ret(out, x)


def stft_polar_fwd(x, window, n_fft, hop, band: Optional[BandedMatrix], contrast, eps, mag_offset, mag_scale, phase_mode,
                   method, weighted, ph_offset, ph_scale, drop_first: bool = False, midside: int = 0):
    """Fused [MidSide ->] STFT -> Polar / PolarIF: x [..., L] (midside: [..., 2, L]) -> float32 stacked [..., T, 2, F']
    (raw.py:145-162, stft.py:101-102, spectral_repr.py:431-440); one kernel, the spectrum never reaches HBM.
    Fused phase modes: raw phase and forward-difference IF (`fusable_phase`)."""
    lib = _lib.load()
    xd = _dev(x)
    _check_stft_input(xd, n_fft)
    if midside and (xd.ndim < 2 or xd.shape[-2] != 2):
        raise RuntimeError("acids_b200: the fused MidSide prologue needs a stereo [..., 2, L] input")
    xf, batch = _flat_batch(xd, 1)
    B, L = xf.shape
    T = n_frames_centered(L, hop)
    F = n_fft // 2 + 1
    if band is not None and band.n_in != F:
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (B * T, F, band.n_in, band.n_out))
    n_mag = (band.n_out if band is not None else F) - int(drop_first)
    n_keep = F - int(drop_first)
    if n_mag != n_keep:
        raise RuntimeError("stack expects each tensor to be equal size, but got [%d] and [%d] bins" % (n_mag, n_keep))
    dev = xf.device
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((B, T, 2, n_keep), dtype=torch.float32, device=dev)
    mo, ms = _scalar(mag_offset, dev), _scalar(mag_scale, dev)
    po, ps = _scalar(ph_offset, dev), _scalar(ph_scale, dev)
    flat = out.view(-1)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stft_polar_fwd, _ptr(xf), B, L, L, _ptr(w), n_fft, hop, T, int(midside), _band(band, dev),
                                            _cid(contrast), float(eps), _ptr(mo), _ptr(ms), int(phase_mode), _mid(method),
                                            int(bool(weighted)), _ptr(po), _ptr(ps), int(drop_first),
                                            _ptr(flat), T * 2 * n_keep, 2 * n_keep, _ptr(flat[n_keep:]), T * 2 * n_keep, 2 * n_keep,
                                            _stream(dev))
    return _ret(out.reshape(tuple(batch) + (T, 2, n_keep)), x)


def fusable_phase(phase_mode: int, method: int) -> bool:
    """Phase modes acids_stft_polar_fwd evaluates without a scan over the frames."""
    return phase_mode == PHASE_RAW or (phase_mode == PHASE_IF and method == IF_METHOD_IDS["forward"])


# ------------------------------------------------------------------------------------------------
# (2) Magnitude on a spectrum
# ------------------------------------------------------------------------------------------------
def _as_complex64(X):
    if not torch.is_complex(X):
        X = X.to(torch.float32).to(torch.complex64)     # abs() of a real tensor is still defined in the reference
    return X.to(torch.complex64)


def mag_epilogue(X, band: Optional[BandedMatrix], contrast, eps, offset, scale, drop_first=False,
                 out: Optional[torch.Tensor] = None, out_slot: int = 0, out_slots: int = 1):
    """Magnitude.forward: X [..., F] complex -> float32 [..., n_cols - drop]  (spectral_repr.py:215-226)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    Xf, batch = _flat_batch(Xd, 1)
    rows, F = Xf.shape
    if band is not None and band.n_in != F:
        # the reference's matmul raises the same way when Magnitude.n_fft disagrees with the STFT's
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (rows, F, band.n_in, band.n_out))
    n_cols = band.n_out if band is not None else F
    n_keep = n_cols - int(drop_first)
    dev = Xf.device
    own = out is None
    if own:
        out = torch.empty((rows, out_slots, n_keep) if out_slots > 1 else (rows, n_keep), dtype=torch.float32, device=dev)
    row_stride = out_slots * n_keep
    base = out.view(-1)[out_slot * n_keep:] if out_slots > 1 else out
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_mag_epilogue, _ptr(Xf), rows, F, _band(band, dev), _cid(contrast), float(eps), _ptr(off),
                                          _ptr(sc), int(drop_first), _ptr(base), row_stride, _stream(dev))
    if not own:
        return out
    out = out.reshape(tuple(batch) + tuple(out.shape[1:]))
    return _ret(out, X)


def mag_invert(y, inverse_band: Optional[BandedMatrix], contrast, eps, offset, scale, pad_last=False):
    """Magnitude.invert: y [..., n_in] -> float32 [..., n_out]  (spectral_repr.py:228-240)."""
    lib = _lib.load()
    yd = _dev(y).to(torch.float32)
    yf, batch = _flat_batch(yd, 1) if yd.is_contiguous() else (yd.reshape(-1, yd.shape[-1]), yd.shape[:-1])
    if yf.stride(-1) != 1:
        yf = yf.contiguous()
    rows, n_in = yf.shape
    n_val = n_in + int(pad_last)
    if inverse_band is not None and inverse_band.n_in != n_val:
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (rows, n_val, inverse_band.n_in, inverse_band.n_out))
    n_out = inverse_band.n_out if inverse_band is not None else n_val
    dev = yf.device
    out = torch.empty((rows, n_out), dtype=torch.float32, device=dev)
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_mag_invert, _ptr(yf), rows, n_in, yf.stride(0), int(pad_last), _band(inverse_band, dev),
                                        _cid(contrast), float(eps), _ptr(off), _ptr(sc), _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (n_out,)), y)


def melspec_fwd(x, window, n_fft, hop, mel: BandedMatrix, power=2.0, offset=None, scale=None):
    """MFCC.forward (= torchaudio MelSpectrogram): x [..., L] -> [..., n_mels, T]  (mel.py:68-73)."""
    lib = _lib.load()
    xd = _dev(x)
    _check_stft_input(xd, n_fft)
    xf, batch = _flat_batch(xd, 1)
    B, L = xf.shape
    T = n_frames_centered(L, hop)
    dev = xf.device
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((B, mel.n_out, T), dtype=torch.float32, device=dev)
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_melspec_fwd, _ptr(xf), B, L, L, _ptr(w), n_fft, hop, T, mel.on(dev), float(power), _ptr(off), _ptr(sc),
                                         _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (mel.n_out, T)), x)


def mfcc_dct(mel, dct, top_db: Optional[float] = 80.0, tensor_cores: Optional[bool] = None):
    """dB + top_db floor + DCT-II: mel [..., n_mels, T] -> [..., n_mfcc, T]  (torchaudio MFCC, _transforms.py:701-718).
    tensor_cores: the DCT as a tcgen05 3xTF32 GEMM (needs n_mels % 8 == 0, n_mels <= 128, n_mfcc <= 48); None picks it
    whenever the shape fits, False forces the FP32 register-tiled kernel."""
    lib = _lib.load()
    md = _dev(mel).to(torch.float32)
    mf, batch = _flat_batch(md, 2)
    B, n_mels, T = mf.shape
    # torchaudio's amplitude_to_DB packs dim -3 as "channels": inputs with <= 3 dims share ONE max
    group = B if md.ndim <= 3 else int(md.shape[-3])
    dev = mf.device
    d = _dev(dct).to(torch.float32).contiguous()
    n_mfcc = d.shape[1]
    out = torch.empty((B, n_mfcc, T), dtype=torch.float32, device=dev)
    gmax = torch.empty((max(B // max(group, 1), 1),), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if tensor_cores is None:
            tensor_cores = n_mels % 8 == 0 and n_mels <= 128 and n_mfcc <= 48
        fn = lib.acids_mfcc_dct_tc if tensor_cores else lib.acids_mfcc_dct
        _run(out, fn, _ptr(mf), B, n_mels, T, _ptr(d), n_mfcc, float(-1.0 if top_db is None else top_db),
                      max(group, 1), _ptr(gmax), _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (n_mfcc, T)), mel)


# ------------------------------------------------------------------------------------------------
# (3) phase / IF
# ------------------------------------------------------------------------------------------------
def phase_fwd(X, mode: int, method="forward", weighted=False, offset=None, scale=None, drop_first=False,
              out: Optional[torch.Tensor] = None, out_slot: int = 0, out_slots: int = 1):
    """Phase.forward / IF.forward on X [..., T, F]  (spectral_repr.py:270-278, :319-357; utils/misc.py:12-26)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    if Xd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] spectrum)")
    Xf, batch = _flat_batch(Xd, 2)
    B, T, F = Xf.shape
    n_keep = F - int(drop_first)
    dev = Xf.device
    own = out is None
    if own:
        out = torch.empty((B, T, out_slots, n_keep) if out_slots > 1 else (B, T, n_keep), dtype=torch.float32, device=dev)
    row_stride = out_slots * n_keep
    base = out.view(-1)[out_slot * n_keep:] if out_slots > 1 else out
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_phase_fwd, _ptr(Xf), B, T, F, mode, _mid(method), int(bool(weighted)), _ptr(off), _ptr(sc),
                                       int(drop_first), _ptr(base), T * row_stride, row_stride, _stream(dev))
    if not own:
        return out
    out = out.reshape(tuple(batch) + tuple(out.shape[1:]))
    return _ret(out, X)


def polar_fwd(X, contrast, eps, mag_offset, mag_scale, phase_mode: int, method="forward", weighted=False, ph_offset=None,
              ph_scale=None, drop_first=False):
    """Band-less Magnitude.forward and Phase/IF.forward of X [..., T, F] from one read of the spectrum, stacked:
    -> float32 [..., T, 2, F - drop]  (SpectralRepresentation.forward with stack=-2, spectral_repr.py:434-440)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    if Xd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] spectrum)")
    Xf, batch = _flat_batch(Xd, 2)
    B, T, F = Xf.shape
    n_keep = F - int(drop_first)
    dev = Xf.device
    out = torch.empty((B, T, 2, n_keep), dtype=torch.float32, device=dev)
    flat = out.view(-1)
    mo, ms = _scalar(mag_offset, dev), _scalar(mag_scale, dev)
    po, ps = _scalar(ph_offset, dev), _scalar(ph_scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_polar_fwd, _ptr(Xf), B, T, F, _cid(contrast), float(eps), _ptr(mo), _ptr(ms), phase_mode, _mid(method),
                                       int(bool(weighted)), _ptr(po), _ptr(ps), int(drop_first),
                                       _ptr(flat), T * 2 * n_keep, 2 * n_keep, _ptr(flat[n_keep:]), T * 2 * n_keep, 2 * n_keep,
                                       _stream(dev))
    return _ret(out.reshape(tuple(batch) + (T, 2, n_keep)), X)


POLAR_ROWS_MAX_BINS = 4352


def polar_rows_pays(phase_mode: int, n_bins: int) -> bool:
    """Where acids_polar_rows_fwd beats acids_mag_epilogue + acids_phase_fwd on B200 (tools/polar_rows_probe.py):
    the raw phase on rows of more than 544 bins (1.83 vs 2.04 ms at n_fft 1024 ... 0.46 vs 0.48 ms at 8192).  Both halves are
    issue bound, not memory bound, so the forward-difference IF — an extra pass over the tile — is 3-13 % SLOWER fused
    (0.96 vs 0.89 ms at n_fft 4096) and keeps the two kernels; the entry point supports it and the tests cover it."""
    return phase_mode == PHASE_RAW and 544 < n_bins <= POLAR_ROWS_MAX_BINS


def polar_rows_fwd(X, band: Optional[BandedMatrix], contrast, eps, mag_offset, mag_scale, phase_mode: int, method="forward",
                   weighted=False, ph_offset=None, ph_scale=None, drop_first=False):
    """Magnitude.forward (with its mel bank) and Phase / forward-difference IF of X [..., T, F] from ONE read of the spectrum,
    stacked -> float32 [..., T, 2, F - drop]  (spectral_repr.py:431-440); raw phase or IF `forward` only (`fusable_phase`)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    if Xd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] spectrum)")
    Xf, batch = _flat_batch(Xd, 2)
    B, T, F = Xf.shape
    if band is not None and band.n_in != F:
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)" % (B * T, F, band.n_in, band.n_out))
    n_keep = F - int(drop_first)
    n_mag = (band.n_out if band is not None else F) - int(drop_first)
    if n_mag != n_keep:
        raise RuntimeError("stack expects each tensor to be equal size, but got [%d] and [%d] bins" % (n_mag, n_keep))
    dev = Xf.device
    out = torch.empty((B, T, 2, n_keep), dtype=torch.float32, device=dev)
    flat = out.view(-1)
    mo, ms = _scalar(mag_offset, dev), _scalar(mag_scale, dev)
    po, ps = _scalar(ph_offset, dev), _scalar(ph_scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_polar_rows_fwd, _ptr(Xf), B, T, F, _band(band, dev), _cid(contrast), float(eps), _ptr(mo), _ptr(ms),
             int(phase_mode), _mid(method), int(bool(weighted)), _ptr(po), _ptr(ps), int(drop_first), _ptr(flat), 2 * n_keep,
             _ptr(flat[n_keep:]), 2 * n_keep, _stream(dev))
    return _ret(out.reshape(tuple(batch) + (T, 2, n_keep)), X)


def phase_inv(y, mode: int, method="forward", offset=None, scale=None, pad_last=False):
    """Phase.invert / IF.invert: y [..., T, n_in] -> phase [..., T, n_in + pad]  (spectral_repr.py:46-53, :359-375)."""
    lib = _lib.load()
    yd = _dev(y).to(torch.float32)
    if yd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] tensor)")
    yf = yd.reshape((-1,) + tuple(yd.shape[-2:]))
    if yf.stride(-1) != 1 or (yf.shape[0] > 1 and yf.stride(0) < yf.shape[1] * yf.stride(1)):
        yf = yf.contiguous()
    B, T, n_in = yf.shape
    dev = yf.device
    out = torch.empty((B, T, n_in + int(pad_last)), dtype=torch.float32, device=dev)
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_phase_inv, _ptr(yf), B, T, n_in, yf.stride(0) if B > 1 else T * yf.stride(1), yf.stride(1),
                                       int(pad_last), mode, _mid(method), _ptr(off), _ptr(sc), _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(yd.shape[:-1]) + (n_in + int(pad_last),)), y)


def phase_inv_polar(y, mag, mode: int, method="forward", offset=None, scale=None, pad_last=False):
    """polar_to_complex(mag, phase_inv(y, ...)) without the phase round trip through memory
    (SpectralRepresentation.invert, spectral_repr.py:447-452).  IF method `central` takes the two-kernel route."""
    if mode == 2 and _mid(method) == 2:
        return polar_to_complex(mag, phase_inv(y, mode, method, offset, scale, pad_last))
    lib = _lib.load()
    yd = _dev(y).to(torch.float32)
    if yd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] tensor)")
    yf = yd.reshape((-1,) + tuple(yd.shape[-2:]))
    if yf.stride(-1) != 1 or (yf.shape[0] > 1 and yf.stride(0) < yf.shape[1] * yf.stride(1)):
        yf = yf.contiguous()
    B, T, n_in = yf.shape
    dev = yf.device
    shape = tuple(yd.shape[:-1]) + (n_in + int(pad_last),)
    md = _dev(mag).to(device=dev, dtype=torch.float32)
    if tuple(md.shape) != shape:
        md = md.expand(shape)          # raises like the reference's broadcast would
    md = md.contiguous()
    out = torch.empty(shape, dtype=torch.complex64, device=dev)
    off, sc = _scalar(offset, dev), _scalar(scale, dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_phase_inv_polar, _ptr(yf), B, T, n_in, yf.stride(0) if B > 1 else T * yf.stride(1), yf.stride(1),
                                             int(pad_last), mode, _mid(method), _ptr(off), _ptr(sc), _ptr(md), _ptr(out), _stream(dev))
    return _ret(out, y)


def polar_to_complex(mag, phase):
    """mag * exp(i phase) -> complex64  (spectral_repr.py:452)."""
    lib = _lib.load()
    md = _dev(mag).to(torch.float32).contiguous()
    pd = _dev(phase).to(torch.float32).contiguous()
    if md.shape != pd.shape:
        md, pd = torch.broadcast_tensors(md, pd)
        md, pd = md.contiguous(), pd.contiguous()
    out = torch.empty(md.shape, dtype=torch.complex64, device=md.device)
    with torch.cuda.device(md.device):
        _run(out, lib.acids_polar_to_complex, _ptr(md), _ptr(pd), md.numel(), _ptr(out), _stream(md.device))
    return _ret(out, mag)


def griffinlim_update(rebuilt, tprev, mag, momentum: float, out: Optional[torch.Tensor] = None):
    """One fast-Griffin-Lim update: a = rebuilt - momentum/(1+momentum) * tprev; mag * a / (|a| + 1e-16).
    complex64 in / out (torchaudio functional.py:336-350 as one kernel)."""
    lib = _lib.load()
    rd = _as_complex64(_dev(rebuilt)).resolve_conj().contiguous()
    td = _as_complex64(_dev(tprev)).resolve_conj().contiguous()
    md = _dev(mag).to(torch.float32).contiguous()
    if rd.shape != td.shape or rd.shape != md.shape:
        raise RuntimeError("griffinlim_update: rebuilt %s, tprev %s and mag %s must have the same shape" %
                           (tuple(rd.shape), tuple(td.shape), tuple(md.shape)))
    n = rd.numel()
    if n & 1:        # odd bin count (n_fft/2+1 bins x odd frames x odd batch): process a padded flat copy
        pad = lambda t: torch.cat([t.reshape(-1), t.new_zeros(1)])
        return griffinlim_update(pad(rd), pad(td), pad(md), momentum)[:n].reshape(rd.shape)
    if out is None:
        out = torch.empty(rd.shape, dtype=torch.complex64, device=rd.device)
    with torch.cuda.device(rd.device):
        _run(out, lib.acids_griffinlim_update, _ptr(rd), _ptr(td), _ptr(md), ctypes.c_float(momentum), n, _ptr(out),
             _stream(rd.device))
    return _ret(out, rebuilt)


# ------------------------------------------------------------------------------------------------
# (4) inverse
# ------------------------------------------------------------------------------------------------
_ENVELOPE_CACHE = {}


def istft_envelope_ok(window: torch.Tensor, n_fft: int, hop: int, n_frames: int) -> bool:
    """torch.istft's `window overlap add min` check (_refs/__init__.py:3794-3797), evaluated on the host.

    It depends only on the window, so the reference's RuntimeError can be raised without looking at the
    data.  The envelope is periodic away from the edges, hence a clip of at most 2*ceil(N/hop)+2
    frames has the same minimum over its trimmed core as the full-length one.  The verdict is cached on
    the window buffer's identity and version: a CUDA-resident window is read back ONCE (64 KB, at the
    first invert after construction / set_params / load_state_dict), never again — `invert` can be
    captured in a CUDA graph after one warm-up call and does not synchronise the host.
    """
    t_eff = min(int(n_frames), 2 * ((n_fft + hop - 1) // hop) + 2)
    key = (window.data_ptr(), window._version, str(window.device), int(n_fft), int(hop), t_eff)
    hit = _ENVELOPE_CACHE.get(key)
    if hit is not None:
        return hit[0]
    w2 = window.detach().to("cpu", torch.float64)[:n_fft].numpy() ** 2
    length = n_fft + hop * (t_eff - 1)
    env = np.zeros(length, np.float64)
    for t in range(t_eff):
        env[t * hop:t * hop + n_fft] += w2
    core = env[n_fft // 2:length - n_fft // 2]
    ok = core.size == 0 or bool(np.abs(core).min() > 1e-11)
    if len(_ENVELOPE_CACHE) > 256:
        _ENVELOPE_CACHE.clear()
    _ENVELOPE_CACHE[key] = (ok, window)      # the entry keeps the buffer alive: its address cannot be handed to another window
    return ok


def istft_ola(X, window, n_fft, hop, check_envelope: bool = True):
    """STFT.invert / DGT.invert complex branch: X [..., T, F] -> [..., hop (T-1)]  (stft.py:119-128, dgt.py:85-93)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    Xf, batch = _flat_batch(Xd, 2)
    B, T, F = Xf.shape
    if F != n_fft // 2 + 1:
        raise RuntimeError("istft: expected %d frequency bins for n_fft=%d, got %d" % (n_fft // 2 + 1, n_fft, F))
    if check_envelope and not istft_envelope_ok(window, n_fft, hop, T):
        raise RuntimeError("istft(CUDA): window overlap add min: 1")
    dev = Xf.device
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((B, hop * (T - 1)), dtype=torch.float32, device=dev)
    ws_bytes = int(lib.acids_istft_workspace_bytes(B, T, n_fft, hop))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
    with torch.cuda.device(dev):
        _run(out, lib.acids_istft_ola, _ptr(Xf), B, T, n_fft, hop, _ptr(w), _ptr(out), _ptr(ws), ws_bytes, _stream(dev))
    return _ret(out.reshape(tuple(batch) + (hop * (T - 1),)), X)


def irfft_frames(X, window, n_fft):
    """RealtimeSTFT.invert complex branch: irfft(X) * window, X [..., F] -> [..., n_fft]  (stft.py:259-266)."""
    lib = _lib.load()
    Xd = _as_complex64(_dev(X)).resolve_conj()
    Xf, batch = _flat_batch(Xd, 1)
    rows, F = Xf.shape
    if F != n_fft // 2 + 1:
        raise RuntimeError("irfft: expected %d frequency bins for n_fft=%d, got %d" % (n_fft // 2 + 1, n_fft, F))
    dev = Xf.device
    w = _dev(window).to(torch.float32).contiguous()
    out = torch.empty((rows, n_fft), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_irfft_frames, _ptr(Xf), rows, n_fft, _ptr(w), _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (n_fft,)), X)


def ola_stream(frames, hop, keep, carry_in, gain) -> Tuple[torch.Tensor, torch.Tensor]:
    """OverlapAdd.invert: frames [..., n, N] (+ carry [..., keep]) -> (out [..., (n-1) hop + N - keep], carry_out)  (oadd.py:91-104)."""
    lib = _lib.load()
    fd = _dev(frames).to(torch.float32)
    ff, batch = _flat_batch(fd, 2)
    B, n, N = ff.shape
    dev = ff.device
    total = (n - 1) * hop + N
    out = torch.empty((B, total - keep), dtype=torch.float32, device=dev)
    carry_out = torch.empty((B, keep), dtype=torch.float32, device=dev)
    ci = None
    if carry_in is not None:
        ci = _dev(carry_in).to(torch.float32).reshape(B, keep).contiguous()
    g = float(gain)
    with torch.cuda.device(dev):
        _run(out, lib.acids_ola_stream, _ptr(ff), B, n, N, hop, keep, _ptr(ci), g, _ptr(out), _ptr(carry_out), _stream(dev))
    return (_ret(out.reshape(tuple(batch) + (total - keep,)), frames),
            _ret(carry_out.reshape(tuple(batch) + (keep,)), frames))


# ------------------------------------------------------------------------------------------------
# (4b) the block-by-block streaming step as one kernel (stream.cu)
# ------------------------------------------------------------------------------------------------
def _stream_state(state: torch.Tensor, B: int, keep: int, what: str) -> torch.Tensor:
    if not (state.is_cuda and state.dtype == torch.float32 and state.is_contiguous() and state.numel() == B * keep):
        raise RuntimeError("acids_b200: stream step: `%s` must be a contiguous float32 CUDA tensor of %d x %d samples "
                           "(it is advanced in place)" % (what, B, keep))
    return state


def stream_analysis(x, window, n_fft, hop, tail, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OverlapAdd.forward + RealtimeSTFT/DGT.forward on one block (oadd.py:70-74, stft.py:248-253): x [..., n hop] ->
    complex64 [..., n, n_fft/2+1]; `tail` [..., n_fft - hop] (OverlapAdd.input_buffer) is advanced IN PLACE."""
    lib = _lib.load()
    xd = _dev(x).to(torch.float32)
    xf, batch = _flat_batch(xd, 1)
    B, C = xf.shape
    if C % hop:
        raise RuntimeError("acids_b200: stream step: the block length %d is not a multiple of hop=%d" % (C, hop))
    n, F = C // hop, n_fft // 2 + 1
    dev = xf.device
    _stream_state(tail, B, n_fft - hop, "tail")
    w = window.to(dev, torch.float32).contiguous()
    if out is None:
        out = torch.empty((B, n, F, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stream_analysis, _ptr(xf), B, n, n_fft, hop, _ptr(w), _ptr(tail), _ptr(out), _stream(dev))
    return torch.view_as_complex(out.reshape(tuple(batch) + (n, F, 2)))


def stream_synthesis(X, inv_window, n_fft, hop, gain, carry, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RealtimeSTFT/DGT.invert + OverlapAdd.invert on one block (stft.py:259-266, oadd.py:91-104): complex64
    [..., n, n_fft/2+1] -> [..., n hop]; `carry` [..., n_fft - hop] (OverlapAdd.output_buffer) is advanced IN PLACE."""
    lib = _lib.load()
    Xd = _dev(X)
    Xr = torch.view_as_real(Xd.to(torch.complex64)).contiguous()
    Xf, batch = _flat_batch(Xr, 3)
    B, n, F, _ = Xf.shape
    if F != n_fft // 2 + 1:
        raise RuntimeError("acids_b200: stream step: %d bins do not belong to n_fft=%d" % (F, n_fft))
    dev = Xf.device
    _stream_state(carry, B, n_fft - hop, "carry")
    w = inv_window.to(dev, torch.float32).contiguous()
    if out is None:
        out = torch.empty((B, n * hop), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stream_synthesis, _ptr(Xf), B, n, n_fft, hop, _ptr(w), float(gain), _ptr(carry), _ptr(out), _stream(dev))
    return out.reshape(tuple(batch) + (n * hop,))


def stream_roundtrip(x, window, inv_window, n_fft, hop, gain, tail, carry, out: Optional[torch.Tensor] = None,
                     spectrum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """stream_analysis followed by stream_synthesis in ONE kernel, the spectrum staying in registers (optionally also
    written to `spectrum`, float32 [B, n, F, 2])."""
    lib = _lib.load()
    xd = _dev(x).to(torch.float32)
    xf, batch = _flat_batch(xd, 1)
    B, C = xf.shape
    if C % hop:
        raise RuntimeError("acids_b200: stream step: the block length %d is not a multiple of hop=%d" % (C, hop))
    n = C // hop
    dev = xf.device
    _stream_state(tail, B, n_fft - hop, "tail")
    _stream_state(carry, B, n_fft - hop, "carry")
    w = window.to(dev, torch.float32).contiguous()
    iw = inv_window.to(dev, torch.float32).contiguous()
    if out is None:
        out = torch.empty((B, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stream_roundtrip, _ptr(xf), B, n, n_fft, hop, _ptr(w), _ptr(iw), float(gain), _ptr(tail), _ptr(carry),
             _ptr(spectrum), _ptr(out), _stream(dev))
    return out.reshape(tuple(batch) + (C,))


# ------------------------------------------------------------------------------------------------
# (4c) phase-gradient heap integration (pghi.cu)
# ------------------------------------------------------------------------------------------------
def pghi(mag, gamma: float, n_fft: int, hop: int, tol: float, eps: float) -> torch.Tensor:
    """DGT.pghi (dgt.py:156-236) for a batch: mag [..., T, F] -> phase [..., T, F]; one CTA per clip."""
    lib = _lib.load()
    md = _dev(mag).to(torch.float32)
    if md.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] magnitude)")
    mf, batch = _flat_batch(md, 2)
    B, T, F = mf.shape
    dev = mf.device
    out = torch.empty((B, T, F), dtype=torch.float32, device=dev)
    nbytes = int(lib.acids_pghi_workspace_bytes(B, T, F))
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_pghi, _ptr(mf), B, T, F, float(gamma), int(n_fft), int(hop), float(tol), float(eps), _ptr(ws), nbytes,
             _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (T, F)), mag)


def rt_pghi(mag, hist_mag, hist_phase, gamma: float, n_fft: int, hop: int, tol: float, eps: float, noise=None) -> torch.Tensor:
    """RealtimeDGT.pghi (dgt.py:338-452): mag [..., n, F] new frames, hist_mag [..., 2, F], hist_phase [..., F] ->
    phase [..., n, F]; one CTA per stream, heap in shared memory.  `noise` [..., n, F]: the phase of the bins below the
    tolerance (None: zeros)."""
    lib = _lib.load()
    md = _dev(mag).to(torch.float32)
    if md.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] magnitude)")
    mf, batch = _flat_batch(md, 2)
    B, n, F = mf.shape
    dev = mf.device
    hm = hist_mag.to(dev, torch.float32).reshape(B, 2, F).contiguous()
    hp = hist_phase.to(dev, torch.float32).reshape(B, F).contiguous()
    nz = None if noise is None else noise.to(dev, torch.float32).reshape(B, n, F).contiguous()
    out = torch.empty((B, n, F), dtype=torch.float32, device=dev)
    nbytes = int(lib.acids_rt_pghi_workspace_bytes(B, n, F))
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_rt_pghi, _ptr(mf), _ptr(hm), _ptr(hp), _ptr(nz) if nz is not None else None, B, n, F, float(gamma),
             int(n_fft), int(hop), float(tol), float(eps), _ptr(ws), nbytes, _ptr(out), _stream(dev))
    return _ret(out.reshape(tuple(batch) + (n, F)), mag)


def mel_tc(spec, bank) -> torch.Tensor:
    """Dense mel projection on the tensor cores (3xTF32, csrc/mel_tc.cu): spec [..., T, F] (|X| or |X|^2) x bank [F, n_mels]
    (n_mels <= 128) -> [..., n_mels, T], the layout of torchaudio's MelSpectrogram (mel.py:38-44)."""
    lib = _lib.load()
    sd = _dev(spec).to(torch.float32)
    if sd.ndim < 2:
        raise IndexError("Dimension out of range (expected a [..., frames, bins] spectrum)")
    sf, batch = _flat_batch(sd, 2)
    B, T, F = sf.shape
    bk = bank.to(sf.device, torch.float32).contiguous()
    if bk.ndim != 2 or bk.shape[0] != F:
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %s)" % (B * T, F, "x".join(str(v) for v in bk.shape)))
    n_mels = int(bk.shape[1])
    out = torch.empty((B, n_mels, T), dtype=torch.float32, device=sf.device)
    nbytes = int(lib.acids_mel_tc_workspace_bytes(F))
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=sf.device)
    with torch.cuda.device(sf.device):
        _run(out, lib.acids_mel_tc, _ptr(sf), B, T, F, _ptr(bk), n_mels, _ptr(ws), nbytes, _ptr(out), _stream(sf.device))
    return _ret(out.reshape(tuple(batch) + (n_mels, T)), spec)


# ------------------------------------------------------------------------------------------------
# (5) mu-law / one-hot
# ------------------------------------------------------------------------------------------------
def _log1p_mu(channels: int) -> float:
    # the reference evaluates torch.log1p(torch.tensor(mu)) on the host in float32 (functional.py:697-698)
    return float(torch.log1p(torch.tensor(channels - 1.0, dtype=torch.float32)))


def mulaw_encode(x, channels=256, one_hot="none", reciprocal_divide: Optional[bool] = None):
    """MuLaw.forward: float32 [..., L] -> int64 ([..., L] | [..., L, C] | [..., C, L])  (raw.py:280-292)."""
    lib = _lib.load()
    if isinstance(one_hot, int):
        one_hot = ("none", "categorical", "channel")[one_hot]
    if reciprocal_divide is None:
        # match the eager chain of the device the caller's data lives on: CUDA eager multiplies by the
        # reciprocal of a host scalar, CPU eager divides (mu-law)
        reciprocal_divide = x.is_cuda
    xd = _dev(x)
    if not xd.is_floating_point():
        xd = xd.to(torch.float32)
    xd = xd.to(torch.float32).contiguous()
    dev = xd.device
    L = xd.shape[-1] if xd.ndim else 1
    outer = xd.numel() // max(L, 1)
    if one_hot == "categorical":
        out = torch.empty(tuple(xd.shape) + (channels,), dtype=torch.int64, device=dev)
    elif one_hot == "channel":
        out = torch.empty(tuple(xd.shape[:-1]) + (channels, L), dtype=torch.int64, device=dev)
    else:
        one_hot = "none"
        out = torch.empty(xd.shape, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_mulaw_encode, _ptr(xd), outer, L, channels, _log1p_mu(channels), int(bool(reciprocal_divide)),
                                          (one_hot if isinstance(one_hot, int) else ONEHOT_IDS[one_hot]), _ptr(out), _stream(dev))
    return _ret(out, x)


def mulaw_decode(q, channels=256, reciprocal_divide: Optional[bool] = None):
    """MuLaw.invert: int64 -> float32  (raw.py:314-316)."""
    lib = _lib.load()
    if reciprocal_divide is None:
        reciprocal_divide = q.is_cuda
    qd = _dev(q).to(torch.int64).contiguous()
    out = torch.empty(qd.shape, dtype=torch.float32, device=qd.device)
    if qd.numel() == 0:
        return _ret(out, q)
    with torch.cuda.device(qd.device):
        _run(out, lib.acids_mulaw_decode, _ptr(qd), qd.numel(), channels, _log1p_mu(channels), int(bool(reciprocal_divide)),
                                          _ptr(out), _stream(qd.device))
    return _ret(out, q)


def one_hot(q, n_classes: int):
    """OneHot.forward: int64 [...] -> int64 [..., n_classes]  (misc.py:176-179)."""
    lib = _lib.load()
    qd = _dev(q)
    if qd.dtype != torch.int64:
        raise RuntimeError("one_hot is only applicable to index tensor of type LongTensor.")
    qd = qd.contiguous()
    if n_classes < 1:
        raise RuntimeError("one_hot: n_classes is not set (call scale_data first)")
    out = torch.empty(tuple(qd.shape) + (n_classes,), dtype=torch.int64, device=qd.device)
    with torch.cuda.device(qd.device):
        _run(out, lib.acids_one_hot, _ptr(qd), qd.numel(), n_classes, _ptr(out), _stream(qd.device))
    return _ret(out, q)


# ------------------------------------------------------------------------------------------------
# statistics, raw-domain prologues
# ------------------------------------------------------------------------------------------------
def stats(x, contrast=None, eps=0.0, abs_contrast: bool = False) -> torch.Tensor:
    """(min, max, mean, unbiased std) as a float64[4] DEVICE tensor — no host sync.
    Complex input: statistics of contrast(|x|), what Magnitude.scale_data feeds Normalize (spectral_repr.py:242-245).
    Real input: the values as they are (Normalize.scale_data, norm.py:26-38), or — with `abs_contrast`, what
    Magnitude.scale_data does for a real-valued spectrogram — contrast(|x|) as well."""
    lib = _lib.load()
    xd = _dev(x)
    dev = xd.device
    if torch.is_complex(xd):
        xd = xd.to(torch.complex64).resolve_conj().contiguous()
        kind = _lib.STATS_CABS_CONTRAST
    else:
        xd = xd.to(torch.float32).contiguous()
        kind = _lib.STATS_ABS_CONTRAST if abs_contrast else _lib.STATS_REAL
    if xd.numel() == 0:
        raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0.")
    scratch = torch.empty((int(lib.acids_stats_scratch_bytes()),), dtype=torch.uint8, device=dev)
    out = torch.empty((4,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stats, _ptr(xd), xd.numel(), kind, _cid(contrast), float(eps), _ptr(scratch), _ptr(out), _stream(dev))
    return out


def stft_stats(x, window, n_fft, hop, contrast=None, eps=0.0) -> torch.Tensor:
    """(min, max, mean, unbiased std) of contrast(|STFT(x)|) over every bin as a float64[4] DEVICE tensor — what
    Magnitude.scale_data(STFT(x)) fits (spectral_repr.py:242-245), from ONE pass of the fused forward kernel: the spectrum
    is never written."""
    lib = _lib.load()
    xd = _dev(x)
    _check_stft_input(xd, n_fft)
    xf, _ = _flat_batch(xd, 1)
    B, L = xf.shape
    if B == 0:
        raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0.")
    T = n_frames_centered(L, hop)
    dev = xf.device
    w = _dev(window).to(torch.float32).contiguous()
    scratch = torch.empty((int(lib.acids_stats_scratch_bytes()),), dtype=torch.uint8, device=dev)
    out = torch.empty((4,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _run(out, lib.acids_stft_stats, _ptr(xf), B, L, L, _ptr(w), n_fft, hop, T, _cid(contrast), float(eps), _ptr(scratch),
                                        _ptr(out), _stream(dev))
    return out


def mono_mix(x):
    """Mono(mode="mix"): [..., 2, L] -> [..., L] = (l + r) / 2  (raw.py:37-39)."""
    lib = _lib.load()
    xd = _dev(x).to(torch.float32)
    xf, batch = _flat_batch(xd, 2)
    B, C, L = xf.shape
    assert C == 2
    out = torch.empty((B, L), dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        _run(out, lib.acids_mono_mix, _ptr(xf), B, L, _ptr(out), _stream(xf.device))
    return _ret(out.reshape(tuple(batch) + (L,)), x)


def midside(x, pad_mid=True, inverse=False):
    """MidSide.forward / invert on [..., 2, L]  (raw.py:145-180)."""
    lib = _lib.load()
    xd = _dev(x).to(torch.float32)
    xf, batch = _flat_batch(xd, 2)
    B, C, L = xf.shape
    assert C == 2
    out = torch.empty_like(xf)
    with torch.cuda.device(xf.device):
        _run(out, lib.acids_midside, _ptr(xf), B, L, int(bool(pad_mid)), int(bool(inverse)), _ptr(out), _stream(xf.device))
    return _ret(out.reshape(tuple(batch) + (2, L)), x)
