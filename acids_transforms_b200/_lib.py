"""ctypes binding of libacids_b200.so (the C ABI declared in include/acids_b200.h).

There is no CPU fallback: if the library is missing or cannot be loaded, every op raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACIDS_B200_LIB") or os.path.join(_HERE, "libacids_b200.so")   # override: tuning builds only
ABI_VERSION = 5

ACIDS_OK, ACIDS_EINVAL, ACIDS_ENOTSUP, ACIDS_ECUDA, ACIDS_EWORKSPACE = 0, -1, -2, -3, -4
CONTRAST_IDS = {None: 0, "none": 0, "log1p": 1, "log": 2, "log10": 3}
IF_METHOD_IDS = {"forward": 0, "backward": 1, "central": 2}
PHASE_RAW, PHASE_UNWRAP, PHASE_IF = 0, 1, 2
STATS_REAL, STATS_CABS_CONTRAST, STATS_ABS_CONTRAST = 0, 1, 2
ONEHOT_IDS = {"none": 0, "categorical": 1, "channel": 2}


class Band(Structure):
    """acids_band: banded (column-sparse) matrix descriptor."""
    _fields_ = [("meta", c_void_p), ("coef", c_void_p), ("n_out", c_int32), ("coef_len", c_int32), ("n_in", c_int32)]


class AcidsError(RuntimeError):
    pass


_P = c_void_p
_PROTOTYPES = {
    # name: (restype, argtypes)   — keep in the order of include/acids_b200.h
    "acids_abi_version": (c_int, []),
    "acids_last_error": (c_char_p, []),
    "acids_stft_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, c_int64, _P, _P]),
    "acids_midside_stft_fwd": (c_int, [_P, c_int64, c_int64, _P, c_int, c_int, c_int64, c_int, _P, _P]),
    "acids_stft_mag_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, c_int64, Band, c_int,
                                   c_float, _P, _P, c_int, _P, c_int64, c_int64, _P]),
    "acids_stft_polar_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int64, c_int, Band, c_int, c_float,
                                     _P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, c_int64, c_int64, _P, c_int64, c_int64, _P]),
    "acids_mag_epilogue": (c_int, [_P, c_int64, c_int, Band, c_int, c_float, _P, _P, c_int, _P, c_int64, _P]),
    "acids_mag_invert": (c_int, [_P, c_int64, c_int, c_int64, c_int, Band, c_int, c_float, _P, _P, _P, _P]),
    "acids_melspec_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int64, Band, c_float, _P, _P, _P, _P]),
    "acids_mfcc_dct": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int, c_float, c_int64, _P, _P, _P]),
    "acids_mfcc_dct_tc": (c_int, [_P, c_int64, c_int, c_int64, _P, c_int, c_float, c_int64, _P, _P, _P]),
    "acids_phase_fwd": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, c_int64, c_int64, _P]),
    "acids_polar_fwd": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_float, _P, _P, c_int, c_int, c_int, _P, _P, c_int,
                                _P, c_int64, c_int64, _P, c_int64, c_int64, _P]),
    "acids_polar_rows_fwd": (c_int, [_P, c_int64, c_int64, c_int, Band, c_int, c_float, _P, _P, c_int, c_int, c_int, _P, _P, c_int,
                                     _P, c_int64, _P, c_int64, _P]),
    "acids_phase_inv": (c_int, [_P, c_int64, c_int64, c_int, c_int64, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "acids_phase_inv_polar": (c_int, [_P, c_int64, c_int64, c_int, c_int64, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "acids_polar_to_complex": (c_int, [_P, _P, c_int64, _P, _P]),
    "acids_griffinlim_update": (c_int, [_P, _P, _P, c_float, c_int64, _P, _P]),
    "acids_istft_workspace_bytes": (c_int64, [c_int64, c_int64, c_int, c_int]),
    "acids_istft_ola": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "acids_irfft_frames": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "acids_ola_stream": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int64, _P, c_float, _P, _P, _P]),
    "acids_stream_analysis": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P]),
    "acids_stream_synthesis": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, c_float, _P, _P, _P]),
    "acids_stream_roundtrip": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, c_float, _P, _P, _P, _P, _P]),
    "acids_pghi_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "acids_mel_tc_workspace_bytes": (c_int64, [c_int]),
    "acids_mel_tc": (c_int, [_P, c_int64, c_int64, c_int, _P, c_int, _P, c_int64, _P, _P]),
    "acids_rt_pghi_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "acids_rt_pghi": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_float, c_int, c_int, c_float, c_float, _P, c_int64, _P, _P]),
    "acids_pghi": (c_int, [_P, c_int64, c_int64, c_int, c_float, c_int, c_int, ctypes.c_double, c_float, _P, c_int64, _P, _P]),
    "acids_mulaw_encode": (c_int, [_P, c_int64, c_int64, c_int, c_float, c_int, c_int, _P, _P]),
    "acids_mulaw_decode": (c_int, [_P, c_int64, c_int, c_float, c_int, _P, _P]),
    "acids_one_hot": (c_int, [_P, c_int64, c_int, _P, _P]),
    "acids_stats_scratch_bytes": (c_int64, []),
    "acids_stats": (c_int, [_P, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "acids_stft_stats": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int64, c_int, c_float, _P, _P, _P]),
    "acids_mono_mix": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "acids_midside": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P]),
}
EXPORTS = tuple(_PROTOTYPES)

_lib = None


def load():
    """Load the shared library once; fail loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AcidsError(
            "libacids_b200.so is not built (%s). Run `python -m acids_transforms_b200.build`; "
            "this package has no CPU or eager fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.acids_abi_version() != ABI_VERSION:
        raise AcidsError("libacids_b200.so ABI %d != expected %d; rebuild" % (lib.acids_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, lib=None):
    """Map a negative return code to the exception class the reference would raise for it."""
    if rc == ACIDS_OK:
        return
    lib = lib or load()
    msg = lib.acids_last_error().decode("utf-8", "replace")
    if rc in (ACIDS_EINVAL, ACIDS_ENOTSUP):
        # torch raises RuntimeError for bad shapes (e.g. stft expected 0 < n_fft <= length); SURVEY §8b
        raise RuntimeError("acids_b200: " + msg)
    raise AcidsError("acids_b200 (code %d): %s" % (rc, msg))
