"""Regression tests for the round-1 review findings (VERDICT.md / ADVICE.md):

* a mel bank built for another row length is refused (the reference's matmul raises) — host side and C ABI;
* `invert` does not synchronise the host after its first call (the envelope verdict is cached) and can be captured
  in a CUDA graph;
* `Magnitude.scale_data` on a real-valued spectrogram fits contrast(|x|) like the reference;
* `phase_buffer` holds the reference's values when tracking is on, `track_phase=True` restores the reference
  default, a `keep_input` request without a buffer warns;
* reference checkpoints load (MFCC's torchaudio keys), Python mirrors follow `load_state_dict`.
"""
import ctypes
import warnings

import numpy as np
import pytest
import torch

from conftest import assert_parity, branch_cut, load_golden


# ---------------------------------------------------------------------------------------------
# host-only
# ---------------------------------------------------------------------------------------------
def test_banded_matrix_carries_its_input_size():
    from acids_transforms_b200 import ops
    from acids_transforms_b200.transforms import Magnitude
    m = Magnitude(n_fft=512)
    b = ops.as_band(m.mel_meta, m.mel_coef)
    assert b.n_in == 257 and b.n_out == 257
    assert ops.BandedMatrix(m.mel_bank).on(torch.device("cpu")).n_in == 257      # travels in the C struct


def test_envelope_verdict_is_cached_per_window_version():
    from acids_transforms_b200 import ops
    w = torch.hann_window(1024)
    ops._ENVELOPE_CACHE.clear()
    assert ops.istft_envelope_ok(w, 1024, 256, 690)
    n = len(ops._ENVELOPE_CACHE)
    assert ops.istft_envelope_ok(w, 1024, 256, 345) and len(ops._ENVELOPE_CACHE) == n       # same effective frame count
    w.mul_(0.0)                                                                             # in-place edit bumps _version
    assert not ops.istft_envelope_ok(w, 1024, 256, 690)


def test_track_phase_switch_and_state_dict_mirrors():
    from acids_transforms_b200.transforms import STFT, DGT, Magnitude, OverlapAdd, MFCC
    assert STFT().track_phase is False and STFT(track_phase=True).track_phase is True
    assert STFT(inversion_mode="keep_input").track_phase is True and DGT(track_phase=True).track_phase is True
    # Python mirrors follow load_state_dict
    a, b = Magnitude(eps=1e-3), Magnitude()
    b.load_state_dict(a.state_dict())
    assert b._eps == pytest.approx(1e-3)
    o1, o2 = OverlapAdd(512, 64), OverlapAdd(512, 64)
    sd = o1.state_dict()
    sd["gain_compensation"] = torch.tensor(3.0)
    o2.load_state_dict(sd)
    assert o2._gain == 3.0 and o2.keep == 448
    # a reference MFCC checkpoint holds torchaudio's buffers (mel.py:43-44)
    m = MFCC(n_fft=512, hop_length=128, n_mels=40)
    sd = m.state_dict()
    sd["transform.spectrogram.window"] = torch.hann_window(512)
    sd["transform.mel_scale.fb"] = m.dense_bank()
    m.load_state_dict(sd, strict=True)
    sd["transform.mel_scale.fb"] = torch.zeros(100, 40)
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd, strict=True)


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
def test_mel_bank_shape_mismatch_raises():
    """STFT(512) + Magnitude() (n_fft 1024 bank): the reference's matmul raises; so must every path here."""
    from acids_transforms_b200 import transforms as T, ops, _lib
    x = torch.randn(2, 8192, device="cuda")
    stft, mag = T.STFT(n_fft=512, hop_length=128).cuda(), T.Magnitude().cuda()
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        mag(stft(x))
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        mag.invert(torch.rand(2, 65, 257, device="cuda"))
    band = ops.as_band(mag.mel_meta, mag.mel_coef)
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        ops.stft_mag_fwd(x, stft.window, 512, 128, band, "log1p", 1e-7, None, None)
    # and the C ABI itself (no Python check in between)
    lib = _lib.load()
    out = torch.empty(2, 65, 513, device="cuda")
    w = stft.window[:512].contiguous()
    rc = lib.acids_stft_mag_fwd(ctypes.c_void_p(x.data_ptr()), 2, 8192, 8192, ctypes.c_void_p(w.data_ptr()), 512, 128, 1, 65,
                                band.on(x.device), 1, 1e-7, None, None, 0, ctypes.c_void_p(out.data_ptr()), 65 * 513, 513, None)
    assert rc == _lib.ACIDS_EINVAL and b"cannot be multiplied" in lib.acids_last_error()
    # a chain whose STFT is re-parameterised after the fused plan was built must not read out of bounds either
    ch = (T.STFT(n_fft=1024, hop_length=256) + T.Magnitude(mode=None)).cuda()
    ch[0].set_params(512, 128)
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        ch(x)


@pytest.mark.gpu
def test_invert_is_graph_capturable_and_sync_free():
    """After one warm-up call STFT.invert / DGT.invert enqueue work only: capture them in a CUDA graph (any host
    synchronisation or blocking copy would make the capture fail) and replay."""
    from acids_transforms_b200 import transforms as T
    for mod in (T.STFT(n_fft=1024, hop_length=256).cuda(), T.DGT(n_fft=512, hop_length=128, inversion_mode="random").cuda()):
        x = torch.randn(4, 16384, device="cuda")
        X = mod(x)
        ref = mod.invert(X)                 # warm-up: reads the window back once, caches the verdict
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                y = mod.invert(X)
        y.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, ref)


@pytest.mark.gpu
def test_magnitude_scale_data_on_real_spectrogram():
    """spectral_repr.py:242-245 fits contrast(x.abs()) for real input too."""
    from acids_transforms_b200 import transforms as T
    g = load_golden("magnitude_1024")
    mag_in = np.abs(g["X"]).astype(np.float32)
    m = T.Magnitude().cuda()
    m.scale_data(cu(mag_in))
    assert abs(float(m.norm.offset) - float(g["default_offset"])) <= 1e-6
    assert abs(float(m.norm.scale) - float(g["default_scale"])) <= 1e-5 * float(g["default_scale"])
    assert_parity(m(cu(mag_in)).cpu().numpy(), g["default_y"], 1e-4, "Magnitude on |X|")
    # signed real input: |.| is taken first
    m2 = T.Magnitude(mel=False, contrast="log").cuda()
    v = torch.randn(3, 9, 513, device="cuda")
    m2.scale_data(v)
    ref = torch.log(v.abs().clamp_min(m2._eps))
    assert abs(float(m2.norm.offset) - float(ref.min())) <= 1e-5 and abs(float(m2.norm.scale) - float(ref.max() - ref.min())) <= 1e-4


@pytest.mark.gpu
def test_phase_buffer_matches_reference():
    """Row A4: with tracking on, `phase_buffer` is the reference's x_fft.angle(), flattened [B, T, F] (stft.py:103, :134)."""
    from acids_transforms_b200 import transforms as T
    g = load_golden("stft_1024_256")
    for mod in (T.STFT(n_fft=1024, hop_length=256, track_phase=True), T.STFT(n_fft=1024, hop_length=256, inversion_mode="keep_input")):
        mod = mod.cuda()
        X = mod(cu(g["x"]))
        pb = mod.phase_buffer.cpu().numpy()
        assert pb.shape == g["phase_buffer"].shape
        ok = ~branch_cut(g["X"])
        assert ok.mean() > 0.9
        d = np.abs(pb - g["phase_buffer"])[ok]
        assert d.max() <= 1e-4 * np.pi, d.max()
        # keep_input inversion consumes it: exact reconstruction of the input on the trimmed span
        y = mod.invert(X.abs(), inversion_mode="keep_input").cpu().numpy()
        assert_parity(y, g["y"], 1e-4, "keep_input")
    # the reference's own test flow: default module, per-call keep_input -> warns once, then records
    mod = T.STFT(n_fft=1024, hop_length=256).cuda()
    X = mod(cu(g["x"]))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mod.invert(X.abs(), inversion_mode="keep_input")
    assert any("keep_input" in str(i.message) for i in w) and mod.track_phase
    X = mod(cu(g["x"]))
    assert_parity(mod.invert(X.abs(), inversion_mode="keep_input").cpu().numpy(), g["y"], 1e-4, "keep_input after warning")
