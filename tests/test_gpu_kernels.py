"""GPU parity tests: every kernel is called through the C ABI (acids_transforms_b200.ops -> ctypes ->
libacids_b200.so) and compared with (a) the golden vectors minted from the unmodified reference and
(b) the numpy oracle on seeded inputs.  Tolerance: SURVEY.md §8(d) "rel 1e-4" (conftest.assert_parity);
integer outputs bit-exact.  Run on the B200 box: python -m pytest tests -m gpu
"""
import math

import numpy as np
import pytest
import torch

from conftest import assert_parity, branch_cut, dense, if_mask, load_golden, unwrap_mask
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-4   # north_star: within rel 1e-4 for fp32 spectra, mel/MFCC and round-trip reconstruction


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from acids_transforms_b200 import ops as _ops, _lib
    _lib.load()          # fails loudly if libacids_b200.so is missing: no fallback may hide it
    return _ops


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().resolve_conj().numpy()


def band_from_golden(ops, name, inverse=False):
    g = load_golden(name)
    m = dense(g["inv_rows" if inverse else "rows"], g["inv_cols" if inverse else "cols"],
              g["inv_vals" if inverse else "vals"], g["shape"])
    return ops.BandedMatrix(torch.from_numpy(m)), m


def synth(n_clips, L, seed):
    rng = np.random.default_rng(seed)
    x = 0.5 * (2 * rng.random((n_clips, L), dtype=np.float32) - 1)
    n = np.arange(L)
    for i in range(n_clips):
        x[i] += (0.25 * np.sin(2 * np.pi * 55.0 * 2 ** ((i % 84) / 12) * n / 44100)).astype(np.float32)
    return x.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# (1) STFT forward / (4) ISTFT against the reference's golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["stft_1024_256", "stft_512_128", "stft_2048_512", "stft_4096_1024",
                                  "stft_256_64", "stft_ragged_512_128", "stft_stereo_512_128"])
def test_stft_golden(ops, name):
    g = load_golden(name)
    n, h = int(g["n_fft"]), int(g["hop"])
    w = torch.hann_window(n).cuda()
    X = ops.stft_fwd(cu(g["x"]), w, n, h)
    assert X.dtype == torch.complex64 and X.is_contiguous() and tuple(X.shape) == g["X"].shape
    assert_parity(host(X), g["X"], REL, name + " fwd")
    # DC / Nyquist carry an exact +0 imaginary part like the reference's r2c (phase 0 or +pi)
    Xh = host(X)
    assert not np.signbit(Xh.imag[..., 0]).any() and not np.signbit(Xh.imag[..., -1]).any()
    y = ops.istft_ola(cu(g["X"]), w, n, h)
    assert_parity(host(y), g["y"], REL, name + " inv")
    # round trip on the trimmed span (Hann COLA)
    y2 = host(ops.istft_ola(X, w, n, h))
    assert_parity(y2, g["x"][..., :y2.shape[-1]], REL, name + " roundtrip")


@pytest.mark.parametrize("name", ["dgt_1024_256", "dgt_512_128"])
def test_dgt_golden(ops, name):
    g = load_golden(name)
    n, h = int(g["n_fft"]), int(g["hop"])
    X = ops.stft_fwd(cu(g["x"]), cu(g["window"]), n, h)
    assert_parity(host(X), g["X"], REL, name + " fwd")
    y = ops.istft_ola(cu(g["X"]), cu(g["inv_window"]), n, h)
    assert_parity(host(y), g["y"], REL, name + " inv")     # quirk A15: not the identity, must match the reference


@pytest.mark.parametrize("n_fft,hop,L,B", [(32, 8, 100, 3), (64, 16, 333, 2), (128, 32, 1000, 3), (256, 128, 2049, 2),
                                           (512, 512, 4096, 2), (1024, 256, 20000, 5), (2048, 512, 30001, 3),
                                           (4096, 1024, 40000, 2), (8192, 2048, 50000, 2), (16384, 4096, 70000, 1),
                                           (1024, 100, 5000, 2), (1024, 128, 9000, 2), (512, 127, 3000, 2), (256, 85, 2001, 3),
                                           (1024, 256, 20002, 3), (512, 128, 10002, 3), (1024, 256, 176400, 2)])
def test_stft_istft_oracle(ops, n_fft, hop, L, B):
    """All supported n_fft (32..16384), odd lengths and rows that are only 8-byte aligned (staged load path), odd
    hops, hop == n_fft, and a full 4 s clip (the persistent inverse starts segments inside a clip: halo frames)."""
    x = synth(B, L, n_fft + hop)
    w = O.periodic_window("hann", n_fft)
    Xo = O.stft(x, n_fft, hop, w)
    X = ops.stft_fwd(cu(x), cu(w), n_fft, hop)
    assert_parity(host(X), Xo, REL, "fwd %d/%d" % (n_fft, hop))
    if O.istft_envelope(Xo.shape[-2], n_fft, hop, w)[n_fft // 2:-(n_fft // 2)].min() > 1e-11 and Xo.shape[-2] > 1:
        y = ops.istft_ola(cu(Xo), cu(w), n_fft, hop)
        assert_parity(host(y), O.istft(Xo, n_fft, hop, w), REL, "inv %d/%d" % (n_fft, hop))


def test_fuzz_against_torch_on_device(ops):
    """Seeded random geometries (n_fft, hop, length, batch, window) against torch.stft / torch.istft on the same GPU
    (cuFFT: an implementation independent of both this code and the numpy oracle)."""
    rng = np.random.default_rng(20261018)
    for case in range(24):
        n_fft = int(2 ** rng.integers(5, 13))                       # 32 .. 4096
        hop = int(rng.choice([n_fft // 8, n_fft // 4, n_fft // 2, max(1, n_fft // 4 - 1), max(1, n_fft // 3)]))
        B = int(rng.integers(1, 6))
        L = int(n_fft // 2 + 1 + rng.integers(0, 12 * n_fft))
        g = torch.Generator(device="cuda").manual_seed(case)
        x = 2 * torch.rand((B, L), generator=g, device="cuda") - 1
        w = torch.hann_window(n_fft, device="cuda") if case % 2 == 0 else torch.hamming_window(n_fft, device="cuda")
        want = torch.stft(x, n_fft, hop, window=w, return_complex=True).transpose(-2, -1)
        got = ops.stft_fwd(x, w, n_fft, hop)
        tag = "case %d n_fft=%d hop=%d L=%d B=%d" % (case, n_fft, hop, L, B)
        assert got.shape == want.shape, tag
        assert_parity(host(torch.view_as_real(got)), host(torch.view_as_real(want.contiguous())), REL, "fuzz fwd " + tag)
        if want.shape[-2] > 1 and ops.istft_envelope_ok(w, n_fft, hop, want.shape[-2]):
            y_want = torch.istft(want.transpose(-2, -1), n_fft, hop, window=w)
            y = ops.istft_ola(want.contiguous(), w, n_fft, hop)
            assert_parity(host(y), host(y_want), REL, "fuzz inv " + tag)


def test_istft_partial_sum_mode(ops):
    """n_fft >= 4096 with n_fft % hop == 0 keeps partial overlap-add sums instead of a frame ring: long clips in a small
    batch (so that CTAs start and stop inside clips and recompute halo frames), every supported overlap, clip edges,
    against torch.istft on the same GPU; the unaligned hop takes the frame-ring route of the same kernel."""
    for n_fft, hop, B, T in ((4096, 1024, 3, 700), (4096, 2048, 2, 300), (4096, 512, 2, 500), (4096, 4096 // 16, 1, 400),
                             (8192, 2048, 3, 250), (8192, 1024, 1, 300), (16384, 4096, 2, 90), (4096, 1000, 2, 300),
                             (4096, 1024, 5, 2), (4096, 1024, 4, 3)):
        g = torch.Generator(device="cuda").manual_seed(n_fft + hop + T)
        w = torch.hann_window(n_fft, device="cuda")
        X = torch.view_as_complex(torch.randn((B, T, n_fft // 2 + 1, 2), generator=g, device="cuda"))
        X[..., 0] = X[..., 0].real + 0j
        X[..., -1] = X[..., -1].real + 0j
        tag = "n_fft=%d hop=%d B=%d T=%d" % (n_fft, hop, B, T)
        assert ops.istft_envelope_ok(w, n_fft, hop, T), tag
        want = torch.istft(X.transpose(-2, -1), n_fft, hop, window=w)
        got = ops.istft_ola(X, w, n_fft, hop)
        assert got.shape == want.shape, tag
        assert_parity(host(got), host(want), REL, "istft partial sums " + tag)
        again = ops.istft_ola(X, w, n_fft, hop)
        assert torch.equal(got, again), "not deterministic: " + tag


def test_stft_known_answers(ops):
    n, h = 1024, 256
    w = torch.hann_window(n).cuda()
    # impulse in the middle of frame t: flat magnitude w[n0]
    x = torch.zeros(1, 8192).cuda()
    x[0, 4096] = 1.0
    X = ops.stft_fwd(x, w, n, h)            # frame 16 is centred on sample 4096 -> tap w[512] = 1
    assert_parity(host(X[0, 16].abs()), np.ones(513, np.float32), REL, "impulse")
    # bin-centred sinusoid: |X[k0]| = sum(w)/2
    k0 = 64
    t = torch.arange(8192, dtype=torch.float64)
    x = torch.cos(2 * math.pi * k0 * t / n).float().cuda()[None]
    X = ops.stft_fwd(x, w, n, h)
    mid = host(X[0, 16].abs())
    assert abs(mid[k0] - float(w.sum()) / 2) <= REL * float(w.sum()) / 2
    assert mid[k0 + 3:].max() < 1e-3 * mid[k0]


def test_stft_errors(ops):
    w = torch.hann_window(1024).cuda()
    with pytest.raises(RuntimeError):
        ops.stft_fwd(torch.zeros(2, 300).cuda(), w, 1024, 256)           # reflect pad longer than the input
    with pytest.raises(RuntimeError):
        ops.stft_fwd(torch.zeros(2, 3000).cuda(), torch.hann_window(1000).cuda(), 1000, 250)   # not a power of two
    with pytest.raises(RuntimeError):
        ops.istft_ola(torch.zeros(2, 5, 513, dtype=torch.complex64).cuda(), w, 1024, 1024)     # Hann at hop=N: zero envelope


def test_empty_batch(ops):
    w = torch.hann_window(512).cuda()
    X = ops.stft_fwd(torch.zeros(0, 4096).cuda(), w, 512, 128)
    assert tuple(X.shape) == (0, 33, 257)
    assert tuple(ops.istft_ola(X, w, 512, 128).shape) == (0, 4096)


# ---------------------------------------------------------------------------------------------
# (2) Magnitude: fused and stand-alone, forward and inverse
# ---------------------------------------------------------------------------------------------
MAG_CASES = {
    "default": dict(mode="unipolar", contrast="log1p", mel=True, keep_nyquist=True),
    "bipolar_log": dict(mode="bipolar", contrast="log", mel=True, keep_nyquist=True),
    "gauss_log10": dict(mode="gaussian", contrast="log10", mel=True, keep_nyquist=True),
    "none_none_nomel": dict(mode=None, contrast=None, mel=False, keep_nyquist=True),
    "unipolar_log1p_nomel": dict(mode="unipolar", contrast="log1p", mel=False, keep_nyquist=True),
    "nonyq": dict(mode="unipolar", contrast="log1p", mel=True, keep_nyquist=False),
}
EPS = float(np.finfo(np.float32).eps)


@pytest.mark.parametrize("tag", sorted(MAG_CASES))
def test_magnitude_golden(ops, tag):
    g = load_golden("magnitude_1024")
    c = MAG_CASES[tag]
    bank = "mel_bank_1024" if c["keep_nyquist"] else "mel_bank_1024_nonyq"
    fwd = band_from_golden(ops, bank)[0] if c["mel"] else None
    inv = band_from_golden(ops, bank, inverse=True)[0] if c["mel"] else None
    off = cu(g[tag + "_offset"].reshape(1)) if c["mode"] else None
    sc = cu(g[tag + "_scale"].reshape(1)) if c["mode"] else None
    y = ops.mag_epilogue(cu(g["X"]), fwd, c["contrast"], EPS, off, sc, drop_first=not c["keep_nyquist"])
    assert_parity(host(y), g[tag + "_y"], REL, tag + " fwd")
    m = ops.mag_invert(cu(g[tag + "_y"]), inv, c["contrast"], EPS, off, sc, pad_last=not c["keep_nyquist"])
    assert_parity(host(m), g[tag + "_inv"], REL, tag + " inv")
    if c["mode"]:
        st = host(ops.stats(cu(g["X"]), c["contrast"], EPS))
        o_off, o_sc = O.normalize_stats(O.magnitude_stats_input(g["X"], c["contrast"]), c["mode"])
        mn, mx, mean, std = st
        got = {"unipolar": (mn, mx - mn), "bipolar": ((mx + mn) / 2, mx - (mx + mn) / 2), "gaussian": (mean, std)}[c["mode"]]
        assert abs(got[0] - float(g[tag + "_offset"])) <= REL * max(abs(float(g[tag + "_scale"])), 1e-6)
        assert abs(got[1] - float(g[tag + "_scale"])) <= REL * abs(float(g[tag + "_scale"]))


@pytest.mark.parametrize("name,window", [("chain_cfg2", "gauss"), ("chain_cfg1", "hann")])
def test_fused_chain_golden(ops, name, window):
    """BASELINE.json cfg 1 / cfg 2 at fixture size: wave -> normalised log-mel in ONE kernel."""
    g = load_golden(name)
    band, _ = band_from_golden(ops, "mel_bank_1024")
    x = cu(g["x"])
    if name == "chain_cfg1":
        x = ops.mono_mix(x)
    w = cu(O.dgt_window(1024)) if window == "gauss" else torch.hann_window(1024).cuda()
    y = ops.stft_mag_fwd(x, w, 1024, 256, band, "log1p", EPS, cu(g["offset"].reshape(1)), cu(g["scale"].reshape(1)))
    assert_parity(host(y), g["y"], REL, name)
    # and the unfused path gives the same thing
    y2 = ops.mag_epilogue(ops.stft_fwd(x, w, 1024, 256), band, "log1p", EPS, cu(g["offset"].reshape(1)), cu(g["scale"].reshape(1)))
    assert_parity(host(y2), g["y"], REL, name + " unfused")


@pytest.mark.parametrize("n_fft,hop", [(128, 32), (512, 128), (1024, 256), (1024, 100), (1024, 1024), (2048, 512), (4096, 1024), (8192, 2048)])
def test_fused_magnitude_oracle(ops, n_fft, hop):
    """Small plans (many frames per CTA, several row tiles per unit) up to a bank that does not fit in shared memory
    (n_fft = 8192: banded matrix read from global memory).  The clip length is odd, so every row but the first is unaligned
    and its frames go through the staged path — at n_fft = 1024 that is the one-exchange plan with the rows parked in the
    exchange buffers (hop 100: unaligned frame starts as well; hop = n_fft: no overlap)."""
    x = synth(3, 6 * n_fft + 17, n_fft)
    w = O.periodic_window("hann", n_fft)
    fwd, _ = O.magnitude_banks(44100, n_fft)
    band = ops.BandedMatrix(torch.from_numpy(fwd))
    X = O.stft(x, n_fft, hop, w)
    lin = O.magnitude_forward(X, fwd, None)
    for contrast in ("log1p", "log", None):
        yo = O.magnitude_forward(X, fwd, contrast, offset=0.3, scale=1.7)
        y = host(ops.stft_mag_fwd(cu(x), cu(w), n_fft, hop, band, contrast, EPS, 0.3, 1.7))
        if contrast == "log":
            # log() turns the RELATIVE error of a bin into an absolute one: bins 1000x below the peak carry
            # 1e-3 relative FFT rounding in any fp32 transform, so judge log-contrast on the others
            ok = lin > 1e-3 * lin.max()
            assert ok.mean() > 0.7          # 109 of the 513 square-bank columns are all-zero
            y, yo = np.where(ok, y, 0), np.where(ok, yo, 0)
        assert_parity(y, yo, REL, "fused %d %s" % (n_fft, contrast))
    yo = O.magnitude_forward(X, None, "log1p", keep_nyquist=False)
    y = ops.stft_mag_fwd(cu(x), cu(w), n_fft, hop, None, "log1p", EPS, None, None, drop_first=True)
    assert_parity(host(y), yo, REL, "fused nomel nonyq %d" % n_fft)


# ---------------------------------------------------------------------------------------------
# (3) Phase / unwrap / IF / Polar
# ---------------------------------------------------------------------------------------------
def masked(a, ok):
    return np.where(ok, a, 0)


def test_phase_if_golden(ops):
    from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW, PHASE_UNWRAP
    g = load_golden("phase_if_256")
    X = g["X"]
    Xc = cu(X)
    raw_ok = ~branch_cut(X)
    assert raw_ok.mean() > 0.95
    assert_parity(masked(host(ops.phase_fwd(Xc, PHASE_RAW)), raw_ok), masked(g["phase_raw_y"], raw_ok), REL, "phase")
    un_ok = unwrap_mask(X)
    assert_parity(masked(host(ops.phase_fwd(Xc, PHASE_UNWRAP)), un_ok), masked(g["phase_unwrap_y"], un_ok), REL, "unwrap")
    off, sc = cu(g["phase_unwrap_bipolar_offset"].reshape(1)), cu(g["phase_unwrap_bipolar_scale"].reshape(1))
    assert_parity(masked(host(ops.phase_fwd(Xc, PHASE_UNWRAP, offset=off, scale=sc)), un_ok),
                  masked(g["phase_unwrap_bipolar_y"], un_ok), REL, "unwrap bipolar")
    y = host(ops.phase_fwd(Xc, PHASE_RAW, drop_first=True))
    assert_parity(masked(y, raw_ok[..., 1:]), masked(g["phase_nonyq_y"], raw_ok[..., 1:]), REL, "phase nonyq")
    assert_parity(host(ops.phase_inv(cu(g["phase_nonyq_y"]), PHASE_RAW, pad_last=True)), g["phase_nonyq_inv"], 1e-6, "phase nonyq inv")
    for method in ("forward", "backward", "central"):
        ok = if_mask(X, method)
        off, sc = cu(g["if_%s_offset" % method].reshape(1)), cu(g["if_%s_scale" % method].reshape(1))
        y = host(ops.phase_fwd(Xc, PHASE_IF, method, False, off, sc))
        assert_parity(masked(y, ok), masked(g["if_%s_y" % method], ok), REL, "if " + method)
        yw = host(ops.phase_fwd(Xc, PHASE_IF, method, True))
        assert_parity(masked(yw, ok), masked(g["if_%s_w_y" % method], ok), REL, "if weighted " + method)
        inv = host(ops.phase_inv(cu(g["if_%s_y" % method]), PHASE_IF, method, off, sc))
        assert_parity(inv, g["if_%s_inv" % method], REL, "if inv " + method)
    # statistics of the IF for Normalize("gaussian"): mean / unbiased std
    raw_if = ops.phase_fwd(Xc, PHASE_IF, "forward")
    st = host(ops.stats(raw_if))
    assert abs(st[2] - float(g["if_forward_offset"])) <= 5e-3 * float(g["if_forward_scale"])   # a few +-pi flips on the branch cut
    assert abs(st[3] - float(g["if_forward_scale"])) <= 5e-3 * float(g["if_forward_scale"])


def test_polar_golden(ops):
    from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW
    g = load_golden("polar_1024")
    X = g["X"]
    Xc = cu(X)
    fwd, _ = band_from_golden(ops, "mel_bank_1024")
    inv, _ = band_from_golden(ops, "mel_bank_1024", inverse=True)
    for tag, mode in (("polar", PHASE_RAW), ("polarif", PHASE_IF)):
        mo, ms = cu(g[tag + "_mag_offset"].reshape(1)), cu(g[tag + "_mag_scale"].reshape(1))
        po, ps = cu(g[tag + "_ph_offset"].reshape(1)), cu(g[tag + "_ph_scale"].reshape(1))
        out = torch.empty(X.shape[:-1] + (2, X.shape[-1]), dtype=torch.float32, device="cuda")
        ops.mag_epilogue(Xc, fwd, "log1p", EPS, mo, ms, out=out.view(-1, 2, X.shape[-1]), out_slot=0, out_slots=2)
        ops.phase_fwd(Xc, mode, "forward", False, po, ps, out=out.view(-1, X.shape[-2], 2, X.shape[-1]), out_slot=1, out_slots=2)
        y = host(out)
        ok = if_mask(X) if mode == PHASE_IF else ~branch_cut(X)
        assert_parity(y[..., 0, :], g[tag + "_y"][..., 0, :], REL, tag + " mag slot")
        assert_parity(masked(y[..., 1, :], ok), masked(g[tag + "_y"][..., 1, :], ok), REL, tag + " phase slot")
        gy = cu(g[tag + "_y"])
        m = ops.mag_invert(gy[..., 0, :], inv, "log1p", EPS, mo, ms)
        p = ops.phase_inv(gy[..., 1, :], mode, "forward", po, ps)
        Z = host(ops.polar_to_complex(m, p))
        assert_parity(Z, g[tag + "_inv"], REL, tag + " inv")
        # the one-pass form the modules use: same phase, SFU sine / cosine (see test_polar_recombination_accuracy)
        Zf = host(ops.phase_inv_polar(gy[..., 1, :], m, mode, "forward", po, ps))
        assert_parity(Zf, g[tag + "_inv"], REL, tag + " fused inv")
        assert float(np.abs(Zf - Z).max()) <= 4e-6 * float(np.abs(Z).max()), tag + " fused vs two kernels"


def test_polar_fwd_one_pass(ops):
    """polar_fwd (one read of the spectrum) == mag_epilogue(band=None) + phase_fwd: every contrast, mode, method,
    weighting and keep_nyquist=False; frame counts around the 8-frame blocks of the kernel."""
    from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW, PHASE_UNWRAP
    g = torch.Generator(device="cuda").manual_seed(9)
    mo, ms = torch.tensor([0.3], device="cuda"), torch.tensor([2.5], device="cuda")
    po, ps = torch.tensor([-0.1], device="cuda"), torch.tensor([0.75], device="cuda")
    cases = [(PHASE_RAW, "forward", False), (PHASE_UNWRAP, "forward", False), (PHASE_IF, "forward", False),
             (PHASE_IF, "backward", True), (PHASE_IF, "central", False)]
    for T in (2, 7, 8, 9, 26):
        X = torch.view_as_complex(torch.randn(3, 2, T, 70, 2, device="cuda", generator=g))
        for contrast in ("none", "log1p", "log", "log10"):
            for mode, method, weighted in cases:
                for drop in (False, True):
                    got = ops.polar_fwd(X, contrast, EPS, mo, ms, mode, method, weighted, po, ps, drop)
                    m = ops.mag_epilogue(X, None, contrast, EPS, mo, ms, drop)
                    ph = ops.phase_fwd(X, mode, method, weighted, po, ps, drop)
                    assert got.shape == (3, 2, T, 2, 70 - int(drop))
                    torch.testing.assert_close(got[..., 0, :], m, rtol=1e-6, atol=1e-6)
                    assert torch.equal(got[..., 1, :], ph), (T, contrast, mode, method, drop)


def test_polar_recombination_accuracy(ops):
    """phase_inv_polar evaluates exp(i phase) with a 2-pi reduction + the SFU sine / cosine: absolute error against
    float64 stays below 2e-6 * mag for integrated phases of thousands of radians (parity budget 1e-4), exp(i 0) is
    exact; polar_to_complex (library sincosf) is held to 5e-7."""
    from acids_transforms_b200._lib import PHASE_RAW
    g = torch.Generator(device="cuda").manual_seed(21)
    for span in (3.2, 700.0, 3.0e4):
        ph = (2 * torch.rand((4, 16, 4099), generator=g, device="cuda") - 1) * span
        mag = torch.rand((4, 16, 4099), generator=g, device="cuda") + 0.5
        p64, m64 = ph.double().cpu(), mag.double().cpu()
        want = torch.stack([m64 * torch.cos(p64), m64 * torch.sin(p64)], -1)
        for fn, bound in ((lambda: ops.phase_inv_polar(ph, mag, PHASE_RAW), 2e-6), (lambda: ops.polar_to_complex(mag, ph), 5e-7)):
            got = torch.view_as_real(fn()).double().cpu()
            err = float(((got - want).abs() / m64[..., None]).max())
            assert err < bound, "span %g: %.3e" % (span, err)
    one = ops.phase_inv_polar(torch.zeros(2, 8, device="cuda"), torch.full((2, 8), 3.0, device="cuda"), PHASE_RAW)
    assert torch.equal(torch.view_as_real(one).cpu(), torch.tensor([3.0, 0.0]).expand(2, 8, 2))


def test_phase_inv_polar_variants(ops):
    """phase_inv_polar == polar_to_complex(mag, phase_inv(...)) up to the SFU sine / cosine (4e-6 * mag): every mode /
    method, strided input (the stacked [.., 2, F] layout), the appended zero bin, and the central method's two-kernel
    route (identical bits); the one-pass kernel itself is deterministic."""
    from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW, PHASE_UNWRAP
    g = torch.Generator(device="cuda").manual_seed(5)
    stacked = torch.randn(3, 2, 17, 2, 96, device="cuda", generator=g)
    y = stacked[..., 1, :]
    off = torch.tensor([0.25], device="cuda")
    sc = torch.tensor([1.5], device="cuda")
    for mode, method in ((PHASE_RAW, "forward"), (PHASE_UNWRAP, "forward"), (PHASE_IF, "forward"), (PHASE_IF, "backward"), (PHASE_IF, "central")):
        for pad in (False, True):
            mag = torch.rand(3, 2, 17, 96 + int(pad), device="cuda", generator=g)
            want = ops.polar_to_complex(mag, ops.phase_inv(y, mode, method, off, sc, pad))
            got = ops.phase_inv_polar(y, mag, mode, method, off, sc, pad)
            assert got.shape == want.shape and got.dtype == torch.complex64
            a, b = torch.view_as_real(got), torch.view_as_real(want)
            if method == "central":
                assert torch.equal(a, b), (mode, method, pad)
            else:
                assert float(((a - b).abs() / mag[..., None]).max()) <= 4e-6, (mode, method, pad)
                assert torch.equal(a, torch.view_as_real(ops.phase_inv_polar(y, mag, mode, method, off, sc, pad)))
    one = ops.phase_inv_polar(y[0, 0], torch.ones(17, 96, device="cuda"), PHASE_IF, "backward")
    ref = ops.polar_to_complex(torch.ones(17, 96, device="cuda"), ops.phase_inv(y[0, 0], PHASE_IF, "backward"))
    assert float((torch.view_as_real(one) - torch.view_as_real(ref)).abs().max()) <= 4e-6


# ---------------------------------------------------------------------------------------------
# MFCC (= MelSpectrogram) and the DCT variant
# ---------------------------------------------------------------------------------------------
def test_mfcc_dct_tensor_cores(ops):
    """The tcgen05 3xTF32 DCT against the FP32 kernel and a float64 reference (dB-scaled data, top_db floor)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    for B, n_mels, n_mfcc, T in [(3, 128, 40, 300), (2, 64, 13, 129), (1, 128, 48, 128)]:
        mel = torch.rand((B, n_mels, T), generator=g, device="cuda") ** 4 * 50.0 + 1e-9
        k = torch.arange(n_mfcc, dtype=torch.float64)[None, :]
        n = torch.arange(n_mels, dtype=torch.float64)[:, None]
        dct = torch.cos(math.pi / n_mels * (n + 0.5) * k) * math.sqrt(2.0 / n_mels)
        dct[:, 0] *= 1.0 / math.sqrt(2.0)
        dct = dct.float().cuda()
        ref32 = ops.mfcc_dct(mel, dct, 80.0, tensor_cores=False)
        got = ops.mfcc_dct(mel, dct, 80.0, tensor_cores=True)
        db = 10.0 * torch.log10(torch.clamp(mel.double(), min=1e-10))
        db = torch.maximum(db, db.amax() - 80.0)
        want = torch.einsum("bmt,mk->bkt", db, dct.double())
        assert_parity(host(got), host(want.float()), REL, "tc dct vs float64 (%d, %d, %d)" % (n_mels, n_mfcc, T))
        assert_parity(host(got), host(ref32), REL, "tc dct vs fp32 kernel")


def test_mel_projection_tensor_cores(ops):
    """The dense mel projection on tcgen05 (3xTF32, bank chunks by bulk asynchronous copy) against a float64 matmul, against
    the banded FP32 kernel on the same spectrum, and on the reference's own MelSpectrogram fixture; ragged frame counts,
    bin counts that are not a multiple of the 32-bin chunk, banks narrower than the 128-column tile."""
    from acids_transforms_b200.transforms.spectral_repr import melscale_fbanks
    g = torch.Generator(device="cuda").manual_seed(17)
    for B, T, n_fft, n_mels in [(3, 300, 2048, 128), (2, 129, 1024, 128), (1, 128, 512, 64), (2, 77, 256, 40), (1, 1, 64, 8)]:
        F = n_fft // 2 + 1
        spec = torch.rand((B, T, F), generator=g, device="cuda") ** 4 * 100.0
        fb = melscale_fbanks(F, 0.0, 22050.0, n_mels, 44100).cuda()
        got = ops.mel_tc(spec, fb)
        assert tuple(got.shape) == (B, n_mels, T)
        want = torch.einsum("btf,fm->bmt", spec.double(), fb.double())
        assert_parity(host(got), host(want.float()), 1e-5, "tc mel vs float64 (n_fft %d, %d mels, %d frames)" % (n_fft, n_mels, T))
        banded = ops.mag_epilogue(torch.complex(spec, torch.zeros_like(spec)), ops.BandedMatrix(fb.cpu()), None, 1e-7, None, None, False)
        assert_parity(host(got), host(banded.transpose(-1, -2)), 1e-5, "tc mel vs the banded FP32 kernel")
    # a dense random matrix (nothing banded about it) and the error message of a mismatched bank
    spec = torch.rand((2, 200, 513), generator=g, device="cuda")
    w = torch.randn((513, 96), generator=g, device="cuda")
    want = torch.einsum("btf,fm->bmt", spec.double(), w.double())
    # signed weights cancel: the 3xTF32 error is ~2^-22 of sum |a||b|, i.e. of the peak, not of each (small) element
    assert_parity(host(ops.mel_tc(spec, w)), host(want.float()), REL, "tc projection, dense random matrix")
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        ops.mel_tc(spec, w[:512])
    # the reference's MelSpectrogram (golden "mfcc": n_fft 2048-class fixture): power spectrum from the STFT kernel, then the GEMM
    gm = load_golden("mfcc")
    if "y" in gm and "x" in gm:
        x = cu(gm["x"])
        n_mels, Tm = gm["y"].shape[-2], gm["y"].shape[-1]
        L = x.shape[-1]
        hop = L // (Tm - 1)
        for n_fft in (4 * hop,):
            X = ops.stft_fwd(x.reshape(-1, L), torch.hann_window(n_fft).cuda(), n_fft, hop, True)
            fb = melscale_fbanks(n_fft // 2 + 1, 0.0, 22050.0, n_mels, 44100).cuda()
            got = ops.mel_tc(X.abs() ** 2, fb).reshape(gm["y"].shape)
            assert_parity(host(got), gm["y"], REL, "tc mel vs the reference's MelSpectrogram")


def test_griffinlim_update(ops):
    """The fused fast-Griffin-Lim update against the eager formula (torchaudio functional.py:336-350)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape in [(3, 33, 513), (1, 7, 17)]:          # even and odd bin counts
        reb = torch.randn(shape, generator=g, device="cuda", dtype=torch.complex64)
        tpr = torch.randn(shape, generator=g, device="cuda", dtype=torch.complex64)
        mag = torch.rand(shape, generator=g, device="cuda")
        a = reb - tpr * (0.99 / 1.99)
        want = mag * (a / (a.abs() + 1e-16))
        got = ops.griffinlim_update(reb, tpr, mag, 0.99)
        assert_parity(host(torch.view_as_real(got)), host(torch.view_as_real(want)), 1e-5, "griffin-lim update %s" % (shape,))


def test_mfcc_golden(ops):
    g = load_golden("mfcc")
    fb = dense(g["fb_rows"], g["fb_cols"], g["fb_vals"], g["fb_shape"])
    band = ops.BandedMatrix(torch.from_numpy(fb))
    y = ops.melspec_fwd(cu(g["x"]), torch.hann_window(2048).cuda(), 2048, 512, band, 2.0)
    assert tuple(y.shape) == g["y"].shape          # frequency-major [B, n_mels, T]
    assert_parity(host(y), g["y"], REL, "melspec 2048")
    d = ops.mfcc_dct(y, cu(g["dct"]), 80.0)
    assert_parity(host(d), g["y_dct40"], REL, "mfcc dct40")
    fb1k = O.melscale_fbanks(513, 0.0, 22050.0, 128, 44100)
    y2 = ops.melspec_fwd(cu(g["x2"]), torch.hann_window(1024).cuda(), 1024, 256, ops.BandedMatrix(torch.from_numpy(fb1k)), 2.0,
                         cu(g["y2_gauss_offset"].reshape(1)), cu(g["y2_gauss_scale"].reshape(1)))
    assert_parity(host(y2), g["y2_gauss"], 3e-4, "melspec gaussian (oracle-built bank)")
    y1 = ops.melspec_fwd(cu(g["x"]), torch.hann_window(2048).cuda(), 2048, 512, band, 1.0)
    assert_parity(host(y1), O.mel_spectrogram(g["x"], 44100, 2048, 512, 128, power=1.0, fb=fb), REL, "melspec power 1")


# ---------------------------------------------------------------------------------------------
# (5) mu-law / one-hot: bit-exact
# ---------------------------------------------------------------------------------------------
def test_mulaw_golden(ops):
    g = load_golden("mulaw")
    x = cu(g["x"])
    # CPU-eager semantics (IEEE division): compare with the reference's CPU output
    q = host(ops.mulaw_encode(x, 256, reciprocal_divide=False))
    assert q.dtype == np.int64
    mism = int((q != g["q"]).sum())
    assert mism <= 2 and (np.abs(q - g["q"]).max() <= 1), "mu-law vs CPU reference: %d mismatches" % mism
    # same-device eager chain (what the reference computes once moved .to('cuda')): must be bit-exact
    import torchaudio
    ref_cuda = torchaudio.functional.mu_law_encoding(x, 256)
    got = ops.mulaw_encode(x, 256)
    assert torch.equal(got, ref_cuda), "mu-law vs CUDA eager: %d mismatches" % int((got != ref_cuda).sum())
    assert host(ops.mulaw_encode(torch.tensor([-1.0, 0.0, 1.0]).cuda())).tolist() == [0, 128, 255]
    dec = ops.mulaw_decode(cu(g["q"]), 256)
    assert_parity(host(dec), g["dec"], 1e-6, "decode")
    assert torch.equal(dec, torchaudio.functional.mu_law_decoding(cu(g["q"]), 256))
    assert torch.equal(ops.mulaw_encode(x, 64), torchaudio.functional.mu_law_encoding(x, 64))
    assert_parity(host(ops.mulaw_decode(cu(g["q_c64"]), 64)), g["dec_c64"], 1e-6, "decode 64")
    xs = x[:, :64].contiguous()
    qs = ops.mulaw_encode(xs, 256, reciprocal_divide=False)
    assert np.array_equal(host(ops.mulaw_encode(xs, 256, "channel", reciprocal_divide=False)), O.one_hot(host(qs), 256, "channel"))
    assert np.array_equal(host(ops.mulaw_encode(xs, 256, "categorical", reciprocal_divide=False)), O.one_hot(host(qs), 256))
    assert np.array_equal(host(ops.one_hot(cu(g["q"][:, :64]), 256)), g["onehot64"])
    # odd class count exercises the scalar row writer
    assert np.array_equal(host(ops.mulaw_encode(xs, 255, "categorical", reciprocal_divide=False)),
                          O.one_hot(host(ops.mulaw_encode(xs, 255, reciprocal_divide=False)), 255))


def test_raw_golden(ops):
    g = load_golden("raw")
    x = cu(g["x"])
    assert_parity(host(ops.mono_mix(x)), g["mono"], 1e-6, "mono")
    assert_parity(host(ops.midside(x)), g["midside"], 1e-6, "midside")
    assert_parity(host(ops.midside(cu(g["midside"]), inverse=True)), g["midside_inv"], 1e-6, "midside inv")
    assert_parity(host(ops.midside(x, pad_mid=False)), g["midside_nopad"], 1e-6, "midside nopad")
    # bit-exact against the reference's arithmetic (raw.py:155-178) evaluated by torch on the CPU (IEEE division; torch's
    # CUDA kernel multiplies by the reciprocal of a scalar divisor instead): vector (L % 4 == 0) and scalar (odd L) routes
    gen = torch.Generator(device="cuda").manual_seed(2)
    for shape in ((3, 2, 4096), (5, 7, 2, 1001), (1, 2, 3), (70, 2, 260)):
        z = 2 * torch.rand(shape, generator=gen, device="cuda") - 1
        zc = z.cpu()
        left, right = zc[..., 0, :], zc[..., 1, :]
        mid, side = (left + right) / 2, (left - right) / 2
        want = torch.stack([mid / math.sqrt(2), side], -2)
        got = ops.midside(z)
        assert torch.equal(got.cpu(), want), shape
        assert torch.equal(ops.mono_mix(z).cpu(), zc.sum(-2) / 2), shape          # raw.py:39
        back = ops.midside(got, inverse=True).cpu()
        m2 = want[..., 0, :] * math.sqrt(2)
        assert torch.equal(back, torch.stack([m2 + want[..., 1, :], m2 - want[..., 1, :]], -2)), shape


def test_normalize_stats_golden(ops):
    g = load_golden("normalize")
    mn, mx, mean, std = host(ops.stats(cu(g["x"])))
    assert abs(mn - float(g["unipolar_offset"])) < 1e-6 and abs((mx - mn) - float(g["unipolar_scale"])) < 1e-5
    assert abs(mean - float(g["gaussian_offset"])) < 1e-5 and abs(std - float(g["gaussian_scale"])) < 1e-5


# ---------------------------------------------------------------------------------------------
# streaming: pre-framed rfft / irfft and overlap-add with carry
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["rtstft", "rtdgt"])
def test_streaming_golden(ops, tag):
    g = load_golden("stream_" + tag)
    w, iw = cu(g["window"]), cu(g["inv_window"])
    carry = None
    for i in range(3):
        fr = cu(g["frames_%d" % i])
        X = ops.stft_fwd(fr, w, 512, 128, center=False)
        assert_parity(host(X), g["X_%d" % i], REL, "rt fwd")
        yi = ops.irfft_frames(cu(g["X_%d" % i]), iw, 512)
        assert_parity(host(yi), g["inv_frames_%d" % i], REL, "rt inv")
        out, carry = ops.ola_stream(cu(g["inv_frames_%d" % i]), 128, 3 * 128, carry, float(g["gain"]))
        assert_parity(host(out), g["out_%d" % i], REL, "oadd inv")


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties on one full-size shard
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(ops):
    """cfg 2 geometry (4 s clips, n_fft 1024, hop 256) on 64 clips: the oracle on a subset, round trip,
    linearity and batch-independence on the whole batch."""
    B, L, n, h = 64, 176400, 1024, 256
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = 0.5 * (2 * torch.rand((B, L), generator=gen, device="cuda") - 1)
    w = torch.hann_window(n).cuda()
    X = ops.stft_fwd(x, w, n, h)
    assert tuple(X.shape) == (B, 690, 513)
    Xo = O.stft(host(x[:2]), n, h, O.periodic_window("hann", n))
    assert_parity(host(X[:2]), Xo, REL, "full-size fwd subset")
    y = ops.istft_ola(X, w, n, h)
    assert tuple(y.shape) == (B, 256 * 689)
    err = float((y - x[:, :y.shape[1]]).abs().max())
    assert err <= REL * float(x.abs().max()), "round trip %.3e" % err
    # linearity: STFT(a x1 + b x2) = a STFT(x1) + b STFT(x2)
    X2 = ops.stft_fwd(2.0 * x[:32] - 0.5 * x[32:], w, n, h)
    lin = 2.0 * X[:32] - 0.5 * X[32:]
    assert float((X2 - lin).abs().max()) <= REL * float(lin.abs().max())
    # a clip's result does not depend on its batch position / batch size
    assert torch.equal(ops.stft_fwd(x[7:8], w, n, h)[0], X[7])
    # fused forward == unfused forward on the whole batch
    fwd, _ = O.magnitude_banks(44100, n)
    band = ops.BandedMatrix(torch.from_numpy(fwd))
    yf = ops.stft_mag_fwd(x, w, n, h, band, "log1p", EPS, 0.1, 2.0)
    yu = ops.mag_epilogue(X, band, "log1p", EPS, 0.1, 2.0)
    assert float((yf - yu).abs().max()) <= REL * float(yu.abs().max())
    assert_parity(host(yf[:2]), O.magnitude_forward(Xo, fwd, "log1p", offset=0.1, scale=2.0), REL, "full-size fused subset")
    # run-to-run determinism, bit for bit, on a batch large enough to fill the persistent grids several times: a
    # shared-memory race (exchange buffers, the double-buffered |X| rows, the frame ring) would show up as noise;
    # and every clip's result equals the same clip processed alone (independent of which CTA / segment got it)
    xb = x.repeat(4, 1)
    y1 = ops.stft_mag_fwd(xb, w, n, h, band, "log1p", EPS, 0.1, 2.0)
    for _ in range(3):
        assert torch.equal(ops.stft_mag_fwd(xb, w, n, h, band, "log1p", EPS, 0.1, 2.0), y1)
    assert torch.equal(y1[:B], y1[3 * B:]) and torch.equal(y1[:B], yf)
    Xb = ops.stft_fwd(xb, w, n, h)
    assert torch.equal(Xb[B:2 * B], X)
    z1 = ops.istft_ola(Xb, w, n, h)
    for _ in range(3):
        assert torch.equal(ops.istft_ola(Xb, w, n, h), z1)
    assert torch.equal(z1[:B], z1[2 * B:3 * B]) and torch.equal(z1[:B], y)


def test_host_tensor_roundtrip(ops):
    """CPU tensors are staged to the GPU and back (what the reference's own CPU-tensor tests would feed)."""
    x = torch.from_numpy(synth(2, 4096, 5))
    w = torch.hann_window(512)
    X = ops.stft_fwd(x, w, 512, 128)
    assert X.device.type == "cpu"
    assert_parity(X.numpy(), O.stft(x.numpy(), 512, 128, w.numpy()), REL, "host in/out")
