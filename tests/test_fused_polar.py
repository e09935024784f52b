"""[MidSide +] STFT + Polar / PolarIF as ONE kernel (acids_stft_polar_fwd, VERDICT r1 "missing" #1):
the fused stage of ComposeAudioTransform against the children run one after the other, and against the
golden vectors minted from the unmodified reference (raw.py:145-162, stft.py:101-102, spectral_repr.py:431-440).

The fused phase half evaluates the forward-difference IF as the wrapped difference of two consecutive raw phases;
the reference differences the unwrapped phases (a float32 running sum over the frames), so the two agree up to the
rounding of that sum — far inside the 1e-4 budget — except where a raw phase sits on the +-pi branch cut (masked
the way the golden tests mask it, conftest.if_mask)."""
import numpy as np
import pytest
import torch

from conftest import assert_parity, branch_cut, if_mask, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    assert torch.cuda.is_available()
    from acids_transforms_b200 import transforms, _lib
    _lib.load()
    return transforms


def host(t):
    return t.detach().cpu().resolve_conj().numpy()


def _fit(ch, x):
    ch.scale_data(x)
    for st in ch._plan:                      # the one-kernel path is opt-in (fused.FusedSTFTPolar): switch it on
        if type(st).__name__ == "FusedSTFTPolar":
            st.one_kernel = True
    return ch


def wrap_ambiguous(ref, norm, weighted=False):
    """IF rows whose phase advance is +-pi to within rounding: the value is +0.5 or -0.5 by the last bit of the two raw
    phases (the same frequency either way) — not comparable between two evaluations of the spectrum."""
    v = ref * float(norm.scale) + float(norm.offset)
    T = ref.shape[-2]
    if weighted:
        n = np.arange(T, dtype=np.float32)
        a = (n - (T / 2.0 - 1.0)) / (T / 2.0)
        w = (1.5 * T) / (T * T - 1.0) * (1.0 - a * a)
        v = v / np.where(w == 0, 1, w)[:, None]
    amb = np.abs(np.abs(v) - 0.5) < 1e-4
    amb[..., 0, :] = False
    amb[..., T - 1, :] = np.abs(np.abs(v[..., T - 1, :]) - 0.5 * np.pi) < 3e-4         # the last row is not divided by pi
    return amb


def assert_phase_close(a, b, X, rad_per_unit, diff, what):
    """Phase-like outputs of two evaluations of the same spectrum.  The two FFTs agree to ~3e-7 of the spectrum's peak
    (different fusion / contraction of the same arithmetic); the phase of a bin of modulus |X| then moves by up to that
    over |X|, so the budget is 1e-5 * pi (conftest.assert_parity's atol for a +-pi quantity) plus 2e-6 peak / |X| per
    frame involved (`diff`: the value differences frames t and t - 1).  Branch-cut bins are masked by the caller."""
    absX = np.abs(X)
    weak = absX < 1e-3 * absX.max()           # +-pi flips of a (nearly) real bin that conftest.branch_cut's 1e-4 misses
    if diff:
        weak[..., 1:, :] |= weak[..., :-1, :].copy()
    a, b = np.where(weak, 0, a), np.where(weak, 0, b)
    per = 2e-6 * absX.max() / np.maximum(absX, 1e-30)
    allow = 1e-5 * np.pi + per
    if diff:
        allow[..., 1:, :] += per[..., :-1, :]
    err = np.abs(a - b) * rad_per_unit
    bad = err > allow
    assert not bad.any(), "%s: %d of %d elements beyond the phase budget, worst %.3e rad (allowed %.3e)" % (
        what, int(bad.sum()), bad.size, float(err[bad].max()), float(allow[bad][err[bad].argmax()]))


@pytest.mark.parametrize("n_fft,hop", [(256, 64), (512, 128), (1024, 256), (2048, 512), (4096, 1024), (8192, 2048)])
@pytest.mark.parametrize("kind", ["polar", "polarif", "polarif_weighted"])
def test_fused_equals_children(T, n_fft, hop, kind):
    torch.manual_seed(n_fft + len(kind))
    B, L = (40, 6 * n_fft + 3 * hop) if n_fft <= 1024 else (6, 5 * n_fft + hop)
    x = torch.randn(B, L, device="cuda")
    margs = {"mode": "bipolar", "n_fft": n_fft}
    if kind == "polar":
        rep = T.Polar(magnitude_args=margs)
    else:
        rep = T.PolarIF(magnitude_args=margs, phase_args={"mode": "bipolar", "weighted": kind.endswith("weighted")})
    ch = _fit((T.STFT(n_fft=n_fft, hop_length=hop) + rep).cuda(), x)
    assert type(ch._plan[0]).__name__ == "FusedSTFTPolar"
    y = ch(x)
    ref = ch.forward_unfused(x)
    assert y.shape == ref.shape == (B, 1 + L // hop, 2, n_fft // 2 + 1)
    assert_parity(host(y[..., 0, :]), host(ref[..., 0, :]), 1e-5, "magnitude slot")
    X = host(ch[0](x))
    ok = if_mask(X) if kind != "polar" else ~branch_cut(X)
    if kind != "polar":
        ok &= ~wrap_ambiguous(host(ref[..., 1, :]), ch[1].phase.norm, kind.endswith("weighted"))
    assert ok.mean() > 0.9
    unit = float(ch[1].phase.norm.scale) * (np.pi if kind != "polar" else 1.0)      # IF rows are phase differences / (2) pi
    assert_phase_close(np.where(ok, host(y[..., 1, :]), 0), np.where(ok, host(ref[..., 1, :]), 0), X, unit, kind != "polar", "phase slot")
    # scripted chain, same kernel
    ys = torch.jit.script(ch)(x)
    assert torch.equal(ys, y)


def test_fused_midside_and_options(T):
    """MidSide folded into the sample loads (both pad_mid settings), keep_nyquist=False (drops bin 0), mel=False."""
    torch.manual_seed(7)
    x = torch.randn(5, 2, 20000, device="cuda")
    for pad_mid in (True, False):
        for keep in (True, False):
            for mel in (True, False):
                rep = T.PolarIF(magnitude_args={"mode": "bipolar", "n_fft": 1024, "mel": mel}, keep_nyquist=keep)
                ch = _fit((T.MidSide(pad_mid=pad_mid) + T.STFT(n_fft=1024, hop_length=256) + rep).cuda(), x)
                assert type(ch._plan[0]).__name__ == "FusedSTFTPolar" and ch._plan[0].has_midside
                y, ref = ch(x), ch.forward_unfused(x)
                assert y.shape == ref.shape == (5, 2, 79, 2, 513 - (0 if keep else 1))
                assert_parity(host(y[..., 0, :]), host(ref[..., 0, :]), 1e-5, "magnitude slot")
                X = host(ch[1](ch[0](x)))[..., (0 if keep else 1):]
                ok = if_mask(X) & ~wrap_ambiguous(host(ref[..., 1, :]), ch[2].phase.norm)
                assert_phase_close(np.where(ok, host(y[..., 1, :]), 0), np.where(ok, host(ref[..., 1, :]), 0), X,
                                   float(ch[2].phase.norm.scale) * np.pi, True, "IF slot")
                xi = ch.invert(y)
                assert xi.shape == (5, 2, 256 * 78)


def test_fused_falls_back(T):
    """Configurations the one-kernel path does not cover run the children in turn with identical results:
    unwrapped phase, backward / central IF, mono input through MidSide, MidSide(normalize=True)."""
    torch.manual_seed(11)
    x = torch.randn(3, 2, 9000, device="cuda")
    for rep in (T.Polar(phase_args={"mode": "bipolar", "unwrap": True}), T.PolarIF(phase_args={"mode": "bipolar", "method": "backward"}),
                T.PolarIF(phase_args={"mode": "bipolar", "method": "central"})):
        # MidSide rides in the STFT's sample loads (acids_midside_stft_fwd): mid = (l + r) * (1 / (2 sqrt 2)) there, a true
        # division by sqrt 2 in the stand-alone kernel — one ulp of the waveform apart
        ch = _fit((T.MidSide() + T.STFT(n_fft=1024, hop_length=256) + rep).cuda(), x)
        y, ref = ch(x), ch.forward_unfused(x)
        assert_parity(host(y[..., 0, :]), host(ref[..., 0, :]), 1e-5, "magnitude slot, folded MidSide")
        chp = _fit((T.MidSide(pad_mid=False) + T.STFT(n_fft=1024, hop_length=256) + rep).cuda(), x)
        assert torch.equal(chp(x), chp.forward_unfused(x))          # without the sqrt 2 the arithmetic is the same: bit-identical
    ch = _fit((T.MidSide(normalize=True) + T.STFT(n_fft=1024, hop_length=256) + T.PolarIF()).cuda(), x)
    assert torch.equal(ch(x), ch.forward_unfused(x))
    mono = torch.randn(3, 1, 9000, device="cuda")
    ch = _fit((T.MidSide() + T.STFT(n_fft=1024, hop_length=256) + T.PolarIF()).cuda(), mono)
    assert torch.equal(ch(mono), ch.forward_unfused(mono))


def test_fused_partition_independent(T):
    """A batch that gives every CTA of the persistent grid a run starting inside a clip (the one-unit halo that
    supplies the previous frame's phase row): clip b of the big batch equals clip b transformed alone, bit for bit."""
    torch.manual_seed(3)
    x = torch.randn(700, 16384 + 256, device="cuda")
    ch = _fit((T.STFT(n_fft=1024, hop_length=256) + T.PolarIF()).cuda(), x[:4])
    y = ch(x)
    for b in (0, 1, 123, 350, 699):
        assert torch.equal(y[b], ch(x[b:b + 1])[0])
    x4 = torch.randn(300, 2, 4096 * 6, device="cuda")
    ch4 = _fit((T.MidSide() + T.STFT(n_fft=4096, hop_length=1024) + T.PolarIF(magnitude_args={"mode": "bipolar", "n_fft": 4096})).cuda(), x4[:2])
    y4 = ch4(x4)
    for b in (0, 77, 299):
        assert torch.equal(y4[b], ch4(x4[b:b + 1])[0])


def test_cfg4_full_size_chain(T):
    """cfg 4 at BASELINE size per clip (stereo 4 s @ 44.1 kHz, n_fft 4096, hop 1024): the fused chain against the
    children run in turn, and the round trip through the inverse chain (PolarIF.invert -> ISTFT -> MidSide.invert)."""
    torch.manual_seed(5)
    L = 176400
    x = 0.5 * (2 * torch.rand(16, 2, L, device="cuda") - 1)
    ch = _fit((T.MidSide() + T.STFT(n_fft=4096, hop_length=1024) + T.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})).cuda(), x[:4])
    y, ref = ch(x), ch.forward_unfused(x)
    assert y.shape == (16, 2, 173, 2, 2049)
    assert_parity(host(y[..., 0, :]), host(ref[..., 0, :]), 1e-5, "cfg4 magnitude slot")
    X = host(ch[1](ch[0](x)))
    ok = if_mask(X) & ~wrap_ambiguous(host(ref[..., 1, :]), ch[2].phase.norm)
    assert ok.mean() > 0.95
    assert_phase_close(np.where(ok, host(y[..., 1, :]), 0), np.where(ok, host(ref[..., 1, :]), 0), X,
                       float(ch[2].phase.norm.scale) * np.pi, True, "cfg4 IF slot")
    # round trip: the mel bank's inverse is a row-normalised transpose (not an inverse), so compare the two inverse
    # chains with each other rather than with x; both must reproduce the fused and the unfused forward alike
    xi, xr = ch.invert(y), ch.invert(ref)
    assert xi.shape == (16, 2, 1024 * 172)
    assert_parity(host(xi), host(xr), 2e-3, "cfg4 inverse of fused vs unfused forward")


def test_cfg4_golden_through_the_c_abi():
    """acids_stft_polar_fwd against the golden cfg-4 chain of the unmodified reference, called through ops (ctypes)."""
    from acids_transforms_b200 import ops, transforms as Tr
    g = load_golden("chain_cfg4")
    x = torch.from_numpy(g["x"]).cuda()
    mag = Tr.Magnitude(n_fft=4096, mode="bipolar").cuda()
    w = Tr.STFT(n_fft=4096, hop_length=1024).window.cuda()
    y = ops.stft_polar_fwd(x, w, 4096, 1024, ops.as_band(mag.mel_meta, mag.mel_coef), "log1p", mag._eps, torch.from_numpy(g["mag_offset"]), torch.from_numpy(g["mag_scale"]), 2, 0, False,
                           torch.from_numpy(g["ph_offset"]), torch.from_numpy(g["ph_scale"]), False, midside=2)
    assert tuple(y.shape) == g["y"].shape
    assert_parity(host(y[..., 0, :]), g["y"][..., 0, :], 1e-4, "cfg4 golden magnitude")
    X = host(ops.stft_fwd(ops.midside(x), w, 4096, 1024))
    from types import SimpleNamespace
    ok = if_mask(X) & ~wrap_ambiguous(g["y"][..., 1, :], SimpleNamespace(scale=g["ph_scale"], offset=g["ph_offset"]))
    assert_parity(np.where(ok, host(y[..., 1, :]), 0), np.where(ok, g["y"][..., 1, :], 0), 2e-4, "cfg4 golden IF")


# ---------------------------------------------------------------------------------------------
# scale_data statistics from the forward kernel's statistics mode
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_fft,hop", [(64, 16), (1024, 256), (2048, 512), (4096, 1024), (16384, 4096)])
@pytest.mark.parametrize("contrast", ["log1p", "log", None])
def test_stft_stats_equals_stats_of_the_spectrum(n_fft, hop, contrast):
    """acids_stft_stats (no spectrum in HBM) against acids_stats on the materialised spectrum and against float64 numpy."""
    from acids_transforms_b200 import ops
    torch.manual_seed(n_fft)
    B, L = (33, 5 * n_fft + 7 * hop) if n_fft <= 4096 else (3, 3 * n_fft)
    x = torch.randn(B, L, device="cuda")
    w = torch.hann_window(n_fft, device="cuda")
    st = ops.stft_stats(x, w, n_fft, hop, contrast, 1e-6).cpu().numpy()
    X = ops.stft_fwd(x, w, n_fft, hop)
    ref = ops.stats(X, contrast, 1e-6).cpu().numpy()
    a = np.abs(host(X)).astype(np.float64)
    v = {"log1p": np.log1p(a), "log": np.log(np.maximum(a, 1e-6)), None: a}[contrast]
    want = np.array([v.min(), v.max(), v.mean(), v.std(ddof=1)])
    span = want[1] - want[0]
    tol = lambda r: 2e-5 * span + 1e-4 * np.abs(r) * (contrast != "log")
    lo = 0
    if contrast == "log":
        # the minimum of log(|X|) is the logarithm of the bin closest to zero, i.e. of the FFT's rounding noise on it
        # (|X| ~ 1e-5 of the peak): only bounded here — not below the clamp, within a factor e^1.5 of the other evaluations
        lo = 1
        assert st[0] >= np.log(1e-6) - 1e-4 and abs(st[0] - want[0]) <= 1.5 and abs(st[0] - ref[0]) <= 1.5, (st, want, ref)
    assert np.all(np.abs(st - want)[lo:] <= tol(want)[lo:]), (st, want)
    assert np.all(np.abs(st - ref)[lo:] <= tol(ref)[lo:]), (st, ref)


def test_fused_chain_scale_data_is_one_pass(T):
    """ComposeAudioTransform.scale_data on STFT|DGT + Magnitude: the fitted offset / scale equal the unfused fit
    (golden cfg 2 values from the reference) and the spectrum is never allocated."""
    g = load_golden("chain_cfg2")
    x = torch.from_numpy(g["x"]).cuda()
    ch = (T.DGT(sr=44100, n_fft=1024, hop_length=256, inversion_mode="random") + T.Magnitude(mel=True, mode="unipolar", contrast="log1p")).cuda()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    ch._plan[0].scale_data(x)
    peak = torch.cuda.max_memory_allocated() - base
    spectrum_bytes = x.shape[0] * (1 + x.shape[1] // 256) * 513 * 8
    assert peak < spectrum_bytes // 2, (peak, spectrum_bytes)
    assert abs(float(ch[1].norm.offset) - float(g["offset"])) <= 1e-5
    assert abs(float(ch[1].norm.scale) - float(g["scale"])) <= 1e-4 * float(g["scale"])
    for mode in ("bipolar", "gaussian"):
        a = (T.STFT(n_fft=1024, hop_length=256) + T.Magnitude(mode=mode)).cuda()
        a.scale_data(x)
        b = T.Magnitude(mode=mode).cuda()
        b.scale_data(a[0](x))
        assert abs(float(a[1].norm.offset) - float(b.norm.offset)) <= 1e-5 * max(1.0, abs(float(b.norm.offset)))
        assert abs(float(a[1].norm.scale) - float(b.norm.scale)) <= 1e-4 * float(b.norm.scale)
    sc = torch.jit.script(ch)
    sc.scale_data(x)
    assert_parity(host(sc(x)), g["y"], 1e-4, "cfg2 scripted after fused scale_data")


def test_midside_folded_into_the_stft(T):
    """acids_midside_stft_fwd (MidSide.forward inside the STFT's sample loads, raw.py:145-161 + stft.py:101-102) against the
    two modules run in turn, at n_fft 1024 and 4096, both pad_mid settings, through torch.ops and through the chain plan."""
    torch.manual_seed(9)
    x = 0.5 * (2 * torch.rand(3, 2, 40000, device="cuda") - 1)
    for n_fft, hop in ((1024, 256), (4096, 1024)):
        for pad_mid in (False, True):
            ms, st = T.MidSide(pad_mid=pad_mid).cuda(), T.STFT(n_fft=n_fft, hop_length=hop).cuda()
            want = st(ms(x))
            got = torch.ops.acids_b200.midside_stft_fwd(x, st.window, n_fft, hop, 2 if pad_mid else 1)
            assert got.shape == want.shape
            assert_parity(host(torch.view_as_real(got)), host(torch.view_as_real(want)), 1e-5, "midside+stft %d pad_mid=%s" % (n_fft, pad_mid))
    with pytest.raises(RuntimeError):
        torch.ops.acids_b200.midside_stft_fwd(x[:, 0], T.STFT().cuda().window, 1024, 256, 1)      # not stereo


@pytest.mark.parametrize("n_fft,hop", [(64, 16), (512, 128), (1024, 256), (2048, 512), (4096, 1024), (8192, 2048)])
@pytest.mark.parametrize("kind", ["polar", "polarif", "polarif_weighted", "polarif_drop"])
def test_polar_rows_against_the_two_kernels(T, n_fft, hop, kind):
    """acids_polar_rows_fwd (mel bank + raw phase / forward IF from ONE read of the spectrum; what Polar.forward runs on long rows)
    against acids_mag_epilogue + acids_phase_fwd writing the two slots.  Enough clips that the persistent grid hands most
    CTAs a run starting inside a clip (the recomputed carry row), odd frame counts so that row tiles straddle clips."""
    from acids_transforms_b200 import ops
    torch.manual_seed(n_fft + len(kind))
    B, L = (300, 5 * n_fft + 3 * hop) if n_fft <= 1024 else (37, 5 * n_fft + hop)
    x = torch.randn(B, L, device="cuda")
    keep = kind != "polarif_drop"
    margs = {"mode": "bipolar", "n_fft": n_fft}
    if kind == "polar":
        rep = T.Polar(magnitude_args=margs, keep_nyquist=keep).cuda()
    else:
        rep = T.PolarIF(magnitude_args=margs, phase_args={"mode": "bipolar", "weighted": kind.endswith("weighted")}, keep_nyquist=keep).cuda()
    st = T.STFT(n_fft=n_fft, hop_length=hop).cuda()
    X = st(x)
    rep.scale_data(X[:8])
    mag, ph = rep.magnitude, rep.phase

    def rows(Xi):      # the entry point itself: the modules only dispatch to it where it is the faster path (ops.polar_rows_pays)
        return ops.polar_rows_fwd(Xi, ops.as_band(mag.band_meta(), mag.band_coef()), mag.contrast_mode, mag._eps, mag.norm.get_offset(),
                                  mag.norm.get_scale(), rep._phase_mode(), rep._phase_method(), rep._phase_weighted(), ph.norm.get_offset(),
                                  ph.norm.get_scale(), not keep)

    y = rows(X)
    assert (kind == "polar" and n_fft >= 2048) == (type(rep).__name__ == "Polar" and ops.polar_rows_pays(rep._phase_mode(), n_fft // 2 + 1))
    Fk = n_fft // 2 + 1 - (0 if keep else 1)
    ref = torch.empty((B, X.shape[1], 2, Fk), device="cuda")
    ops.mag_epilogue(X, ops.as_band(mag.band_meta(), mag.band_coef()), mag.contrast_mode, mag._eps, mag.norm.get_offset(), mag.norm.get_scale(),
                     not keep, out=ref, out_slot=0, out_slots=2)
    ops.phase_fwd(X, rep._phase_mode(), rep._phase_method(), rep._phase_weighted(), ph.norm.get_offset(), ph.norm.get_scale(), not keep,
                  out=ref, out_slot=1, out_slots=2)
    assert y.shape == ref.shape
    assert_parity(host(y[..., 0, :]), host(ref[..., 0, :]), 1e-6, "magnitude slot")
    Xh = host(X)[..., (0 if keep else 1):]
    if kind == "polar":
        assert torch.equal(y[..., 1, :], ref[..., 1, :])                  # the same arctangent of the same bins
        return
    ok = if_mask(Xh) & ~wrap_ambiguous(host(ref[..., 1, :]), ph.norm, kind.endswith("weighted"))
    assert ok.mean() > 0.9
    assert_phase_close(np.where(ok, host(y[..., 1, :]), 0), np.where(ok, host(ref[..., 1, :]), 0), Xh, float(ph.norm.scale) * np.pi, True, "IF slot")
    # position independence: a clip alone (its run starts at frame 0) equals the clip inside the batch, bit for bit
    for b in (1, B // 2, B - 1):
        assert torch.equal(rows(X[b:b + 1])[0], y[b])
