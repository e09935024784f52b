"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is imported as is; the only shim is a stub ``turtle`` module, because
``acids_transforms/transforms/misc.py:1`` does ``from turtle import forward`` and the
image has no tkinter.  All tensors are float32 / complex64 / int64 on
CPU (torch 2.11.0, torchaudio 2.11.0).  Outputs: ``tests/golden/*.npz``.
"""
import os
import sys
import types
import math
import numpy as np

REF = os.environ.get("ACIDS_REFERENCE", "/root/reference")
sys.modules.setdefault("turtle", types.ModuleType("turtle"))
sys.modules["turtle"].forward = None
sys.path.insert(0, REF)

import torch  # noqa: E402
import torchaudio  # noqa: E402
import acids_transforms as ref  # noqa: E402
from acids_transforms import transforms as T  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SR = 44100


def synth(n_clips, length, seed, channels=None):
    """SURVEY.md §8(d) synthetic clips: 0.5*(2U-1) + 0.25*sin(2 pi f_i n / sr), f_i = 55 * 2^((i mod 84)/12)."""
    g = torch.Generator().manual_seed(seed)
    shape = (n_clips, length) if channels is None else (n_clips, channels, length)
    x = 0.5 * (2 * torch.rand(shape, generator=g) - 1)
    n = torch.arange(length, dtype=torch.float64)
    for i in range(n_clips):
        f = 55.0 * 2 ** ((i % 84) / 12)
        s = (0.25 * torch.sin(2 * math.pi * f * n / SR)).float()
        x[i] = x[i] + (s if channels is None else s * torch.tensor([1.0, 0.6]).view(2, 1)[:channels])
    return x.float()


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().resolve_conj().numpy()
    return np.asarray(t)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print("%-28s %8.1f KB  %s" % (name, os.path.getsize(path) / 1024, sorted(arrays)))


def sparse(m):
    """COO of a mostly-zero matrix (the square mel banks are 99.6 % zeros)."""
    m = npy(m)
    r, c = np.nonzero(m)
    return r.astype(np.int32), c.astype(np.int32), m[r, c], np.array(m.shape, np.int32)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)

    # ---- windows (A1, A2) -------------------------------------------------
    out = {}
    for n, h in ((256, 64), (512, 128), (1024, 256), (2048, 512), (4096, 1024)):
        s = T.STFT(n_fft=n, hop_length=h)
        out["hann_%d" % n] = s.window[:n]
        if n <= 2048:  # the dual-window python double loop is O(N * N/H)
            d = T.DGT(n_fft=n, hop_length=h)
            out["gauss_%d" % n] = d.window[:n]
            out["dual_%d_%d" % (n, h)] = d.inv_window[:n]
            out["gamma_%d" % n] = d.gamma
    for name in ("hamming", "blackman", "bartlett"):
        out[name + "_512"] = T.STFT(n_fft=512, hop_length=128, window=name).window[:512]
    save("windows", **out)

    # ---- STFT / DGT forward + complex inverse (A3, A4, A15) ----------------
    for n, h, L, seed in ((1024, 256, 8192, 11), (512, 128, 4096, 12), (2048, 512, 8192, 13),
                          (4096, 1024, 12288, 14), (256, 64, 4096, 15)):
        x = synth(2 if n < 4096 else 1, L, seed)
        s = T.STFT(n_fft=n, hop_length=h)
        X = s(x)
        y = s.invert(X)
        arrays = dict(x=x, X=X, y=y, n_fft=n, hop=h)
        if n == 1024:
            arrays["phase_buffer"] = s.phase_buffer
            Xt, tt = s.forward_with_time(x, torch.tensor([0.0, 1.5]))
            arrays["time_in"] = torch.tensor([0.0, 1.5])
            arrays["time_out"] = tt
        save("stft_%d_%d" % (n, h), **arrays)
    # ragged / short: L not a multiple of hop, L just above n_fft/2
    x = synth(3, 1000, 16)
    s = T.STFT(n_fft=512, hop_length=128)
    X = s(x)
    save("stft_ragged_512_128", x=x, X=X, y=s.invert(X), n_fft=512, hop=128)
    # batch dims [2, 2, L] keep their shape
    x = synth(2, 4096, 17, channels=2)
    s = T.STFT(n_fft=512, hop_length=128)
    X = s(x)
    save("stft_stereo_512_128", x=x, X=X, y=s.invert(X), n_fft=512, hop=128)

    for n, h, L, seed in ((1024, 256, 8192, 21), (512, 128, 4096, 22)):
        x = synth(2, L, seed)
        d = T.DGT(n_fft=n, hop_length=h)
        X = d(x)
        save("dgt_%d_%d" % (n, h), x=x, X=X, y=d.invert(X), n_fft=n, hop=h,
             window=d.window[:n], inv_window=d.inv_window[:n])

    # ---- Magnitude (A5-A8) -------------------------------------------------
    x = synth(2, 4096, 31)
    X = T.STFT(n_fft=1024, hop_length=256)(x)
    m = T.Magnitude()
    r, c, v, shp = sparse(m.mel_bank[0])
    ri, ci, vi, _ = sparse(m.inverse_mel_bank[0])
    save("mel_bank_1024", rows=r, cols=c, vals=v, shape=shp, inv_rows=ri, inv_cols=ci, inv_vals=vi)
    m512 = T.Magnitude(n_fft=512)
    r, c, v, shp = sparse(m512.mel_bank[0])
    ri, ci, vi, _ = sparse(m512.inverse_mel_bank[0])
    save("mel_bank_512", rows=r, cols=c, vals=v, shape=shp, inv_rows=ri, inv_cols=ci, inv_vals=vi)
    mk = T.Magnitude(keep_nyquist=False)
    r, c, v, shp = sparse(mk.mel_bank[0])
    ri, ci, vi, _ = sparse(mk.inverse_mel_bank[0])
    save("mel_bank_1024_nonyq", rows=r, cols=c, vals=v, shape=shp, inv_rows=ri, inv_cols=ci, inv_vals=vi)
    m4k = T.Magnitude(n_fft=4096)
    r, c, v, shp = sparse(m4k.mel_bank[0])
    ri, ci, vi, _ = sparse(m4k.inverse_mel_bank[0])
    save("mel_bank_4096", rows=r, cols=c, vals=v, shape=shp, inv_rows=ri, inv_cols=ci, inv_vals=vi)

    arrays = dict(X=X)
    for tag, kw in (("default", {}),
                    ("bipolar_log", dict(mode="bipolar", contrast="log")),
                    ("gauss_log10", dict(mode="gaussian", contrast="log10")),
                    ("none_none_nomel", dict(mode=None, contrast=None, mel=False)),
                    ("unipolar_log1p_nomel", dict(mode="unipolar", contrast="log1p", mel=False)),
                    ("nonyq", dict(keep_nyquist=False))):
        m = T.Magnitude(**kw)
        m.scale_data(X)
        y = m(X)
        arrays[tag + "_y"] = y
        arrays[tag + "_inv"] = m.invert(y.clone())
        if not isinstance(m.norm, T.spectral_repr.Dummy):
            arrays[tag + "_offset"] = m.norm.offset
            arrays[tag + "_scale"] = m.norm.scale
    save("magnitude_1024", **arrays)

    # ---- Phase / IF / Polar (A10-A14) --------------------------------------
    x = synth(1, 8192, 41)
    X = T.STFT(n_fft=256, hop_length=64)(x)          # T = 129 frames: long unwrap runs
    arrays = dict(X=X)
    for tag, kw in (("raw", {}), ("unwrap", dict(unwrap=True)),
                    ("unwrap_bipolar", dict(unwrap=True, mode="bipolar")),
                    ("nonyq", dict(keep_nyquist=False))):
        p = T.Phase(**kw)
        p.scale_data(X)
        y = p(X)
        arrays["phase_%s_y" % tag] = y
        arrays["phase_%s_inv" % tag] = p.invert(y.clone())
        if kw.get("mode"):
            arrays["phase_%s_offset" % tag] = p.norm.offset
            arrays["phase_%s_scale" % tag] = p.norm.scale
    for method in ("forward", "backward", "central"):
        f = T.IF(method=method)
        f.scale_data(X)
        y = f(X)
        tag = "if_%s" % method
        arrays[tag + "_y"] = y
        arrays[tag + "_offset"] = f.norm.offset
        arrays[tag + "_scale"] = f.norm.scale
        arrays[tag + "_inv"] = f.invert(y.clone())
        # weighted=True: the reference survives exactly ONE call per module (its cached
        # 1-D weighted_window then fails `.size(-2)`, spectral_repr.py:339), so capture a
        # single un-normalised call on a fresh module.
        arrays[tag + "_w_y"] = T.IF(method=method, weighted=True, mode=None)(X)
    f = T.IF(mode=None)
    arrays["if_forward_nonorm_y"] = f(X)
    save("phase_if_256", **arrays)

    x = synth(2, 4096, 42)
    X = T.STFT(n_fft=1024, hop_length=256)(x)
    arrays = dict(X=X)
    for tag, cls in (("polar", T.Polar), ("polarif", T.PolarIF)):
        p = cls()
        p.scale_data(X)
        y = p(X)
        arrays[tag + "_y"] = y
        arrays[tag + "_inv"] = p.invert(y.clone())
        arrays[tag + "_mag_offset"] = p.magnitude.norm.offset
        arrays[tag + "_mag_scale"] = p.magnitude.norm.scale
        arrays[tag + "_ph_offset"] = p.phase.norm.offset
        arrays[tag + "_ph_scale"] = p.phase.norm.scale
    save("polar_1024", **arrays)

    # ---- MFCC (= MelSpectrogram) and the DCT variant (A9) ------------------
    x = synth(2, 12288, 51)
    mf = T.MFCC(n_fft=2048, hop_length=512, n_mels=128)
    y = mf(x)
    tm = torchaudio.transforms.MFCC(sample_rate=SR, n_mfcc=40, melkwargs=dict(n_fft=2048, hop_length=512, n_mels=128))
    mf2 = T.MFCC()                                     # defaults 1024 / 256 / 128
    x2 = synth(2, 8192, 52)
    mfn = T.MFCC(norm_mode="gaussian")
    mfn.scale_data(mfn.transform(x2))
    r, c, v, shp = sparse(mf.transform.mel_scale.fb)
    save("mfcc", x=x, y=y, y_dct40=tm(x), x2=x2, y2=mf2(x2), y2_gauss=mfn(x2),
         y2_gauss_offset=mfn.norm.offset, y2_gauss_scale=mfn.norm.scale,
         fb_rows=r, fb_cols=c, fb_vals=v, fb_shape=shp, dct=tm.dct_mat)

    # ---- mu-law / one-hot (A18, A19) ---------------------------------------
    g = torch.Generator().manual_seed(61)
    x = torch.cat([2 * torch.rand(2, 16384, generator=g) - 1,
                   torch.tensor([[-1.0, 0.0, 1.0, 0.5, -0.5, 1e-7, -1e-7, 0.999999] + [0.0] * 16376] * 1)], 0)
    ml = T.MuLaw()
    q = ml(x)
    xs = x[:, :64]
    save("mulaw", x=x, q=q, dec=ml.invert(q),
         q64_channel=T.MuLaw(one_hot="channel")(xs), q64_categorical=T.MuLaw(one_hot="categorical")(xs),
         q_c64=T.MuLaw(channels=64)(x), dec_c64=T.MuLaw(channels=64).invert(T.MuLaw(channels=64)(x)),
         onehot64=T.OneHot(n_classes=256)(q[:, :64]))

    # ---- Mono / MidSide (A20) ----------------------------------------------
    x = synth(2, 2048, 71, channels=2)
    ms = T.MidSide()
    save("raw", x=x, mono=T.Mono()(x), midside=ms(x), midside_inv=ms.invert(ms(x)),
         midside_nopad=T.MidSide(pad_mid=False)(x))

    # ---- OverlapAdd + Realtime* streaming (A16, A17) -----------------------
    x = synth(2, 2048 * 3, 81)
    for cls, tag in ((T.RealtimeSTFT, "rtstft"), (T.RealtimeDGT, "rtdgt")):
        oa = T.OverlapAdd(512, 128)
        rt = cls(n_fft=512, hop_length=128)
        arrays = dict(x=x, gain=oa.gain_compensation, window=rt.window[:512], inv_window=rt.inv_window[:512])
        for i, chunk in enumerate(x.split(2048, -1)):
            fr = oa(chunk)
            Xc = rt(fr)
            yi = rt.invert(Xc)
            arrays["frames_%d" % i] = fr.clone()
            arrays["X_%d" % i] = Xc
            arrays["inv_frames_%d" % i] = yi
            arrays["out_%d" % i] = oa.invert(yi)
        save("stream_%s" % tag, **arrays)
    oa = T.OverlapAdd()                                 # defaults 1024 / 128
    xs = synth(2, 2048, 82)
    fr = oa(xs)
    save("oadd_default", x=xs, frames=fr, out=oa.invert(fr.clone()), gain=oa.gain_compensation)

    # ---- Normalize (A7) ----------------------------------------------------
    g = torch.Generator().manual_seed(91)
    x = torch.randn(4, 1000, generator=g) * 3 + 1.5
    arrays = dict(x=x)
    for mode in ("unipolar", "bipolar", "gaussian"):
        nm = T.Normalize(mode)
        nm.scale_data(x)
        arrays[mode + "_offset"], arrays[mode + "_scale"] = nm.offset, nm.scale
        arrays[mode + "_y"] = nm(x)
    save("normalize", **arrays)

    # ---- whole chains: BASELINE.json configs at fixture size ----------------
    # cfg 1: Mono + STFT(1024,256) + Magnitude()
    x = synth(2, 8192, 101, channels=2)
    ch = T.Mono() + T.STFT(n_fft=1024, hop_length=256) + T.Magnitude()
    ch.scale_data(x)
    save("chain_cfg1", x=x, y=ch(x), offset=ch[2].norm.offset, scale=ch[2].norm.scale)
    # cfg 2: (Mono +) DGT(1024,256) + Magnitude(mel, unipolar, log1p)   [README.md:52-54]
    x = synth(3, 8192, 102)
    ch = T.DGT(n_fft=1024, hop_length=256) + T.Magnitude(mel=True, mode="unipolar", contrast="log1p")
    ch.scale_data(x)
    save("chain_cfg2", x=x, y=ch(x), offset=ch[1].norm.offset, scale=ch[1].norm.scale)
    # cfg 4: MidSide + STFT(4096,1024) + PolarIF, forward and inverse
    x = synth(1, 16384, 104, channels=2)
    ch = T.MidSide() + T.STFT(n_fft=4096, hop_length=1024) + T.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})
    ch.scale_data(x)
    y = ch(x)
    save("chain_cfg4", x=x, y=y, x_inv=ch.invert(y.clone()),
         mag_offset=ch[2].magnitude.norm.offset, mag_scale=ch[2].magnitude.norm.scale,
         ph_offset=ch[2].phase.norm.offset, ph_scale=ch[2].phase.norm.scale)

    # ---- PGHI (DGT default inversion mode; dgt.py:156-236): phase of a small magnitude ----------
    # sinusoid + weak noise so that most bins survive the tolerance and the flood fill has real work
    g = torch.Generator().manual_seed(105)
    n = torch.arange(1280, dtype=torch.float64)
    x = (0.5 * torch.sin(2 * math.pi * 3000.0 * n / SR) + 0.25 * torch.sin(2 * math.pi * 9000.0 * n / SR)).float()[None]
    x = x + 0.05 * (2 * torch.rand(x.shape, generator=g) - 1)
    d = T.DGT(n_fft=128, hop_length=32)
    mag = d(x)[0].abs()
    save("pghi_128_32", x=x, mag=mag, phase=d.pghi(mag.clone(), 1e-2), gamma=d.gamma, eps=d.eps, y=d.invert(mag.clone()[None]))

    # ---- PGHI at the DEFAULT sizes (DGT(): n_fft 1024, hop 256), the reference's default inversion of a magnitude ----
    g = torch.Generator().manual_seed(106)
    n = torch.arange(8192, dtype=torch.float64)
    x = (0.5 * torch.sin(2 * math.pi * 440.0 * n / SR) + 0.25 * torch.sin(2 * math.pi * 3520.0 * n / SR)
         + 0.1 * torch.sin(2 * math.pi * 9000.0 * n / SR)).float()[None]
    x = x + 0.05 * (2 * torch.rand(x.shape, generator=g) - 1)
    d = T.DGT()
    mag = d(x)[0].abs()
    save("pghi_1024_256", x=x, mag=mag, phase=d.pghi(mag.clone(), 1e-2), gamma=d.gamma, eps=d.eps, y=d.invert(mag.clone()[None]))

    # ---- real-time PGHI (RealtimeDGT.pghi, dgt.py:338-452): blocks of 3, 1 and 4 frames through the two-frame history ----
    # the noise floor keeps every bin above tolerance * max, so no bin takes the reference's (unpinnable) random phase;
    # `audible` records it.  The reference's time stencil reads one row of `torch.empty` memory BEFORE the first history frame
    # (dgt.py:373-380: Y[:, 0] is never written and feeds the gradient row the first two new frames use), so its output
    # depends on what the allocator returns; for the fixture `torch.empty` hands out zeros while `pghi` runs (the reference
    # source is untouched) and the restatement is pinned with the same row (`row_before="zeros"`).
    g = torch.Generator().manual_seed(107)
    n = torch.arange(1280, dtype=torch.float64)
    x = torch.stack([(0.5 * torch.sin(2 * math.pi * f0 * n / SR) + 0.2 * torch.sin(2 * math.pi * 2.7 * f0 * n / SR)).float()
                     for f0 in (2500.0, 4100.0)])
    x = x + 0.05 * (2 * torch.rand(x.shape, generator=g) - 1)
    mag = T.DGT(n_fft=128, hop_length=32)(x).abs()[:, 4:12]                  # [2, 8, 65]
    rt = T.RealtimeDGT(n_fft=128, hop_length=32, batch_size=2)
    tol = 1e-6
    phases, hist_mag, hist_phase, audible = [], [], [], []
    for a, b in ((0, 3), (3, 4), (4, 8)):
        blk = mag[:, a:b].clone()
        hist_mag.append(rt.hgi_mag_buffer.clone())
        hist_phase.append(rt.hgi_phase_buffer.clone())
        both = torch.cat([rt.hgi_mag_buffer, blk], -2).clamp(rt.eps, None)
        thr = (tol * both.amax((-2, -1), keepdim=True)).clamp(rt.eps, None)
        audible.append(blk.clamp(rt.eps, None) > thr)
        real_empty = torch.empty
        torch.empty = lambda *a, **k: real_empty(*a, **k).zero_()
        try:
            ph = rt.pghi(blk.clone(), tol)
        finally:
            torch.empty = real_empty
        phases.append(ph)
        rt.update_buffers(blk * torch.exp(ph * torch.full(ph.shape, 1j)))
    save("rtpghi_128_32", mag=mag, phase=torch.cat(phases, -2), audible=torch.cat(audible, -2), hist_mag=torch.stack(hist_mag),
         hist_phase=torch.stack(hist_phase), blocks=np.array([[0, 3], [3, 4], [4, 8]]), tol=tol, gamma=rt.gamma, eps=rt.eps)


if __name__ == "__main__":
    main()
