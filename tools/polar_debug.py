"""Debug / timing probe of the fused STFT -> Polar kernel (tools only)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from acids_transforms_b200 import transforms as T, ops
from conftest import if_mask, branch_cut


def host(t):
    return t.detach().cpu().numpy()


def err_report(n_fft, hop, kind):
    torch.manual_seed(n_fft + len(kind))
    B, L = (40, 6 * n_fft + 3 * hop) if n_fft <= 1024 else (6, 5 * n_fft + hop)
    x = torch.randn(B, L, device="cuda")
    rep = T.Polar(magnitude_args={"mode": "bipolar", "n_fft": n_fft}) if kind == "polar" else T.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": n_fft}, phase_args={"mode": "bipolar", "weighted": kind.endswith("w")})
    ch = (T.STFT(n_fft=n_fft, hop_length=hop) + rep).cuda()
    ch.scale_data(x)
    y, ref = host(ch(x)), host(ch.forward_unfused(x))
    X = host(ch[0](x))
    ok = if_mask(X) if kind != "polar" else ~branch_cut(X)
    sc = float(ch[1].phase.norm.scale)
    d = np.abs(y[..., 1, :] - ref[..., 1, :]) * sc * (np.pi if kind != "polar" else 1.0)
    d = np.where(ok, d, 0)
    absX = np.abs(X)
    per = 2e-6 * absX.max() / np.maximum(absX, 1e-30)
    allow = 1e-5 * np.pi + per
    if kind != "polar":
        allow[..., 1:, :] += per[..., :-1, :]
    bad = d > allow
    idx = np.unravel_index(np.argmax(d - allow), d.shape)
    print("%s n_fft=%d: max err %.3e rad, n_bad %d / %d, worst at %s: err %.3e allow %.3e |X| %.3e prev|X| %.3e peak %.3e y %.5f ref %.5f mag err %.2e" % (
        kind, n_fft, d.max(), bad.sum(), bad.size, idx, d[idx], allow[idx], absX[idx], absX[idx[0], max(idx[1] - 1, 0), idx[2]], absX.max(),
        y[idx[0], idx[1], 1, idx[2]], ref[idx[0], idx[1], 1, idx[2]], np.abs(y[..., 0, :] - ref[..., 0, :]).max()))
    if bad.sum():
        bt = np.argwhere(bad)
        print("   bad frames histogram:", np.bincount(bt[:, 1], minlength=d.shape[1]).tolist())


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


def timing():
    L = 176400
    x = 0.5 * (2 * torch.rand(256, 2, L, device="cuda") - 1)
    w = T.STFT(n_fft=4096, hop_length=1024).window.cuda()
    mag = T.Magnitude(n_fft=4096, mode="bipolar").cuda()
    band = ops.as_band(mag.mel_meta, mag.mel_coef)
    one = torch.ones(1, device="cuda")
    zero = torch.zeros(1, device="cuda")
    xm = ops.midside(x)
    print("band coef_len", band.coef_len, "n_out", band.n_out)
    for name, fn in [
        ("complex stft 4096 (512 mono clips)", lambda: ops.stft_fwd(xm, w, 4096, 1024)),
        ("fused mag (band)", lambda: ops.stft_mag_fwd(xm, w, 4096, 1024, band, "log1p", 1e-6, zero, one)),
        ("fused mag (no band)", lambda: ops.stft_mag_fwd(xm, w, 4096, 1024, None, "log1p", 1e-6, zero, one)),
        ("fused polar raw (band)", lambda: ops.stft_polar_fwd(xm, w, 4096, 1024, band, "log1p", 1e-6, zero, one, 0, 0, False, zero, one)),
        ("fused polar IF (band)", lambda: ops.stft_polar_fwd(xm, w, 4096, 1024, band, "log1p", 1e-6, zero, one, 2, 0, False, zero, one)),
        ("fused polar IF (no band)", lambda: ops.stft_polar_fwd(xm, w, 4096, 1024, None, "log1p", 1e-6, zero, one, 2, 0, False, zero, one)),
        ("fused polar IF midside (band)", lambda: ops.stft_polar_fwd(x, w, 4096, 1024, band, "log1p", 1e-6, zero, one, 2, 0, False, zero, one, midside=2)),
    ]:
        print("%-40s %.3f ms" % (name, timeit(fn)))
    x1 = 0.5 * (2 * torch.rand(512, L, device="cuda") - 1)
    w1 = T.STFT().window.cuda()
    m1 = T.Magnitude().cuda()
    b1 = ops.as_band(m1.mel_meta, m1.mel_coef)
    for name, fn in [
        ("1024: fused mag (band) 512 clips", lambda: ops.stft_mag_fwd(x1, w1, 1024, 256, b1, "log1p", 1e-6, zero, one)),
        ("1024: fused polar IF (band)", lambda: ops.stft_polar_fwd(x1, w1, 1024, 256, b1, "log1p", 1e-6, zero, one, 2, 0, False, zero, one)),
        ("1024: stft_stats", lambda: ops.stft_stats(x1, w1, 1024, 256, "log1p", 1e-6)),
        ("1024: complex", lambda: ops.stft_fwd(x1, w1, 1024, 256)),
    ]:
        print("%-40s %.3f ms" % (name, timeit(fn)))


if __name__ == "__main__":
    if "--time" in sys.argv:
        timing()
    else:
        for n_fft, hop in [(256, 64), (512, 128), (1024, 256), (2048, 512), (4096, 1024), (8192, 2048)]:
            for kind in ("polar", "polarif", "polarifw"):
                err_report(n_fft, hop, kind)
        timing()
