"""Quick device-side timing of the headline kernels (not the bench): CUDA events, L2-exceeding inputs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from acids_transforms_b200 import ops
from acids_transforms_b200.transforms.dgt import gaussian_window
from acids_transforms_b200.transforms.spectral_repr import build_mel_banks

PEAK = 6536.4  # GB/s measured copy bandwidth (MEASURED_PEAKS.json)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    nums = [a for a in sys.argv[1:] if a.isdigit()]
    B = int(nums[0]) if nums else 1024
    L, n, h = 176400, 1024, 256
    T, F = 1 + L // h, n // 2 + 1
    x = 0.5 * (2 * torch.rand((B, L), device="cuda") - 1)
    w = gaussian_window(n).float().cuda()
    hw = torch.hann_window(n).cuda()
    fwd = build_mel_banks(44100, n, True)[0]
    fwd = fwd[0] if fwd.dim() == 3 else fwd
    band = ops.BandedMatrix(fwd)
    eps = float(np.finfo(np.float32).eps)
    out = torch.empty((B, T, F), device="cuda")
    res = {}
    ms = timeit(lambda: ops.stft_mag_fwd(x, w, n, h, band, "log1p", eps, 0.1, 2.0, out=out))
    byt = B * (4 * L + 4 * T * F)
    res["fused_fwd_mel_log1p"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK, audio_s_per_s=B * 4 / (ms / 1e3))
    if "--mel-only" in sys.argv:
        print("fused_fwd_mel_log1p", json.dumps({a: round(b, 4) for a, b in res["fused_fwd_mel_log1p"].items()}))
        return
    ms = timeit(lambda: ops.stft_mag_fwd(x, w, n, h, None, "log1p", eps, 0.1, 2.0, out=out))
    res["fused_fwd_nomel_log1p"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    ms = timeit(lambda: ops.stft_mag_fwd(x, w, n, h, None, None, eps, None, None, out=out))
    res["fused_fwd_mag_only"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    if "--fwd-only" in sys.argv:
        for k, v in res.items():
            print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))
        return
    X = ops.stft_fwd(x, hw, n, h)
    ms = timeit(lambda: ops.stft_fwd(x, hw, n, h))
    byt = B * (4 * L + 8 * T * F)
    res["stft_complex"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    ms = timeit(lambda: ops.istft_ola(X, hw, n, h, check_envelope=False))
    byt = B * (8 * T * F + 4 * h * (T - 1))
    res["istft_ola"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK, audio_s_per_s=B * 4 / (ms / 1e3))
    ms = timeit(lambda: ops.mag_epilogue(X, band, "log1p", eps, 0.1, 2.0, out=out))
    byt = B * (8 * T * F + 4 * T * F)
    res["mag_epilogue"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    ms = timeit(lambda: ops.phase_fwd(X, 2, "forward", False, 0.0, 1.0, out=out))
    res["phase_if"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    ms = timeit(lambda: ops.mulaw_encode(x))
    byt = B * L * 12
    res["mulaw"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    # eager torch chain on the same GPU (what the reference becomes after .to('cuda')), for context
    mb = fwd.cuda()[None]
    def eager():
        Xe = torch.stft(x, n, h, window=w, return_complex=True).transpose(-2, -1)
        return (torch.log(1 + torch.matmul(Xe.abs(), mb)) - 0.1) / 2.0
    if B <= 1024:
        ms = timeit(eager, iters=5)
        res["torch_eager_cuda_chain"] = dict(ms=ms)
    if "--all" in sys.argv:
        del X, out
        torch.cuda.empty_cache()
        # cfg 3: MFCC = MelSpectrogram(n_fft=2048, hop=512, 128 mels) on 10 s clips (batch reduced to fit the probe)
        B3, L3, n3, h3 = 512, 441000, 2048, 512
        T3 = 1 + L3 // h3
        x3 = 0.5 * (2 * torch.rand((B3, L3), device="cuda") - 1)
        import torchaudio
        fb = torchaudio.functional.melscale_fbanks(n3 // 2 + 1, 0.0, 22050.0, 128, 44100)
        mel = ops.BandedMatrix(fb)
        hw3 = torch.hann_window(n3).cuda()
        ms = timeit(lambda: ops.melspec_fwd(x3, hw3, n3, h3, mel, 2.0), iters=10)
        byt = B3 * (4 * L3 + 4 * 128 * T3)
        res["cfg3_melspec_2048_128"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK, audio_s_per_s=B3 * 10 / (ms / 1e3))
        X3 = ops.stft_fwd(x3[:256], hw3, n3, h3)
        ms = timeit(lambda: ops.istft_ola(X3, hw3, n3, h3, check_envelope=False), iters=10)
        byt = 256 * (8 * T3 * (n3 // 2 + 1) + 4 * h3 * (T3 - 1))
        res["istft_2048_512"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        del x3, X3
        torch.cuda.empty_cache()
        # cfg 4: stereo clips, STFT(4096, 1024) -> complex; ISTFT back; Phase/IF on the spectrum
        B4, n4, h4 = 1024, 4096, 1024
        x4 = 0.5 * (2 * torch.rand((B4, L), device="cuda") - 1)
        hw4 = torch.hann_window(n4).cuda()
        T4, F4 = 1 + L // h4, n4 // 2 + 1
        X4 = ops.stft_fwd(x4, hw4, n4, h4)
        ms = timeit(lambda: ops.stft_fwd(x4, hw4, n4, h4), iters=10)
        byt = B4 * (4 * L + 8 * T4 * F4)
        res["cfg4_stft_4096"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        ms = timeit(lambda: ops.istft_ola(X4, hw4, n4, h4, check_envelope=False), iters=10)
        byt = B4 * (8 * T4 * F4 + 4 * h4 * (T4 - 1))
        res["cfg4_istft_4096"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        o4 = torch.empty((B4, T4, F4), device="cuda")
        ms = timeit(lambda: ops.phase_fwd(X4, 2, "forward", False, 0.0, 1.0, out=o4), iters=10)
        byt = B4 * 12 * T4 * F4
        res["cfg4_if_4096"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
    for k, v in res.items():
        print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))


if __name__ == "__main__":
    main()
