"""Dense tcgen05 mel projection (csrc/mel_tc.cu) against the banded FP32 forms at the cfg-3 shape (n_fft 2048, 128 mels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import ops
from acids_transforms_b200.transforms.spectral_repr import melscale_fbanks


def timeit(fn, iters=10, warm=2):
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


B, L, n, h, n_mels = 128, 441000, 2048, 512, 128
T, F = 1 + L // h, n // 2 + 1
x = 0.5 * (2 * torch.rand((B, L), device="cuda") - 1)
w = torch.hann_window(n).cuda()
fb = melscale_fbanks(F, 0.0, 22050.0, n_mels, 44100).cuda()
band = ops.BandedMatrix(fb.cpu())
X = ops.stft_fwd(x, w, n, h, True)
P = (X.abs() ** 2).contiguous()
flops = 2.0 * B * T * F * n_mels
ms = timeit(lambda: ops.mel_tc(P, fb))
print("mel_tc (tcgen05 3xTF32), %d x %d x %d -> %d: %.3f ms = %.1f dense TFLOP/s fp32-equivalent (x3 issued), reads %.0f GB/s"
      % (B, T, F, n_mels, ms, flops / ms / 1e9, B * T * F * 4 / ms / 1e6))
ms2 = timeit(lambda: torch.matmul(P, fb))
print("torch.matmul fp32 (cuBLAS) on the same spectrum: %.3f ms" % ms2)
ms3 = timeit(lambda: ops.mag_epilogue(X, band, None, 1e-7, None, None, False))
print("banded FP32 epilogue kernel on the materialised complex spectrum (|X|, 1025 -> 128): %.3f ms" % ms3)
ms4 = timeit(lambda: ops.melspec_fwd(x, w, n, h, band, 2.0, None, None))
ms5 = timeit(lambda: ops.stft_fwd(x, w, n, h, True))
print("fused wave -> mel (banded, spectrum never in HBM): %.3f ms; complex STFT alone: %.3f ms" % (ms4, ms5))
got = ops.mel_tc(P, fb)
want = torch.einsum("btf,fm->bmt", P[:4].double(), fb.double())
print("max rel err vs float64: %.2e" % float(((got[:4].double() - want).abs().max() / want.abs().max())))
