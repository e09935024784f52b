"""acids_polar_rows_fwd (one read of the spectrum) against acids_mag_epilogue + acids_phase_fwd, per row length."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import ops
from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW
from acids_transforms_b200.transforms.spectral_repr import build_mel_banks


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = {}
    one = torch.ones(1, device="cuda")
    for n_fft, B, T in ((512, 2048, 690), (1024, 1024, 690), (2048, 512, 431), (4096, 512, 173), (8192, 256, 87)):
        F = n_fft // 2 + 1
        X = torch.view_as_complex(torch.randn(B, T, F, 2, device="cuda"))
        fwd = build_mel_banks(44100, n_fft, True)[0]
        band = ops.BandedMatrix(fwd[0] if fwd.dim() == 3 else fwd)
        out = torch.empty((B, T, 2, F), device="cuda")
        for tag, mode in (("if", PHASE_IF), ("raw", PHASE_RAW)):
            def two():
                ops.mag_epilogue(X, band, "log1p", 1e-7, one, one, False, out=out, out_slot=0, out_slots=2)
                ops.phase_fwd(X, mode, "forward", False, one, one, False, out=out, out_slot=1, out_slots=2)
            res["%d_%s_two_kernels" % (n_fft, tag)] = round(timeit(two), 4)
            res["%d_%s_rows" % (n_fft, tag)] = round(timeit(lambda: ops.polar_rows_fwd(X, band, "log1p", 1e-7, one, one, mode, "forward", False, one, one)), 4)
        del X, out
    print(json.dumps(res))


if __name__ == "__main__":
    main()
