"""Times phase_inv + polar_to_complex against the one-pass phase_inv_polar at the cfg-4 shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import ops
from acids_transforms_b200._lib import PHASE_IF, PHASE_RAW


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
    B, T, F = 512, 173, 2049
    stacked = torch.randn(B, T, 2, F, device="cuda")
    y = stacked[..., 1, :]
    mag = torch.rand(B, T, F, device="cuda")
    res = {}
    for tag, mode in (("if", PHASE_IF), ("raw", PHASE_RAW)):
        res[tag + "_phase_inv"] = timeit(lambda: ops.phase_inv(y, mode, "forward"))
        ph = ops.phase_inv(y, mode, "forward")
        res[tag + "_polar"] = timeit(lambda: ops.polar_to_complex(mag, ph))
        res[tag + "_fused"] = timeit(lambda: ops.phase_inv_polar(y, mag, mode, "forward"))
    del stacked, mag, ph
    X = torch.view_as_complex(torch.randn(B, T, F, 2, device="cuda"))
    one = torch.ones(1, device="cuda")
    res["fwd_mag_epilogue"] = timeit(lambda: ops.mag_epilogue(X, None, "log1p", 1e-7, one, one))
    res["fwd_phase_if"] = timeit(lambda: ops.phase_fwd(X, PHASE_IF, "forward", False, one, one))
    res["fwd_polar_one_pass"] = timeit(lambda: ops.polar_fwd(X, "log1p", 1e-7, one, one, PHASE_IF, "forward", False, one, one))
    del X
    import acids_transforms_b200.transforms as Tr
    x4 = 0.5 * (2 * torch.rand((256, 2, 176400), device="cuda") - 1)
    ms_ = Tr.MidSide().cuda()
    st = Tr.STFT(n_fft=4096, hop_length=1024).cuda()
    res["fwd_midside"] = timeit(lambda: ms_(x4))
    xm = ms_(x4)
    res["fwd_stft4096"] = timeit(lambda: st(xm))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
