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
    print(json.dumps(res))


if __name__ == "__main__":
    main()
