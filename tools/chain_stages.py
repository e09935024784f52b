"""Per-kernel device time of the cfg-4 forward / inverse chains (torch.profiler, CUDA activities only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import acids_transforms_b200.transforms as Tr


def main():
    x4 = 0.5 * (2 * torch.rand((256, 2, 176400), device="cuda") - 1)
    ch = (Tr.MidSide() + Tr.STFT(n_fft=4096, hop_length=1024) + Tr.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})).cuda()
    ch.scale_data(x4[:8])
    for name, fn in (("forward", lambda: ch(x4)), ("inverse", None)):
        if fn is None:
            y = ch(x4)
            fn = lambda: ch.invert(y)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
        print("==", name)
        rows = [(e.key, e.device_time_total / 5.0, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
        for k, t, c in sorted(rows, key=lambda r: -r[1])[:12]:
            print("%9.1f us  x%.0f  %s" % (t, c, k[:110]))


if __name__ == "__main__":
    main()
