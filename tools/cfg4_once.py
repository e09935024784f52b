"""Two passes of the cfg-4 chain, forward and inverse (for `ncu -k regex:acids`; the last launch of each kernel is the warm one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acids_transforms_b200.transforms as Tr


def main():
    x4 = 0.5 * (2 * torch.rand((256, 2, 176400), device="cuda") - 1)
    ch = (Tr.MidSide() + Tr.STFT(n_fft=4096, hop_length=1024) + Tr.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})).cuda()
    ch.scale_data(x4[:8])
    for _ in range(2):
        y = ch(x4)
        z = ch.invert(y)
    torch.cuda.synchronize()
    print(tuple(y.shape), tuple(z.shape))


if __name__ == "__main__":
    main()
