"""A few warm launches of ONE kernel of interest, for `ncu -k regex:... -s <n> -c 1` (tools only).

    python tools/ncu_targets.py fwd|inv|cfg3|cfg3dct|polar4096|stats
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acids_transforms_b200.transforms as Tr
from acids_transforms_b200 import ops


def main():
    what = sys.argv[1]
    g = torch.Generator(device="cuda").manual_seed(1234)
    L = 176400
    if what in ("fwd", "inv", "stats"):
        x = 0.5 * (2 * torch.rand((1024, L), generator=g, device="cuda") - 1)
        ch = (Tr.DGT(n_fft=1024, hop_length=256, inversion_mode="random") + Tr.Magnitude(mel=True, mode="unipolar", contrast="log1p")).cuda()
        ch.scale_data(x[:16])
        if what == "fwd":
            for _ in range(3):
                y = ch(x)
        elif what == "stats":
            for _ in range(3):
                ch.scale_data(x)
        else:
            X = ch[0](x)
            del x
            for _ in range(3):
                y = ch[0].invert(X)
    elif what in ("cfg3", "cfg3dct"):
        x = 0.5 * (2 * torch.rand((512, 441000), generator=g, device="cuda") - 1)
        m = Tr.MFCC(n_fft=2048, hop_length=512, n_mels=128, **({"n_mfcc": 40} if what == "cfg3dct" else {})).cuda()
        for _ in range(3):
            y = m(x)
    elif what == "polar4096":
        x = 0.5 * (2 * torch.rand((256, 2, L), generator=g, device="cuda") - 1)
        ch = (Tr.MidSide() + Tr.STFT(n_fft=4096, hop_length=1024) + Tr.PolarIF(
            magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})).cuda()
        ch.scale_data(x[:8])
        for _ in range(3):
            y = ch(x)
    torch.cuda.synchronize()
    print(what, tuple(y.shape) if what != "stats" else "ok")


if __name__ == "__main__":
    main()
