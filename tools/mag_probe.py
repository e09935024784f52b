"""Stand-alone launches of the magnitude kernels at the cfg-4 shape (for ncu -k regex:mag_)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acids_transforms_b200.transforms.spectral_repr import Magnitude


def main():
    B, T, F = 512, 173, 2049
    X = torch.view_as_complex(torch.randn(B, T, F, 2, device="cuda"))
    m = Magnitude(n_fft=4096, mode="bipolar").cuda()
    m.scale_data(X[:4])
    for _ in range(3):
        y = m(X)
    torch.cuda.synchronize()
    z = m.invert(y)
    torch.cuda.synchronize()
    print(y.shape, z.shape)


if __name__ == "__main__":
    main()
