import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import ops
n_fft, hop, block = 1024, 256, 1024
F = n_fft // 2 + 1
def run(x, w, label):
    tail = torch.zeros((1, n_fft - hop), device="cuda")
    X = ops.stream_analysis(x, w, n_fft, hop, tail)
    sb = torch.cat([torch.zeros((1, n_fft - hop), device="cuda"), x], -1)
    fr = sb.unfold(-1, n_fft, hop).contiguous()
    Xb = ops.stft_fwd(fr, w, n_fft, n_fft, center=False)
    d = (torch.view_as_real(X) != torch.view_as_real(Xb))
    print(label, "mismatches per frame:", d.sum((-1, -2)).tolist(), "max", float((X - Xb).abs().max()))
    return X, Xb
ones = torch.ones(n_fft, device="cuda")
hann = torch.hann_window(n_fft, device="cuda")
x = torch.zeros((1, block), device="cuda"); x[0, 700] = 1.0
run(x, ones, "impulse, rect window:")
x = torch.zeros((1, block), device="cuda"); x[0, 100] = 1.0; x[0, 613] = 2.0; x[0, 1001] = -3.0
run(x, ones, "3 impulses, rect window:")
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randint(-8, 9, (1, block), generator=g, device="cuda")).float()
run(x, ones, "small integers, rect window:")
run(x, hann, "small integers, hann:")
x = 2 * torch.rand((1, block), generator=g, device="cuda") - 1
X, Xb = run(x, ones, "uniform noise, rect:")
X, Xb = run(x, hann, "uniform noise, hann:")
# where (which bins) do they differ in frame 3
d = (torch.view_as_real(X[0, 3]) != torch.view_as_real(Xb[0, 3])).any(-1).nonzero().flatten().tolist()
print("frame 3 differing bins:", d[:64], "count", len(d))
