import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import transforms as T
x = 0.5 * (2 * torch.rand((256, 2, 176400), device="cuda") - 1)
st = T.STFT(n_fft=4096, hop_length=1024).cuda()
f = lambda: torch.ops.acids_b200.midside_stft_fwd(x, st.window, 4096, 1024, 2)
want = st(T.MidSide().cuda()(x))
got = f()
print("max rel err", float((got - want).abs().max() / want.abs().max()))
for _ in range(3): f()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): f()
e.record(); torch.cuda.synchronize()
print("midside_stft_fwd 4096, 256 stereo x 4 s: %.4f ms" % (s.elapsed_time(e) / 10))
