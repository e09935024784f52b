"""cfg 3 kernel alone: MFCC(2048, 512, 128 mels) on 512 x 10 s clips (one launch of the fused mel-spectrogram kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acids_transforms_b200.transforms as Tr

x = 0.5 * (2 * torch.rand((512, 441000), device="cuda") - 1)
m = Tr.MFCC(n_fft=2048, hop_length=512, n_mels=128).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    m(x)
e.record()
torch.cuda.synchronize()
print("cfg3_mel128_512x10s ms", round(s.elapsed_time(e) / 10, 4))
