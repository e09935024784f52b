"""GPU PGHI (csrc/pghi.cu) against the host restatement (numpy + heapq): time per batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acids_transforms_b200.transforms as Tr
from acids_transforms_b200.transforms import pghi as P

d = Tr.DGT(n_fft=1024, hop_length=256).cuda()
g = torch.Generator(device="cuda").manual_seed(3)
x = 0.5 * (2 * torch.rand((256, 176400), generator=g, device="cuda") - 1)
n = torch.arange(176400, device="cuda")
x += 0.5 * torch.sin(2 * torch.pi * 440.0 * n / 44100)
mag = d(x).abs()
for B in (1, 16, 256):
    m = mag[:B].contiguous()
    d.pghi(m, 1e-2); torch.cuda.synchronize()
    t0 = time.perf_counter(); ph = d.pghi(m, 1e-2); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("GPU pghi, %3d clips x 690 x 513: %.3f s (%.1f ms per clip)" % (B, t1 - t0, 1e3 * (t1 - t0) / B))
t0 = time.perf_counter(); want = P.pghi(mag[0].cpu(), float(d.gamma), 1024, 256, 1e-2, float(d.eps)); t1 = time.perf_counter()
print("host numpy + heapq, 1 clip: %.3f s" % (t1 - t0))
print("max |gpu - host| on clip 0: %.2e rad; visited %.1f %% of the bins" % (float((ph[0].cpu() - want).abs().max()), 100 * float((want != 0).float().mean())))

# ---- frame-by-frame variant: RealtimeDGT.invert(magnitude) in its default mode, one frame per call ----
if "--rt" in sys.argv or True:
    for B in (1, 16, 256):
        rt_dev = Tr.RealtimeDGT(n_fft=1024, hop_length=256, batch_size=B).cuda()
        rt_host = Tr.RealtimeDGT(n_fft=1024, hop_length=256, batch_size=B)
        blocks = [mag[:B, 100 + i:101 + i].contiguous() for i in range(12)]
        for blk in blocks[:2]:
            rt_dev.invert(blk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for blk in blocks[2:]:
            y = rt_dev.invert(blk)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        line = "RealtimeDGT.invert (pghi), %3d streams, one frame per call: device %.3f ms per call" % (B, 1e3 * (t1 - t0) / 10)
        if B <= 16:
            hb = [b.cpu() for b in blocks]
            rt_host.invert(hb[0]); rt_host.invert(hb[1])
            t0 = time.perf_counter()
            for blk in hb[2:6]:
                rt_host.invert(blk)
            t1 = time.perf_counter()
            line += "; host restatement %.1f ms per call" % (1e3 * (t1 - t0) / 4)
        print(line)

# ---- the frame-by-frame kernel alone (CUDA events around ops.rt_pghi, one frame of 513 bins per stream) ----
from acids_transforms_b200 import ops
rt = Tr.RealtimeDGT(n_fft=1024, hop_length=256)
gm, eps = float(rt.gamma), float(rt.eps)
for B in (1, 256):
    hm, mg, hp = mag[:B, 100:102].contiguous(), mag[:B, 102:103].contiguous(), torch.zeros((B, 513), device="cuda")
    for tol in (1e-2, 1e-6):
        for _ in range(3):
            ops.rt_pghi(mg, hm, hp, gm, 1024, 256, tol, eps)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.rt_pghi(mg, hm, hp, gm, 1024, 256, tol, eps)
        e.record(); torch.cuda.synchronize()
        print("rt_pghi kernel, %3d streams, tolerance %g: %.3f ms per frame" % (B, tol, s.elapsed_time(e) / 20))
