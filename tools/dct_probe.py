import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acids_transforms_b200 import ops
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / iters
B, M, T, C = 512, 128, 862, 40
mel = torch.rand((B, M, T), device="cuda") ** 4 * 50 + 1e-9
k = torch.arange(C, dtype=torch.float64)[None, :]; n = torch.arange(M, dtype=torch.float64)[:, None]
dct = (torch.cos(math.pi / M * (n + 0.5) * k) * math.sqrt(2.0 / M)); dct[:, 0] *= 1 / math.sqrt(2.0); dct = dct.float().cuda()
for tc in (False, True):  # FP32 register-tiled kernel vs tcgen05 3xTF32
    ms = timeit(lambda: ops.mfcc_dct(mel, dct, 80.0, tensor_cores=tc))
    print("mfcc tail (group max + dB + DCT) tensor_cores=%s: %.3f ms  (%.1f GFLOP/s useful, %.0f GB/s)" % (tc, ms, 2.0 * B * T * M * C / ms / 1e6, (2 * B * M * T * 4 + B * C * T * 4) / ms / 1e6))
