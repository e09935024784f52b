import sys, time, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from acids_transforms_b200 import transforms as Tr
from acids_transforms_b200.hostpipe import HostPipeline
dev = torch.device("cuda", 0)
x = bench.synth_clips(512, dev, 1234)
chain = (Tr.DGT(sr=44100, n_fft=1024, hop_length=256, inversion_mode="random") + Tr.Magnitude(sr=44100, mel=True, mode="unipolar", contrast="log1p", n_fft=1024)).to(dev)
chain.scale_data(x[:16])
x_host = torch.empty((512, bench.L), dtype=torch.float32, pin_memory=True); x_host.copy_(x)
for chunk in (64,):
    pipe = HostPipeline(chain, chunk_clips=chunk, device=dev)
    out_host = pipe(x_host); torch.cuda.synchronize()
    ts = []
    for i in range(20):
        t = time.perf_counter(); pipe(x_host, out_host); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print("chunk", chunk, " ".join("%.1f" % t for t in ts))
    print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_reserved() / 1e9)
