"""Latency of one streaming round trip (OverlapAdd + RealtimeSTFT, forward and inverse): eager modules vs GraphedStep."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import acids_transforms_b200.transforms as Tr
from acids_transforms_b200.streaming import GraphedStep, StreamStep


def wall_us(fn, x, iters=300, warm=30):
    for _ in range(warm):
        fn(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn(x)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def main():
    res = {}
    for B, block in ((1, 1024), (16, 1024), (64, 4096)):
        eager = (Tr.OverlapAdd(1024, 256) + Tr.RealtimeSTFT(n_fft=1024, hop_length=256)).cuda()
        graphed = copy.deepcopy(eager)
        x = 2 * torch.rand((B, block), device="cuda") - 1
        step = GraphedStep(graphed, lambda b: graphed.invert(graphed(b)), x)
        tag = "B%d_block%d" % (B, block)
        res[tag] = dict(eager_us=round(wall_us(lambda b: eager.invert(eager(b)), x), 1), graph_us=round(wall_us(step, x), 1))
        # the same round trip as ONE kernel (csrc/stream.cu): called through ctypes, and replayed from a CUDA graph
        oadd, rt = Tr.OverlapAdd(1024, 256).cuda(), Tr.RealtimeSTFT(n_fft=1024, hop_length=256).cuda()
        one = StreamStep(oadd, rt, batch_shape=(B,))
        oneg = StreamStep(oadd, rt, batch_shape=(B,), graph=True, block=block)
        res[tag]["one_kernel_us"] = round(wall_us(one.roundtrip, x), 1)
        res[tag]["one_kernel_graph_us"] = round(wall_us(oneg.roundtrip, x), 1)
        # device time of the kernel alone
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(200):
            oneg._graph.replay()
        e.record()
        torch.cuda.synchronize()
        res[tag]["one_kernel_device_us"] = round(s.elapsed_time(e) / 200 * 1e3, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
