"""cfg 3 at BASELINE size per clip (10 s @ 44.1 kHz, n_fft 2048, hop 512, 128 mels [+ 40 coefficients]) and the cross-rank
`scale_data` fit on NCCL.  Full-size checks use what the domain offers at sizes the CPU oracle cannot finish in seconds:
torchaudio's own MelSpectrogram / MFCC on the same GPU (the reference's MFCC is that module, mel.py:38-44, :68-73),
batch independence and run-to-run determinism."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_parity

pytestmark = pytest.mark.gpu


def host(t):
    return t.detach().cpu().numpy()


def test_cfg3_full_size_chain():
    import torchaudio
    from acids_transforms_b200 import transforms as T
    torch.manual_seed(11)
    L = 441000                                          # 10 s clips: 862 frames of 2048 samples
    x = 0.5 * (2 * torch.rand(24, L, device="cuda") - 1)
    x[3] *= torch.linspace(0, 1, L, device="cuda")      # not only stationary noise
    x[5, ::7] = 0.9
    m = T.MFCC(n_fft=2048, hop_length=512, n_mels=128).cuda()
    y = m(x)
    assert tuple(y.shape) == (24, 128, 862)
    ta = torchaudio.transforms.MelSpectrogram(sample_rate=44100, n_fft=2048, hop_length=512, n_mels=128).cuda()
    assert_parity(host(y), host(ta(x)), 1e-4, "cfg3 mel spectrogram vs torchaudio (cuFFT + SGEMM) on the same GPU")
    # the persistent grid hands every CTA a different run of frames: a clip must not depend on its neighbours
    for b in (0, 3, 23):
        assert torch.equal(y[b], m(x[b:b + 1])[0])
    assert torch.equal(y, m(x))
    m40 = T.MFCC(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40).cuda()
    y40 = m40(x)
    assert tuple(y40.shape) == (24, 40, 862)
    ta40 = torchaudio.transforms.MFCC(sample_rate=44100, n_mfcc=40, melkwargs=dict(n_fft=2048, hop_length=512, n_mels=128)).cuda()
    assert_parity(host(y40), host(ta40(x)), 1e-4, "cfg3 MFCC-40 (dB + top_db + tensor-core DCT) vs torchaudio")


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from acids_transforms_b200 import transforms as T
from acids_transforms_b200.sharding import shard, all_reduce_stats
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
g = torch.Generator().manual_seed(0)
x = (0.5 * (2 * torch.rand(4 * world + 1, 44100, generator=g) - 1) * torch.linspace(0.2, 1.0, 4 * world + 1)[:, None]).cuda()   # the whole job
whole = (T.DGT(n_fft=1024, hop_length=256, inversion_mode="random") + T.Magnitude(mel=True, mode="gaussian", contrast="log1p")).cuda()
whole.scale_data(x)                                    # single-process fit: the oracle of this test
mine = shard(x, world, rank)
ch = (T.DGT(n_fft=1024, hop_length=256, inversion_mode="random") + T.Magnitude(mel=True, mode="gaussian", contrast="log1p")).cuda()
# one pass of the forward kernel's statistics mode on this rank's clips, then an exact merge of (min, max, mean, std, n) over NCCL
st = torch.ops.acids_b200.stft_stats(mine, ch[0].window, 1024, 256, 1, ch[1]._eps)     # contrast 1 = log1p
n = mine.shape[0] * (1 + mine.shape[1] // 256) * 513
merged = all_reduce_stats(st.to(torch.float64), n)
ch[1].norm.set_stats(merged.to(torch.float32))
for name in ("offset", "scale"):
    a, b = float(getattr(ch[1].norm, name)), float(getattr(whole[1].norm, name))
    assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (name, a, b)
y = ch(mine)
assert torch.allclose(y, shard(whole(x), world, rank), rtol=1e-4, atol=1e-5)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_scale_data_across_ranks_on_nccl(tmp_path):
    """per-rank statistics from the fused forward kernel, merged exactly over NCCL; every rank ends
    with the offset / scale a single process fits on the whole job.  Needs >= 2 GPUs (`gpurun --gpus 2`)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29623", WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    assert all("ok" in o for o in outs)
