"""CPU tests: batch sharding (incl. a world_size-2 gloo run), the torch-CPU port used as CPU baseline, and
bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_parity, dense, load_golden
from acids_transforms_b200.sharding import merge_stats, shard_range


def test_shard_range_partitions():
    for total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 8, 3) == (24576, 32768)          # cfg 5: 8,192 clips per GPU
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_merge_stats_is_exact():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(10000) * 3 + 1
    parts = np.split(x, [1234, 5000, 5001])
    st = [[p.min(), p.max(), p.mean(), p.std(ddof=1) if p.size > 1 else 0.0] for p in parts]
    m = merge_stats(st, [p.size for p in parts])
    assert np.allclose(m, [x.min(), x.max(), x.mean(), x.std(ddof=1)], rtol=1e-12)


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from acids_transforms_b200.sharding import shard, shard_range, all_reduce_stats
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
g = torch.Generator().manual_seed(0)
x = torch.randn(11, 1000, generator=g, dtype=torch.float64)      # the whole job, identical on both ranks
mine = shard(x, world, rank)
lo, hi = shard_range(11, world, rank)
assert mine.shape[0] == hi - lo
st = torch.tensor([mine.min(), mine.max(), mine.mean(), mine.std()], dtype=torch.float64)
merged = all_reduce_stats(st, mine.numel())
want = torch.tensor([x.min(), x.max(), x.mean(), x.std()], dtype=torch.float64)
assert torch.allclose(merged, want, rtol=1e-12), (merged, want)
# no collective is needed on the data path: gather only to prove the shards tile the batch
parts = [None] * world
dist.all_gather_object(parts, (lo, hi))
assert parts == [shard_range(11, world, r) for r in range(world)]
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_torch_port_matches_reference():
    from oracle import torch_port as P
    g = load_golden("chain_cfg2")
    b = load_golden("mel_bank_1024")
    bank = torch.from_numpy(dense(b["rows"], b["cols"], b["vals"], b["shape"]))[None]
    y = P.cfg2_forward(torch.from_numpy(g["x"]), P.gaussian_window(1024), bank, torch.from_numpy(g["offset"]), torch.from_numpy(g["scale"]))
    assert_parity(y.numpy(), g["y"], 1e-6, "torch port cfg2")
    s = load_golden("stft_1024_256")
    assert_parity(P.istft(torch.from_numpy(s["X"]), P.hann(1024), 1024, 256).numpy(), s["y"], 1e-6, "torch port istft")


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
