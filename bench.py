#!/usr/bin/env python
"""bench.py — headline benchmark of the acids_transforms spectral hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): DGT(n_fft=1024, hop=256) + Magnitude(mel=True, unipolar, log1p) forward on
1024 synthetic 4 s / 44.1 kHz clips PER GPU (batch sharded across ranks, no collective: weak scaling), plus the
inverse ISTFT + overlap-add of the same batch reported under "inverse".  A step = one pass of the path over the
batch.  One JSON line is printed by rank 0.

  value      audio-seconds per second, inputs resident in HBM, timed with CUDA events on the launching stream,
             max over ranks.  Inputs (722 MB) and outputs (1.45 GB) exceed the 126 MB L2 every step.
  roofline   the dominant kernel (stft_fwd_kernel<1024, fused epilogue>): algorithmic bytes per launch
             (4 L + 4 T F per clip) / measured launch duration, against MEASURED_PEAKS.json.
  e2e        same metric through the public host API (HostPipeline over the module chain) with pinned HOST
             buffers: H2D of the clips and D2H of the features inside the timed region.
  cpu_baseline  the oracle's torch-CPU port of the reference chain on this box's host cores (rank 0, N=1).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, CLIP_S, N_FFT, HOP = 44100, 4, 1024, 256
L = SR * CLIP_S                      # 176,400 samples
T = 1 + L // HOP                     # 690 frames
F = N_FFT // 2 + 1                   # 513 bins
CLIPS_PER_GPU = 1024
FWD_BYTES_PER_CLIP = 4 * L + 4 * T * F                 # 2,121,480 B  
INV_BYTES_PER_CLIP = 8 * T * F + 4 * HOP * (T - 1)     # 3,537,296 B
METRIC = "audio-sec/s (STFT+mel+log fwd; ISTFT inv) at 1/2/4/8 B200; % HBM roofline"
WORKLOAD = "cfg2: DGT(n_fft=1024,hop=256)+Magnitude(mel,unipolar,log1p) fwd, %d clips x 4 s @ 44.1 kHz per GPU" % CLIPS_PER_GPU


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    In-process NVML (nvidia_ml_py) from a background thread, this rank's GPU only, every 100 ms.  A looping
    `nvidia-smi` child re-enumerates every GPU of the box on each poll and holds driver locks long enough to stall
    the many small CUDA API calls of the end-to-end pipeline (measured: 15 ms -> 20-55 ms per e2e step); it is only
    the fallback when NVML cannot be loaded."""
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.proc = None
        self.thread = None
        self.rows = []          # (sm_mhz, sm_max_mhz, power_w, [reasons])
        self.how = None

    def _nvml_loop(self, nv, handle, stop):
        bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        try:
            sm_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        except Exception:
            sm_max = None
        while not stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.rows.append((sm, sm_max, pw, [n for b, n in bits if mask & b]))
            except Exception:
                pass
            stop.wait(0.1)

    def __enter__(self):
        import threading
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES so that rank i samples ITS device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if int(self.idx) < len(ids):
                    phys = ids[int(self.idx)]
            handle = nv.nvmlDeviceGetHandleByUUID(phys) if not phys.isdigit() else nv.nvmlDeviceGetHandleByIndex(int(phys))
            self._stop = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle, self._stop), daemon=True)
            self.thread.start()
            self.how = "nvml"
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.idx, "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            self.thread = None
            return
        if self.proc is None:
            return
        time.sleep(0.3)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        self.proc = None
        for line in out.splitlines():
            p = [s.strip() for s in line.split(",")]
            if len(p) >= 8 and p[0] == self.idx and p[1].replace(".", "").isdigit():
                reasons = [n for i, n in enumerate(self.NAMES) if p[4 + i].lower().startswith("active")]
                self.rows.append((float(p[1]), float(p[2]) if p[2].replace(".", "").isdigit() else None,
                                  float(p[3]) if p[3].replace(".", "").isdigit() else None, reasons))

    def summary(self):
        rows = self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        reasons = sorted({r for row in rows for r in row[3]})
        pw = [r[2] for r in rows if r[2] is not None]
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": rows[0][1],
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows), "how": self.how}


def synth_clips(n, device, seed):
    """SURVEY §8d synthetic input: 0.5 (2U - 1) + 0.25 sin(2 pi f_i n / sr), f_i = 55 * 2^((i mod 84)/12)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = 0.5 * (2 * torch.rand((n, L), generator=g, device=device) - 1)
    f = 55.0 * torch.pow(2.0, (torch.arange(n, device=device) % 84).float() / 12)
    t = torch.arange(L, device=device, dtype=torch.float32) / SR
    x += 0.25 * torch.sin(2 * math.pi * f[:, None] * t[None, :])
    return x


def cpu_port_run(n_clips, iters, seed=1234):
    """Times the oracle's torch-CPU port of the reference chain (all host threads).  Returns audio-s/s, detail."""
    import torch
    from oracle import torch_port as P
    from oracle import np_oracle as O
    import numpy as np
    x = synth_clips(n_clips, "cpu", seed)
    w = P.gaussian_window(N_FFT)
    fwd, _ = O.magnitude_banks(SR, N_FFT)
    bank = torch.from_numpy(fwd)[None]
    off, sc = torch.tensor(0.05), torch.tensor(5.0)
    P.cfg2_forward(x[:4], w, bank, off, sc)                 # warm-up (thread pool, MKL plans)
    best = float("inf")
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        P.cfg2_forward(x, w, bank, off, sc)
        dt = time.perf_counter() - t0
        times.append(dt)
        best = min(best, dt)
    return n_clips * CLIP_S / best, {"times_s": [round(t, 4) for t in times]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is pure
    Python on torch CPU ops and cannot travel to this box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    cores = torch.get_num_threads()
    n_clips = 64
    x_steps = max(1, args.steps)
    # warm-up steps are folded into cpu_port_run's own warm-up; time `steps` passes of the 64-clip sample
    t0 = time.perf_counter()
    value, detail = cpu_port_run(n_clips, x_steps)
    wall = time.perf_counter() - t0
    sample = "%d of the %d clips per step (torch-CPU port of DGT.forward incl. its angle() phase buffer + Magnitude.forward), best of %d" % (
        n_clips, CLIPS_PER_GPU, x_steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_clips * CLIP_S / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step_timed": n_clips, "where": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 2), "host": {"cpu_count": os.cpu_count(), "torch_threads": cores},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from acids_transforms_b200 import transforms as Tr
    from acids_transforms_b200 import ops, _lib
    from acids_transforms_b200.hostpipe import HostPipeline
    from acids_transforms_b200.sharding import shard_range

    _lib.load()                                   # no CUDA extension -> fail loudly, there is no fallback
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the job: world * clips clips, this rank's shard (contiguous split, no collective on the path) ----
    n_local = args.clips
    lo, hi = shard_range(world * n_local, world, rank)
    assert hi - lo == n_local
    x = synth_clips(n_local, dev, 1234 + rank)
    chain = (Tr.DGT(sr=SR, n_fft=N_FFT, hop_length=HOP, inversion_mode="random") +
             Tr.Magnitude(sr=SR, mel=True, mode="unipolar", contrast="log1p", n_fft=N_FFT)).to(dev)
    chain.scale_data(x[:16])                      # fitted once on the first 16 clips, then frozen
    stft = Tr.STFT(sr=SR, n_fft=N_FFT, hop_length=HOP).to(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- forward, inputs resident in HBM ----
    y = chain(x)
    assert tuple(y.shape) == (n_local, T, F)
    clk = ClockSampler(local_rank)
    clk.__enter__()                               # sampled across every timed region below
    fwd_ms = timed(lambda: chain(x), args.steps, args.warmup)
    fwd_step_ms = fwd_ms / args.steps
    value = world * n_local * CLIP_S / (fwd_step_ms / 1e3)
    peak, peak_src = measured_peaks()
    fwd_gbs = n_local * FWD_BYTES_PER_CLIP / (fwd_step_ms / 1e3) / 1e9

    # ---- inverse, inputs resident in HBM ----
    X = stft(x)
    del y
    inv_ms = timed(lambda: stft.invert(X), args.steps, args.warmup)
    inv_step_ms = inv_ms / args.steps
    inv_value = world * n_local * CLIP_S / (inv_step_ms / 1e3)
    inv_gbs = n_local * INV_BYTES_PER_CLIP / (inv_step_ms / 1e3) / 1e9
    del X

    # ---- end to end: pinned HOST clips in, pinned HOST features out, through the public API ----
    e2e = None
    if not args.no_e2e:
        n_e2e = min(n_local, 512)                 # 361 MB in / 725 MB out per step of pinned host memory
        x_host = torch.empty((n_e2e, L), dtype=torch.float32, pin_memory=True)
        x_host.copy_(x[:n_e2e])
        pipe = HostPipeline(chain, chunk_clips=64, device=dev)
        out_host = pipe(x_host)
        torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            pipe(x_host, out_host)
        # The host side of this path (PCIe, pinned memory, driver calls) shares the box with other tenants: time three
        # groups of e2e_steps steps and report the MEDIAN group (all three are in the line).
        groups = []
        for _ in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(e2e_steps):
                pipe(x_host, out_host)
            e1.record()
            barrier()
            groups.append(max_over_ranks(e0.elapsed_time(e1)) / e2e_steps)
        e2e_ms = sorted(groups)[1]
        e2e = {"value": world * n_e2e * CLIP_S / (e2e_ms / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "clips_per_step": n_e2e, "ms_per_step": e2e_ms,
               "ms_per_step_groups": [round(g, 3) for g in groups], "steps_per_group": e2e_steps,
               "api": "HostPipeline(DGT + Magnitude)(pinned host tensor) -> pinned host tensor"}
        del x_host, out_host

    clk.__exit__()
    clocks = clk.summary()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = torch.get_num_threads()
        v, detail = cpu_port_run(256, 5)
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": "256 of the 1024 clips (torch-CPU port of the reference chain, oracle/torch_port.py), best of 5 passes"}

    if rank == 0:
        kname = "stft_fwd_kernel<Fwd1024,MODE_REAL>"
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": fwd_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": n_local, "clip_samples": L, "frames": T, "bins": F,
                       "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2" % (n_local * 4 * L / 1e6, n_local * 4 * T * F / 1e6),
                       "parallelism": "batch sharded over %d GPU(s), no collective" % world},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": fwd_gbs, "peak": peak, "unit": "GB/s", "frac": fwd_gbs / peak,
                         "traffic": traffic_from_profiles(kname), "algorithmic_bytes_per_launch": n_local * FWD_BYTES_PER_CLIP,
                         "peak_source": peak_src},
            "inverse": {"value": inv_value, "unit": "audio-s/s", "ms_per_step": inv_step_ms,
                        "roofline": {"bound": "hbm", "kernel": "istft_ola_kernel<Inv1024>", "achieved": inv_gbs, "peak": peak,
                                     "unit": "GB/s", "frac": inv_gbs / peak, "traffic": traffic_from_profiles("istft_ola_kernel<Inv1024>"),
                                     "algorithmic_bytes_per_launch": n_local * INV_BYTES_PER_CLIP}},
            "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
