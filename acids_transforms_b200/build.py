"""Builds libacids_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m acids_transforms_b200.build [--force]

The shared library has no torch dependency: it exports the C ABI declared in include/acids_b200.h.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.environ.get("ACIDS_B200_OBJ") or os.path.join(PKG, "build")
LIB = os.environ.get("ACIDS_B200_LIB") or os.path.join(PKG, "libacids_b200.so")
SHIM = os.path.join(PKG, "libacids_b200_torch.so")       # TORCH_LIBRARY(acids_b200): csrc/torch_shim.cpp on top of the C ABI
SOURCES = ["capi.cu", "stft_fwd.cu", "istft.cu", "spectral_repr.cu", "pointwise.cu", "mfcc_tc.cu", "mel_tc.cu", "stream.cu", "pghi.cu"]
# the fused forward kernels: one translation unit per FFT plan (stft_fwd_plan.cu -DACIDS_FWD_PLAN_N=n), built in parallel
FWD_PLANS = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
HEADERS = ["common.cuh", "tc_common.cuh", "fft_core.cuh", "plans.cuh", "stft_fwd_kernel.cuh", os.path.join(ROOT, "include", "acids_b200.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    units = [(src, src.replace(".cu", ".o"), []) for src in SOURCES]
    units += [("stft_fwd_plan.cu", "stft_fwd_plan_%d.o" % n, ["-DACIDS_FWD_PLAN_N=%d" % n]) for n in FWD_PLANS]
    extra = os.environ.get("ACIDS_NVCC_EXTRA", "").split()      # e.g. -DACIDS_FWD_MINB_SMALL=3 for tuning experiments
    for src, obj, defs in units:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, obj)
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o, defs + extra))
    jobs.sort(key=lambda j: -os.path.getsize(j[0]) if "plan" not in j[1] else -10**9)   # longest (per-plan) units first

    def compile_one(job):
        s, o, defs = job
        cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(compile_one, jobs))
        if verbose:
            print("\n".join(logs))
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if os.environ.get("ACIDS_B200_LIB") is None:       # tuning variants (ACIDS_B200_LIB=...) reuse the product shim
        build_shim(force)
    return LIB


def build_shim(force=False):
    """libacids_b200_torch.so: the dispatcher registration of torch.ops.acids_b200.* in C++ (csrc/torch_shim.cpp), linked
    against libtorch and libacids_b200.so — what a libtorch-only host dlopens before torch::jit::load of a saved chain."""
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(CSRC, "torch_shim.cpp")
    hdr = os.path.join(ROOT, "include", "acids_b200.h")
    if not (force or _stale(SHIM, [src, hdr, os.path.abspath(__file__)])):
        return SHIM
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in ce.include_paths()] + ["-I" + os.path.join(cuda_home, "include")]
    cmd += [src, "-o", SHIM, "-L" + tl, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-L" + PKG, "-l:libacids_b200.so",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tl]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("torch shim build failed:\n%s\n%s" % (r.stdout[-3000:], r.stderr[-3000:]))
    return SHIM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
