"""Phase-gradient heap integration (PGHI) — the host restatement (the whole-spectrogram variant also runs on the GPU:
csrc/pghi.cu via ops.pghi, which DGT.pghi uses for device tensors and which the tests pin against this module).

Průša, Balazs, Søndergaard, "A Noniterative Method for Reconstruction of Phase from STFT Magnitude"
(IEEE/ACM TASLP 2017), as parameterised by the reference (acids_transforms/transforms/dgt.py:156-236 for
whole spectrograms, :338-466 for the frame-by-frame variant).  SURVEY.md §2 row 3 / §8(f) N4 keep it off the
GPU hot path: it is a sequential, data-dependent priority-queue flood fill.  It runs on the host in numpy with
`heapq`; the phase it returns is recombined with the magnitude and inverted by the CUDA ISTFT kernels.

Reference quirks that are reproduced because they shape the output (each is visible in the cited lines):
  * offline: steps along FRAMES integrate `fgrad` (log-magnitude derivative along bins / fmul + 2 pi hop k / n_fft),
    steps along BINS integrate `tgrad` (-fmul * derivative along frames + pi)  (dgt.py:191-224);
    bins below `tol * max` keep phase 0.
  * real time: the roles of the two gradients are swapped (dgt.py:419-441), the time derivative is the 3-point
    stencil (3 y[t+1] - 4 y[t] + y[t-1]) / 2 (dgt.py:380), the gradient rows are indexed two frames late
    (the extra zero rows of dgt.py:393-395), bin 0 is never reached from bin 1 (`> 0` at dgt.py:434), and
    low-magnitude bins get a normal random phase.
One deliberate deviation: the reference's real-time stencil reads rows of uninitialised memory (`torch.empty`,
dgt.py:373) on either side of the block.  The row after the newest frame only feeds a gradient row that is never indexed
(the two-frame lag above); the row BEFORE the first history frame feeds the gradient the first two new frames integrate,
so the reference's phases depend on what the allocator returned.  Here both rows replicate their neighbour; the golden
fixture `rtpghi_128_32` was generated with zero memory there and pins this restatement with `row_before="zeros"`.
"""
import heapq
import math
from typing import Optional, Tuple

import numpy as np
import torch


def _log_clamped(mag: np.ndarray, eps: float) -> np.ndarray:
    return np.log(np.maximum(mag, eps)).astype(np.float32)


def phase_gradients(mag: np.ndarray, gamma: float, n_fft: int, hop: int, eps: float) -> Tuple[np.ndarray, np.ndarray]:
    """(tgrad, fgrad) of a [T, F] magnitude: central differences of log|X| with replicated edges (dgt.py:226-236)."""
    fmul = np.float32(gamma / (hop * n_fft))
    y = np.pad(_log_clamped(mag, eps), 1, mode="edge")
    d_bins = (y[1:-1, 2:] - y[1:-1, :-2]) / np.float32(2)
    d_frames = (y[2:, 1:-1] - y[:-2, 1:-1]) / np.float32(2)
    k = np.arange(n_fft // 2 + 1, dtype=np.float32)[None, :]
    fgrad = d_bins / fmul + np.float32(2 * math.pi * hop / n_fft) * k
    tgrad = -fmul * d_frames + np.float32(math.pi)
    return tgrad.astype(np.float32), fgrad.astype(np.float32)


def heap_integrate(mag: np.ndarray, tgrad: np.ndarray, fgrad: np.ndarray, abstol: float, tol: float) -> np.ndarray:
    """Flood-fill the phase from the loudest bin outwards, always continuing from the loudest visited bin
    (dgt.py:167-224).  mag is [T, F] (modified copy inside); returns the phase [T, F] (float32)."""
    s = np.maximum(mag.astype(np.float32), np.float32(abstol)).copy()
    n_t, n_f = s.shape
    phase = np.zeros_like(s)
    peak = float(s.max())
    s[s < np.float32(peak * tol)] = abstol          # too quiet to trust: never visited, phase stays 0
    done = abstol
    while True:
        flat = int(np.argmax(s))
        t0, k0 = divmod(flat, n_f)
        top = float(s[t0, k0])
        if not top > abstol:
            break
        heap = [(-top, t0, k0)]
        s[t0, k0] = done
        while heap:
            _, t, k = heapq.heappop(heap)
            p = phase[t, k]
            if t + 1 < n_t and s[t + 1, k] > abstol:
                phase[t + 1, k] = p + (fgrad[t, k] + fgrad[t + 1, k]) / np.float32(2)
                heapq.heappush(heap, (-float(s[t + 1, k]), t + 1, k))
                s[t + 1, k] = done
            if t > 0 and s[t - 1, k] > abstol:
                phase[t - 1, k] = p - (fgrad[t, k] + fgrad[t - 1, k]) / np.float32(2)
                heapq.heappush(heap, (-float(s[t - 1, k]), t - 1, k))
                s[t - 1, k] = done
            if k + 1 < n_f and s[t, k + 1] > abstol:
                phase[t, k + 1] = p + (tgrad[t, k] + tgrad[t, k + 1]) / np.float32(2)
                heapq.heappush(heap, (-float(s[t, k + 1]), t, k + 1))
                s[t, k + 1] = done
            if k > 0 and s[t, k - 1] > abstol:
                phase[t, k - 1] = p - (tgrad[t, k] + tgrad[t, k - 1]) / np.float32(2)
                heapq.heappush(heap, (-float(s[t, k - 1]), t, k - 1))
                s[t, k - 1] = done
    return phase


def pghi(mag: torch.Tensor, gamma: float, n_fft: int, hop: int, tol: float, eps: float) -> torch.Tensor:
    """Phase for one [T, F] magnitude (DGT.pghi, dgt.py:156-162).  Host computation; result on mag's device."""
    m = mag.detach().to("cpu", torch.float32).numpy()
    m = np.maximum(m, np.float32(eps))
    tgrad, fgrad = phase_gradients(m, gamma, n_fft, hop, eps)
    ph = heap_integrate(m, tgrad, fgrad, eps, tol)
    return torch.from_numpy(ph).to(mag.device)


# ---------------------------------------------------------------------------------------------
# frame-by-frame variant (RealtimeDGT.pghi, dgt.py:338-466)
# ---------------------------------------------------------------------------------------------
def rt_phase_gradients(mag: np.ndarray, gamma: float, n_fft: int, hop: int, row_before: str = "edge") -> Tuple[np.ndarray, np.ndarray]:
    """mag [T, F] is already clamped and includes the two history frames (dgt.py:364-384).  `row_before`: what the stencil
    sees before the first history frame — "edge" (the product: that frame replicated) or "zeros" (what the reference reads
    when its `torch.empty` row happens to be zero memory; used to pin this restatement on the golden fixture)."""
    fmul = np.float32(gamma / (hop * n_fft))
    y = np.log(mag).astype(np.float32)
    yb = np.pad(y, ((0, 0), (1, 1)), mode="edge")
    d_bins = (yb[:, 2:] - yb[:, :-2]) / np.float32(2)
    # (3 y[t+1] - 4 y[t] + y[t-1]) / 2 with the rows outside the block replicated (see module docstring)
    yt = np.pad(y, ((1, 1), (0, 0)), mode="edge")
    if row_before == "zeros":
        yt[0] = 0
    d_frames = (np.float32(3) * yt[2:] - np.float32(4) * yt[1:-1] + yt[:-2]) / np.float32(2)
    k = np.arange(n_fft // 2 + 1, dtype=np.float32)[None, :]
    fgrad = d_bins / fmul + np.float32(2 * math.pi * hop / n_fft) * k
    tgrad = -fmul * d_frames + np.float32(math.pi)
    return tgrad.astype(np.float32), fgrad.astype(np.float32)


def rt_heap_integrate(mag: np.ndarray, prev_phase: np.ndarray, tgrad: np.ndarray, fgrad: np.ndarray, tol: float, eps: float,
                      rng: np.random.Generator, noise: Optional[np.ndarray] = None) -> np.ndarray:
    """One clip: rows 0-1 of mag are the history frames, row 1 has the known phase `prev_phase` (dgt.py:386-452).
    Every new frame is seeded from the previous frame's audible bins and from its own loudest bin."""
    s = mag.astype(np.float32).copy()
    n_t, n_f = s.shape
    abstol = max(float(tol * s.max()), eps)
    phase = np.zeros_like(s)
    phase[1] = prev_phase
    quiet = ~(s[2:] > abstol)
    phase[2:][quiet] = rng.standard_normal(int(quiet.sum())).astype(np.float32) if noise is None else noise[quiet]
    # the reference prepends two zero rows to gradients that already cover the history frames
    zeros = np.zeros((2, n_f), np.float32)
    tg = np.concatenate([zeros, tgrad], 0)
    fg = np.concatenate([zeros, fgrad], 0)
    hist = s.copy()
    for f in range(2, n_t):
        top = float(s[f].max())
        if top <= abstol:
            continue
        k0 = int(np.argmax(s[f]))
        heap = [(-top, f, k0)]
        for k in np.nonzero(hist[f - 1] > abstol)[0]:
            heapq.heappush(heap, (-float(hist[f - 1, k]), f - 1, int(k)))
        while top > abstol:
            while heap:
                top, t, k = heapq.heappop(heap)
                if t == f - 1:
                    if s[f, k] > abstol:
                        phase[f, k] = phase[t, k] + np.float32(0.5) * (tg[t, k] + tg[f, k])
                        heapq.heappush(heap, (-float(s[f, k]), f, k))
                        s[f, k] = abstol
                else:
                    if k + 1 < n_f and s[f, k + 1] > abstol:
                        phase[f, k + 1] = phase[f, k] + np.float32(0.5) * (fg[f, k] + fg[f, k + 1])
                        heapq.heappush(heap, (-float(s[f, k + 1]), f, k + 1))
                        s[f, k + 1] = abstol
                    if k - 1 > 0 and s[f, k - 1] > abstol:
                        phase[f, k - 1] = phase[f, k] - np.float32(0.5) * (fg[f, k] + fg[f, k - 1])
                        heapq.heappush(heap, (-float(s[f, k - 1]), f, k - 1))
                        s[f, k - 1] = abstol
            top = float(s[f].max())
            k0 = int(np.argmax(s[f]))
            heapq.heappush(heap, (-top, f, k0))
            s[f, k0] = abstol
    return phase[2:]


def rt_pghi(mag: torch.Tensor, hist_mag: torch.Tensor, hist_phase: torch.Tensor, gamma: float, n_fft: int, hop: int,
            tol: float, eps: float, generator: Optional[np.random.Generator] = None, row_before: str = "edge",
            noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mag [B, n, F] new frames, hist_mag [B, 2, F], hist_phase [B, F] -> phase [B, n, F].  Bins below the tolerance take
    `noise` [B, n, F] when given, else draws of `generator`."""
    rng = generator or np.random.default_rng()
    m = torch.cat([hist_mag.to(mag.device, mag.dtype), mag], -2).detach().to("cpu", torch.float32).numpy()
    m = np.maximum(m, np.float32(eps))
    hp = hist_phase.detach().to("cpu", torch.float32).numpy()
    nz = None if noise is None else noise.detach().to("cpu", torch.float32).numpy()
    out = []
    for i in range(m.shape[0]):
        tgrad, fgrad = rt_phase_gradients(m[i], gamma, n_fft, hop, row_before)
        out.append(rt_heap_integrate(m[i], hp[i], tgrad, fgrad, tol, eps, rng, None if nz is None else nz[i]))
    return torch.from_numpy(np.stack(out)).to(mag.device)
