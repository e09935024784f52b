// torch_shim.cpp — TORCH_LIBRARY(acids_b200): the dispatcher-registered ops the nn.Module mirror calls, in C++.
//
// Thin by construction: tensor -> (device pointer, sizes, current stream) -> the C ABI of include/acids_b200.h.  No
// arithmetic lives here.  Built next to libacids_b200.so as libacids_b200_torch.so (acids_transforms_b200/build.py), so a
// chain exported with torch.jit.script(...).save() loads in a libtorch-only (C++) host: the reference's stated use
// (README.md:4, :43-67; test/test_transforms.py:62-68).  The host loads it with torch::jit::load after
// dlopen("libacids_b200_torch.so"); Python loads it with torch.ops.load_library (acids_transforms_b200/_torch_ops.py).
// Schemas and semantics are those of the Python registration it replaces (ops.py is the ctypes twin the kernel tests use).
//
// Host (CPU) tensors are accepted like the reference's modules accept them: staged to the current CUDA device, processed
// there, copied back.  There is no CPU path: without a CUDA device every op raises.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/acids_b200.h"

namespace {

using at::Tensor;
using OptTensor = c10::optional<Tensor>;

void check(int rc) {
    if (rc == ACIDS_OK) return;
    const char* msg = acids_last_error();
    TORCH_CHECK(false, "acids_b200: ", msg ? msg : "unknown error", rc == ACIDS_EINVAL || rc == ACIDS_ENOTSUP ? "" : " (CUDA / workspace error)");
}

Tensor dev(const Tensor& t) {
    if (t.is_cuda()) return t;
    TORCH_CHECK(at::cuda::is_available(), "acids_transforms_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback");
    return t.to(at::Device(at::kCUDA, c10::cuda::current_device()), /*non_blocking=*/true);
}
Tensor ret(const Tensor& y, const Tensor& like) { return like.is_cuda() ? y : y.to(like.device()); }
void* stream_of(const Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.device().index()).stream(); }
const float* fptr(const OptTensor& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }

// Normalize.offset / .scale as a 1-element float32 tensor on `d` (or nothing)
OptTensor scalar(const OptTensor& t, at::Device d) {
    if (!t.has_value()) return c10::nullopt;
    TORCH_CHECK(t->numel() == 1, "normalisation buffers are not set: call scale_data() first");     // norm.py:22: offset starts as zeros(0)
    return t->detach().to(d, at::kFloat).reshape({1}).contiguous();
}

// flatten the leading dims, keep the last `event` ones (reshape_batches, utils/misc.py:168-178)
Tensor flat(const Tensor& x, int64_t event, std::vector<int64_t>& batch) {
    batch.assign(x.sizes().begin(), x.sizes().end() - event);
    std::vector<int64_t> shape{-1};
    shape.insert(shape.end(), x.sizes().end() - event, x.sizes().end());
    return x.reshape(shape).contiguous();
}
std::vector<int64_t> with_batch(const std::vector<int64_t>& batch, std::initializer_list<int64_t> tail) {
    std::vector<int64_t> s(batch);
    s.insert(s.end(), tail);
    return s;
}

// the banded (meta, coef) buffer pair of a Magnitude / MFCC module -> acids_band.  meta = (cnt, base)[ceil(n_out/32)] ++
// start[n_out] ++ (n_in, coef_len): n_out follows from the length; n_in is read back ONCE per buffer version.
struct Band {
    Tensor meta, coef;
    acids_band b{nullptr, nullptr, 0, 0, 0};
};
// Identity of a tensor's buffer for the two host-side caches below.  A raw data pointer is not one: the caching allocator hands
// a freed bank's address to the next bank of the same size (the metadata of a 513 -> 128 and of a 1025 -> 128 bank are equally
// long).  The StorageImpl's address is, as long as somebody holds a weak reference to it — which every cache entry does; a weak
// reference keeps the small StorageImpl object alive, not the device memory.
using WeakStorage = c10::weak_intrusive_ptr<c10::StorageImpl>;
using BufferId = std::tuple<const void*, int64_t, int64_t, int64_t>;       // StorageImpl, offset, numel, version
BufferId buffer_id(const Tensor& t) {
    return std::make_tuple((const void*)t.storage().unsafeGetStorageImpl(), (int64_t)t.storage_offset(), (int64_t)t.numel(), (int64_t)t._version());
}
Band band_of(const OptTensor& meta, const OptTensor& coef, at::Device d) {
    Band r;
    if (!meta.has_value() || !coef.has_value()) return r;
    static std::mutex mu;
    static std::map<BufferId, std::tuple<int, int, WeakStorage>> cache;      // -> (n_out, n_in, what pins the identity)
    const auto key = buffer_id(*meta);
    std::pair<int, int> dims;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it == cache.end()) {
            const int64_t total = meta->numel() - 2;
            int64_t n_out = -1;
            for (int64_t n = std::max<int64_t>(total - 2 * ((total + 31) / 32) - 2, 0); n <= total; ++n)
                if (n + 2 * ((n + 31) / 32) == total) { n_out = n; break; }
            TORCH_CHECK(n_out >= 0, "acids_b200: malformed banded-matrix metadata");
            const int n_in = meta->numel() >= 2 ? meta->select(0, meta->numel() - 2).item<int>() : -1;
            if (cache.size() > 64) cache.clear();
            it = cache.emplace(key, std::make_tuple((int)n_out, n_in, meta->storage().getWeakStorageImpl())).first;
        }
        dims = std::make_pair(std::get<0>(it->second), std::get<1>(it->second));
    }
    r.meta = meta->to(d, at::kInt).contiguous();
    r.coef = coef->to(d, at::kFloat).contiguous();
    r.b = acids_band{r.meta.data_ptr<int32_t>(), r.coef.data_ptr<float>(), dims.first, (int32_t)r.coef.numel(), dims.second};
    return r;
}

Tensor as_c64(const Tensor& X) {
    Tensor x = X;
    if (!x.is_complex()) x = x.to(at::kFloat);
    return x.to(at::kComplexFloat).resolve_conj();
}

void check_stft_input(const Tensor& x, int64_t n_fft) {
    TORCH_CHECK(x.scalar_type() == at::kFloat, "acids_b200: expected a float32 waveform, got ", x.scalar_type());
    TORCH_CHECK(x.dim() >= 1 && n_fft / 2 > 0 && n_fft / 2 < x.size(-1),
                "Argument #4: Padding size should be less than the corresponding input dimension, but got: padding (", n_fft / 2, ", ",
                n_fft / 2, ") at dimension 2 of input ", x.sizes());
}

at::TensorOptions f32(const Tensor& like) { return like.options().dtype(at::kFloat); }

// ---- (1) STFT forward: stft.py:97-104, dgt.py:63-70; pre-framed input stft.py:248-253 ----
Tensor stft_fwd(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, bool center) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    std::vector<int64_t> batch;
    int64_t B, L, T, hop_k;
    Tensor xf;
    if (center) {
        check_stft_input(xd, n_fft);
        xf = flat(xd, 1, batch);
        B = xf.size(0); L = xf.size(1); T = 1 + L / hop; hop_k = hop;
    } else {
        TORCH_CHECK(xd.size(-1) == n_fft, "acids_b200: pre-framed input must have n_fft=", n_fft, " samples per frame, got ", xd.size(-1));
        TORCH_CHECK(xd.scalar_type() == at::kFloat, "acids_b200: expected float32 frames");
        Tensor rows = flat(xd, 1, batch);
        T = rows.size(0);
        xf = rows.reshape({1, -1});
        B = 1; L = xf.size(1); hop_k = n_fft;
    }
    const int64_t F = n_fft / 2 + 1;
    Tensor w = dev(window).to(at::kFloat).contiguous();
    Tensor out = at::empty({B, T, F}, xf.options().dtype(at::kComplexFloat));
    if (out.numel())
        check(acids_stft_fwd(xf.data_ptr<float>(), B, L, L, w.data_ptr<float>(), (int)n_fft, (int)hop_k, center ? 1 : 0, T,
                             reinterpret_cast<float*>(out.data_ptr()), stream_of(xf)));
    return ret(out.reshape(center ? with_batch(batch, {T, F}) : with_batch(batch, {F})), x);
}

// MidSide.forward -> STFT.forward in one kernel: raw.py:145-161 + stft.py:101-102
Tensor midside_stft_fwd(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, int64_t ms) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    check_stft_input(xd, n_fft);
    TORCH_CHECK(xd.dim() >= 2 && xd.size(-2) == 2, "acids_b200: the fused MidSide prologue needs a stereo [..., 2, L] input");
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 1, batch);
    const int64_t B = xf.size(0), L = xf.size(1), T = 1 + L / hop, F = n_fft / 2 + 1;
    Tensor w = dev(window).to(at::kFloat).contiguous();
    Tensor out = at::empty({B, T, F}, xf.options().dtype(at::kComplexFloat));
    if (out.numel())
        check(acids_midside_stft_fwd(xf.data_ptr<float>(), B, L, w.data_ptr<float>(), (int)n_fft, (int)hop, T, (int)ms,
                                     reinterpret_cast<float*>(out.data_ptr()), stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {T, F})), x);
}

// ---- (2) fused STFT + Magnitude: stft.py:101 + spectral_repr.py:215-226 ----
Tensor stft_mag_fwd(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, const OptTensor& band_meta,
                    const OptTensor& band_coef, int64_t contrast, double eps, const OptTensor& offset, const OptTensor& scale,
                    bool drop_first) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    check_stft_input(xd, n_fft);
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 1, batch);
    const int64_t B = xf.size(0), L = xf.size(1), T = 1 + L / hop, F = n_fft / 2 + 1;
    Band band = band_of(band_meta, band_coef, xf.device());
    TORCH_CHECK(!band.b.meta || band.b.n_in == F, "mat1 and mat2 shapes cannot be multiplied (", B * T, "x", F, " and ", band.b.n_in, "x", band.b.n_out, ")");
    const int64_t n_keep = (band.b.meta ? band.b.n_out : F) - (drop_first ? 1 : 0);
    Tensor w = dev(window).to(at::kFloat).contiguous();
    Tensor out = at::empty({B, T, n_keep}, f32(xf));
    OptTensor off = scalar(offset, xf.device()), sc = scalar(scale, xf.device());
    if (out.numel())
        check(acids_stft_mag_fwd(xf.data_ptr<float>(), B, L, L, w.data_ptr<float>(), (int)n_fft, (int)hop, 1, T, band.b, (int)contrast,
                                 (float)eps, fptr(off), fptr(sc), drop_first ? 1 : 0, out.data_ptr<float>(), T * n_keep, n_keep, stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {T, n_keep})), x);
}

Tensor mag_epilogue_into(const Tensor& Xf, const Band& band, int64_t contrast, double eps, const OptTensor& offset,
                         const OptTensor& scale, bool drop_first, Tensor out, int64_t slot, int64_t slots) {
    const int64_t rows = Xf.size(0), F = Xf.size(1);
    const int64_t n_keep = (band.b.meta ? band.b.n_out : F) - (drop_first ? 1 : 0);
    OptTensor off = scalar(offset, Xf.device()), sc = scalar(scale, Xf.device());
    if (out.numel())
        check(acids_mag_epilogue(reinterpret_cast<const float*>(Xf.data_ptr()), rows, (int)F, band.b, (int)contrast, (float)eps, fptr(off),
                                 fptr(sc), drop_first ? 1 : 0, out.data_ptr<float>() + slot * n_keep, slots * n_keep, stream_of(Xf)));
    return out;
}

// Magnitude.forward on a spectrum: spectral_repr.py:215-226
Tensor mag_epilogue(const Tensor& X, const OptTensor& band_meta, const OptTensor& band_coef, int64_t contrast, double eps,
                    const OptTensor& offset, const OptTensor& scale, bool drop_first) {
    Tensor Xd = as_c64(dev(X));
    c10::cuda::CUDAGuard guard(Xd.device());
    std::vector<int64_t> batch;
    Tensor Xf = flat(Xd, 1, batch);
    const int64_t rows = Xf.size(0), F = Xf.size(1);
    Band band = band_of(band_meta, band_coef, Xf.device());
    TORCH_CHECK(!band.b.meta || band.b.n_in == F, "mat1 and mat2 shapes cannot be multiplied (", rows, "x", F, " and ", band.b.n_in, "x", band.b.n_out, ")");
    const int64_t n_keep = (band.b.meta ? band.b.n_out : F) - (drop_first ? 1 : 0);
    Tensor out = at::empty({rows, n_keep}, f32(Xf));
    mag_epilogue_into(Xf, band, contrast, eps, offset, scale, drop_first, out, 0, 1);
    return ret(out.reshape(with_batch(batch, {n_keep})), X);
}

// Magnitude.invert: spectral_repr.py:228-240
Tensor mag_invert(const Tensor& y, const OptTensor& band_meta, const OptTensor& band_coef, int64_t contrast, double eps,
                  const OptTensor& offset, const OptTensor& scale, bool pad_last) {
    Tensor yd = dev(y).to(at::kFloat);
    c10::cuda::CUDAGuard guard(yd.device());
    std::vector<int64_t> batch(yd.sizes().begin(), yd.sizes().end() - 1);
    Tensor yf = yd.reshape({-1, yd.size(-1)});
    if (yf.stride(-1) != 1) yf = yf.contiguous();
    const int64_t rows = yf.size(0), n_in = yf.size(1), n_val = n_in + (pad_last ? 1 : 0);
    Band band = band_of(band_meta, band_coef, yf.device());
    TORCH_CHECK(!band.b.meta || band.b.n_in == n_val, "mat1 and mat2 shapes cannot be multiplied (", rows, "x", n_val, " and ", band.b.n_in, "x", band.b.n_out, ")");
    const int64_t n_out = band.b.meta ? band.b.n_out : n_val;
    Tensor out = at::empty({rows, n_out}, f32(yf));
    OptTensor off = scalar(offset, yf.device()), sc = scalar(scale, yf.device());
    if (out.numel())
        check(acids_mag_invert(yf.data_ptr<float>(), rows, (int)n_in, rows > 1 ? yf.stride(0) : n_in, pad_last ? 1 : 0, band.b, (int)contrast, (float)eps,
                               fptr(off), fptr(sc), out.data_ptr<float>(), stream_of(yf)));
    return ret(out.reshape(with_batch(batch, {n_out})), y);
}

// MFCC.forward (= torchaudio MelSpectrogram): mel.py:68-73
Tensor melspec_fwd(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, const Tensor& band_meta, const Tensor& band_coef,
                   double power, const OptTensor& offset, const OptTensor& scale) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    check_stft_input(xd, n_fft);
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 1, batch);
    const int64_t B = xf.size(0), L = xf.size(1), T = 1 + L / hop;
    Band mel = band_of(band_meta, band_coef, xf.device());
    Tensor w = dev(window).to(at::kFloat).contiguous();
    Tensor out = at::empty({B, mel.b.n_out, T}, f32(xf));
    OptTensor off = scalar(offset, xf.device()), sc = scalar(scale, xf.device());
    if (out.numel())
        check(acids_melspec_fwd(xf.data_ptr<float>(), B, L, L, w.data_ptr<float>(), (int)n_fft, (int)hop, T, mel.b, (float)power, fptr(off),
                                fptr(sc), out.data_ptr<float>(), stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {mel.b.n_out, T})), x);
}

// dB + top_db floor + DCT-II (torchaudio MFCC, _transforms.py:701-718); top_db < 0: no floor
Tensor mfcc_dct(const Tensor& mel, const Tensor& dct, double top_db) {
    Tensor md = dev(mel).to(at::kFloat);
    c10::cuda::CUDAGuard guard(md.device());
    std::vector<int64_t> batch;
    Tensor mf = flat(md, 2, batch);
    const int64_t B = mf.size(0), n_mels = mf.size(1), T = mf.size(2);
    // torchaudio's amplitude_to_DB packs dim -3 as "channels": inputs with <= 3 dims share ONE max
    const int64_t group = std::max<int64_t>(md.dim() <= 3 ? B : md.size(-3), 1);
    Tensor d = dev(dct).to(md.device(), at::kFloat).contiguous();
    const int64_t n_mfcc = d.size(1);
    Tensor out = at::empty({B, n_mfcc, T}, f32(mf));
    Tensor gmax = at::empty({std::max<int64_t>(B / group, 1)}, f32(mf));
    const bool tc = n_mels % 8 == 0 && n_mels <= 128 && n_mfcc <= 48;
    if (out.numel())
        check((tc ? acids_mfcc_dct_tc : acids_mfcc_dct)(mf.data_ptr<float>(), B, (int)n_mels, T, d.data_ptr<float>(), (int)n_mfcc,
                                                        (float)(top_db < 0 ? -1.0 : top_db), group, gmax.data_ptr<float>(),
                                                        out.data_ptr<float>(), stream_of(mf)));
    return ret(out.reshape(with_batch(batch, {n_mfcc, T})), mel);
}

// ---- (3) phase / IF: spectral_repr.py:270-278, :319-357; utils/misc.py:12-26 ----
Tensor spectrum3(const Tensor& X, std::vector<int64_t>& batch) {
    Tensor Xd = as_c64(dev(X));
    TORCH_CHECK_INDEX(Xd.dim() >= 2, "Dimension out of range (expected a [..., frames, bins] spectrum)");
    return flat(Xd, 2, batch);
}

void phase_fwd_into(const Tensor& Xf, int64_t mode, int64_t method, bool weighted, const OptTensor& offset, const OptTensor& scale,
                    bool drop_first, Tensor out, int64_t slot, int64_t slots) {
    const int64_t B = Xf.size(0), T = Xf.size(1), F = Xf.size(2), n_keep = F - (drop_first ? 1 : 0);
    OptTensor off = scalar(offset, Xf.device()), sc = scalar(scale, Xf.device());
    if (out.numel())
        check(acids_phase_fwd(reinterpret_cast<const float*>(Xf.data_ptr()), B, T, (int)F, (int)mode, (int)method, weighted ? 1 : 0, fptr(off),
                              fptr(sc), drop_first ? 1 : 0, out.data_ptr<float>() + slot * n_keep, T * slots * n_keep, slots * n_keep, stream_of(Xf)));
}

Tensor phase_fwd(const Tensor& X, int64_t mode, int64_t method, bool weighted, const OptTensor& offset, const OptTensor& scale, bool drop_first) {
    std::vector<int64_t> batch;
    Tensor Xf = spectrum3(X, batch);
    c10::cuda::CUDAGuard guard(Xf.device());
    const int64_t T = Xf.size(1), n_keep = Xf.size(2) - (drop_first ? 1 : 0);
    Tensor out = at::empty({Xf.size(0), T, n_keep}, f32(Xf));
    phase_fwd_into(Xf, mode, method, weighted, offset, scale, drop_first, out, 0, 1);
    return ret(out.reshape(with_batch(batch, {T, n_keep})), X);
}

// SpectralRepresentation.forward for Polar / PolarIF with stack = -2: spectral_repr.py:431-440
Tensor polar_fwd(const Tensor& X, const OptTensor& band_meta, const OptTensor& band_coef, int64_t contrast, double eps,
                 const OptTensor& mag_offset, const OptTensor& mag_scale, int64_t phase_mode, int64_t method, bool weighted,
                 const OptTensor& ph_offset, const OptTensor& ph_scale, bool drop_first) {
    std::vector<int64_t> batch;
    Tensor Xf = spectrum3(X, batch);
    c10::cuda::CUDAGuard guard(Xf.device());
    const int64_t B = Xf.size(0), T = Xf.size(1), F = Xf.size(2), n_ph = F - (drop_first ? 1 : 0);
    Band band = band_of(band_meta, band_coef, Xf.device());
    Tensor out;
    if (!band.b.meta) {          // no mel bank: both halves from one read of the spectrum
        out = at::empty({B, T, 2, n_ph}, f32(Xf));
        OptTensor mo = scalar(mag_offset, Xf.device()), ms = scalar(mag_scale, Xf.device());
        OptTensor po = scalar(ph_offset, Xf.device()), ps = scalar(ph_scale, Xf.device());
        if (out.numel())
            check(acids_polar_fwd(reinterpret_cast<const float*>(Xf.data_ptr()), B, T, (int)F, (int)contrast, (float)eps, fptr(mo), fptr(ms),
                                  (int)phase_mode, (int)method, weighted ? 1 : 0, fptr(po), fptr(ps), drop_first ? 1 : 0, out.data_ptr<float>(),
                                  T * 2 * n_ph, 2 * n_ph, out.data_ptr<float>() + n_ph, T * 2 * n_ph, 2 * n_ph, stream_of(Xf)));
    } else {
        TORCH_CHECK(band.b.n_in == F, "mat1 and mat2 shapes cannot be multiplied (", B * T, "x", F, " and ", band.b.n_in, "x", band.b.n_out, ")");
        const int64_t n_mag = band.b.n_out - (drop_first ? 1 : 0);
        TORCH_CHECK(n_mag == n_ph, "stack expects each tensor to be equal size, but got [", n_mag, "] and [", n_ph, "] bins");
        out = at::empty({B, T, 2, n_ph}, f32(Xf));
        // where the row-tile kernel pays on B200 (ops.polar_rows_pays): raw phase, rows of 545 .. 4352 bins
        if (phase_mode == ACIDS_PHASE_RAW && F > 544 && F <= 4352) {
            // mel bank + raw phase: one row-tile kernel, one read of the spectrum
            OptTensor mo = scalar(mag_offset, Xf.device()), ms = scalar(mag_scale, Xf.device());
            OptTensor po = scalar(ph_offset, Xf.device()), ps = scalar(ph_scale, Xf.device());
            if (out.numel())
                check(acids_polar_rows_fwd(reinterpret_cast<const float*>(Xf.data_ptr()), B, T, (int)F, band.b, (int)contrast, (float)eps, fptr(mo),
                                           fptr(ms), (int)phase_mode, (int)method, weighted ? 1 : 0, fptr(po), fptr(ps), drop_first ? 1 : 0,
                                           out.data_ptr<float>(), 2 * n_ph, out.data_ptr<float>() + n_ph, 2 * n_ph, stream_of(Xf)));
        } else {
            mag_epilogue_into(Xf.reshape({B * T, F}), band, contrast, eps, mag_offset, mag_scale, drop_first, out, 0, 2);
            phase_fwd_into(Xf, phase_mode, method, weighted, ph_offset, ph_scale, drop_first, out, 1, 2);
        }
    }
    return ret(out.reshape(with_batch(batch, {T, 2, n_ph})), X);
}

Tensor midside(const Tensor& x, bool pad_mid, bool inverse);

// [MidSide ->] STFT -> Polar / PolarIF: raw.py:145-162, stft.py:101-102, spectral_repr.py:431-440
Tensor stft_polar_fwd(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, const OptTensor& band_meta,
                      const OptTensor& band_coef, int64_t contrast, double eps, const OptTensor& mag_offset, const OptTensor& mag_scale,
                      int64_t phase_mode, int64_t method, bool weighted, const OptTensor& ph_offset, const OptTensor& ph_scale,
                      bool drop_first, int64_t ms) {
    const bool fusable = phase_mode == ACIDS_PHASE_RAW || (phase_mode == ACIDS_PHASE_IF && method == ACIDS_IF_FORWARD);
    if (!fusable) {              // phase modes that need a scan over the frames: spectrum once, two representation kernels
        return polar_fwd(ms ? midside_stft_fwd(x, window, n_fft, hop, ms) : stft_fwd(x, window, n_fft, hop, true), band_meta, band_coef, contrast, eps, mag_offset, mag_scale, phase_mode,
                         method, weighted, ph_offset, ph_scale, drop_first);
    }
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    check_stft_input(xd, n_fft);
    TORCH_CHECK(!ms || (xd.dim() >= 2 && xd.size(-2) == 2), "acids_b200: the fused MidSide prologue needs a stereo [..., 2, L] input");
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 1, batch);
    const int64_t B = xf.size(0), L = xf.size(1), T = 1 + L / hop, F = n_fft / 2 + 1;
    Band band = band_of(band_meta, band_coef, xf.device());
    TORCH_CHECK(!band.b.meta || band.b.n_in == F, "mat1 and mat2 shapes cannot be multiplied (", B * T, "x", F, " and ", band.b.n_in, "x", band.b.n_out, ")");
    const int64_t n_keep = F - (drop_first ? 1 : 0), n_mag = (band.b.meta ? band.b.n_out : F) - (drop_first ? 1 : 0);
    TORCH_CHECK(n_mag == n_keep, "stack expects each tensor to be equal size, but got [", n_mag, "] and [", n_keep, "] bins");
    Tensor w = dev(window).to(at::kFloat).contiguous();
    Tensor out = at::empty({B, T, 2, n_keep}, f32(xf));
    OptTensor mo = scalar(mag_offset, xf.device()), msc = scalar(mag_scale, xf.device());
    OptTensor po = scalar(ph_offset, xf.device()), ps = scalar(ph_scale, xf.device());
    if (out.numel())
        check(acids_stft_polar_fwd(xf.data_ptr<float>(), B, L, L, w.data_ptr<float>(), (int)n_fft, (int)hop, T, (int)ms, band.b, (int)contrast,
                                   (float)eps, fptr(mo), fptr(msc), (int)phase_mode, (int)method, weighted ? 1 : 0, fptr(po), fptr(ps),
                                   drop_first ? 1 : 0, out.data_ptr<float>(), T * 2 * n_keep, 2 * n_keep, out.data_ptr<float>() + n_keep,
                                   T * 2 * n_keep, 2 * n_keep, stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {T, 2, n_keep})), x);
}

// y [..., T, n_in] as [B, T, n_in] with a unit last stride (a slot of a stacked tensor stays a view)
Tensor rows3(const Tensor& y) {
    Tensor yd = dev(y).to(at::kFloat);
    TORCH_CHECK_INDEX(yd.dim() >= 2, "Dimension out of range (expected a [..., frames, bins] tensor)");
    Tensor yf = yd.reshape({-1, yd.size(-2), yd.size(-1)});
    if (yf.stride(-1) != 1 || (yf.size(0) > 1 && yf.stride(0) < yf.size(1) * yf.stride(1))) yf = yf.contiguous();
    return yf;
}

// Phase.invert / IF.invert: spectral_repr.py:46-53, :359-375
Tensor phase_inv(const Tensor& y, int64_t mode, int64_t method, const OptTensor& offset, const OptTensor& scale, bool pad_last) {
    Tensor yf = rows3(y);
    c10::cuda::CUDAGuard guard(yf.device());
    const int64_t B = yf.size(0), T = yf.size(1), n_in = yf.size(2), n_out = n_in + (pad_last ? 1 : 0);
    Tensor out = at::empty({B, T, n_out}, f32(yf));
    OptTensor off = scalar(offset, yf.device()), sc = scalar(scale, yf.device());
    if (out.numel())
        check(acids_phase_inv(yf.data_ptr<float>(), B, T, (int)n_in, B > 1 ? yf.stride(0) : T * yf.stride(1), yf.stride(1), pad_last ? 1 : 0, (int)mode,
                              (int)method, fptr(off), fptr(sc), out.data_ptr<float>(), stream_of(yf)));
    std::vector<int64_t> shape(y.sizes().begin(), y.sizes().end() - 1);
    shape.push_back(n_out);
    return ret(out.reshape(shape), y);
}

// mag * exp(i phase): spectral_repr.py:452
Tensor polar_to_complex(const Tensor& mag, const Tensor& phase) {
    Tensor md = dev(mag).to(at::kFloat), pd = dev(phase).to(md.device(), at::kFloat);
    c10::cuda::CUDAGuard guard(md.device());
    if (md.sizes() != pd.sizes()) {
        auto bt = at::broadcast_tensors({md, pd});
        md = bt[0];
        pd = bt[1];
    }
    md = md.contiguous();
    pd = pd.contiguous();
    Tensor out = at::empty(md.sizes(), md.options().dtype(at::kComplexFloat));
    if (out.numel())
        check(acids_polar_to_complex(md.data_ptr<float>(), pd.data_ptr<float>(), md.numel(), reinterpret_cast<float*>(out.data_ptr()), stream_of(md)));
    return ret(out, mag);
}

// SpectralRepresentation.invert, spectral_repr.py:447-452, without the phase round trip through memory
Tensor phase_inv_polar(const Tensor& y, const Tensor& mag, int64_t mode, int64_t method, const OptTensor& offset, const OptTensor& scale,
                       bool pad_last) {
    if (mode == ACIDS_PHASE_IF && method == ACIDS_IF_CENTRAL) return polar_to_complex(mag, phase_inv(y, mode, method, offset, scale, pad_last));
    Tensor yf = rows3(y);
    c10::cuda::CUDAGuard guard(yf.device());
    const int64_t B = yf.size(0), T = yf.size(1), n_in = yf.size(2), n_out = n_in + (pad_last ? 1 : 0);
    std::vector<int64_t> shape(y.sizes().begin(), y.sizes().end() - 1);
    shape.push_back(n_out);
    Tensor md = dev(mag).to(yf.device(), at::kFloat);
    if (md.sizes() != at::IntArrayRef(shape)) md = md.expand(shape);          // raises like the reference's broadcast would
    md = md.contiguous();
    Tensor out = at::empty(shape, yf.options().dtype(at::kComplexFloat));
    OptTensor off = scalar(offset, yf.device()), sc = scalar(scale, yf.device());
    if (out.numel())
        check(acids_phase_inv_polar(yf.data_ptr<float>(), B, T, (int)n_in, B > 1 ? yf.stride(0) : T * yf.stride(1), yf.stride(1), pad_last ? 1 : 0,
                                    (int)mode, (int)method, fptr(off), fptr(sc), md.data_ptr<float>(), reinterpret_cast<float*>(out.data_ptr()),
                                    stream_of(yf)));
    return ret(out, y);
}

// one fast-Griffin-Lim update: torchaudio functional.py:336-350
Tensor griffinlim_update(const Tensor& rebuilt, const Tensor& tprev, const Tensor& mag, double momentum) {
    Tensor rd = as_c64(dev(rebuilt)).contiguous(), td = as_c64(dev(tprev)).contiguous(), md = dev(mag).to(at::kFloat).contiguous();
    c10::cuda::CUDAGuard guard(rd.device());
    TORCH_CHECK(rd.sizes() == td.sizes() && rd.sizes() == md.sizes(), "griffinlim_update: rebuilt ", rd.sizes(), ", tprev ", td.sizes(), " and mag ",
                md.sizes(), " must have the same shape");
    const int64_t n = rd.numel();
    if (n & 1) {                 // odd element count: process a padded flat copy
        auto pad = [](const Tensor& t) { return at::cat({t.reshape({-1}), at::zeros({1}, t.options())}); };
        return griffinlim_update(pad(rd), pad(td), pad(md), momentum).slice(0, 0, n).reshape(rd.sizes());
    }
    Tensor out = at::empty(rd.sizes(), rd.options());
    if (n)
        check(acids_griffinlim_update(reinterpret_cast<const float*>(rd.data_ptr()), reinterpret_cast<const float*>(td.data_ptr()), md.data_ptr<float>(),
                                      (float)momentum, n, reinterpret_cast<float*>(out.data_ptr()), stream_of(rd)));
    return ret(out, rebuilt);
}

// ---- (4) inverse: stft.py:119-128, dgt.py:85-93 ----
// torch.istft's `window overlap add min` check (_refs/__init__.py:3794-3797) depends only on the window: evaluated on the host
// ONCE per window buffer version (a 64 KB read-back at the first invert), never again — invert stays graph-capturable.
bool envelope_ok(const Tensor& window, int64_t n_fft, int64_t hop, int64_t n_frames) {
    const int64_t t_eff = std::min<int64_t>(n_frames, 2 * ((n_fft + hop - 1) / hop) + 2);
    static std::mutex mu;
    static std::map<std::tuple<BufferId, int64_t, int64_t, int64_t>, std::pair<bool, WeakStorage>> cache;
    const auto key = std::make_tuple(buffer_id(window), n_fft, hop, t_eff);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second.first;
    }
    Tensor w = window.detach().to(at::kCPU, at::kDouble).slice(0, 0, n_fft).contiguous();
    const double* wp = w.data_ptr<double>();
    const int64_t length = n_fft + hop * (t_eff - 1);
    std::vector<double> env(length, 0.0);
    for (int64_t t = 0; t < t_eff; ++t)
        for (int64_t n = 0; n < n_fft; ++n) env[t * hop + n] += wp[n] * wp[n];
    bool ok = true;
    for (int64_t i = n_fft / 2; i < length - n_fft / 2; ++i) ok = ok && std::fabs(env[i]) > 1e-11;
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 256) cache.clear();
    cache.emplace(key, std::make_pair(ok, window.storage().getWeakStorageImpl()));
    return ok;
}

Tensor istft_ola(const Tensor& X, const Tensor& window, int64_t n_fft, int64_t hop) {
    Tensor Xd = as_c64(dev(X));
    c10::cuda::CUDAGuard guard(Xd.device());
    std::vector<int64_t> batch;
    Tensor Xf = flat(Xd, 2, batch);
    const int64_t B = Xf.size(0), T = Xf.size(1), F = Xf.size(2);
    TORCH_CHECK(F == n_fft / 2 + 1, "istft: expected ", n_fft / 2 + 1, " frequency bins for n_fft=", n_fft, ", got ", F);
    TORCH_CHECK(envelope_ok(window, n_fft, hop, T), "istft(CUDA): window overlap add min: 1");
    Tensor w = dev(window).to(Xf.device(), at::kFloat).contiguous();
    Tensor out = at::empty({B, hop * (T - 1)}, f32(Xf));
    const int64_t ws_bytes = acids_istft_workspace_bytes(B, T, (int)n_fft, (int)hop);
    Tensor ws = ws_bytes ? at::empty({ws_bytes}, Xf.options().dtype(at::kByte)) : Tensor();
    if (out.numel())
        check(acids_istft_ola(reinterpret_cast<const float*>(Xf.data_ptr()), B, T, (int)n_fft, (int)hop, w.data_ptr<float>(), out.data_ptr<float>(),
                              ws_bytes ? ws.data_ptr() : nullptr, ws_bytes, stream_of(Xf)));
    return ret(out.reshape(with_batch(batch, {hop * (T - 1)})), X);
}

// RealtimeSTFT.invert complex branch: stft.py:259-266
Tensor irfft_frames(const Tensor& X, const Tensor& window, int64_t n_fft) {
    Tensor Xd = as_c64(dev(X));
    c10::cuda::CUDAGuard guard(Xd.device());
    std::vector<int64_t> batch;
    Tensor Xf = flat(Xd, 1, batch);
    const int64_t rows = Xf.size(0), F = Xf.size(1);
    TORCH_CHECK(F == n_fft / 2 + 1, "irfft: expected ", n_fft / 2 + 1, " frequency bins for n_fft=", n_fft, ", got ", F);
    Tensor w = dev(window).to(Xf.device(), at::kFloat).contiguous();
    Tensor out = at::empty({rows, n_fft}, f32(Xf));
    if (out.numel())
        check(acids_irfft_frames(reinterpret_cast<const float*>(Xf.data_ptr()), rows, (int)n_fft, w.data_ptr<float>(), out.data_ptr<float>(), stream_of(Xf)));
    return ret(out.reshape(with_batch(batch, {n_fft})), X);
}

// OverlapAdd.invert: oadd.py:91-104
std::tuple<Tensor, Tensor> ola_stream(const Tensor& frames, int64_t hop, int64_t keep, const OptTensor& carry_in, double gain) {
    Tensor fd = dev(frames).to(at::kFloat);
    c10::cuda::CUDAGuard guard(fd.device());
    std::vector<int64_t> batch;
    Tensor ff = flat(fd, 2, batch);
    const int64_t B = ff.size(0), n = ff.size(1), N = ff.size(2), total = (n - 1) * hop + N;
    Tensor out = at::empty({B, total - keep}, f32(ff)), carry_out = at::empty({B, keep}, f32(ff));
    Tensor ci;
    if (carry_in.has_value()) ci = dev(*carry_in).to(ff.device(), at::kFloat).reshape({B, keep}).contiguous();
    if (out.numel())
        check(acids_ola_stream(ff.data_ptr<float>(), B, n, (int)N, (int)hop, keep, ci.defined() ? ci.data_ptr<float>() : nullptr, (float)gain,
                               out.data_ptr<float>(), carry_out.data_ptr<float>(), stream_of(ff)));
    return std::make_tuple(ret(out.reshape(with_batch(batch, {total - keep})), frames), ret(carry_out.reshape(with_batch(batch, {keep})), frames));
}

// ---- (5) mu-law / one-hot: raw.py:280-316, misc.py:176-179 ----
float log1p_mu(int64_t channels) {
    // the reference evaluates torch.log1p(torch.tensor(mu)) on the host in float32 (functional.py:697-698)
    return at::log1p(at::scalar_tensor((double)channels - 1.0, at::TensorOptions().dtype(at::kFloat))).item<float>();
}

Tensor mulaw_encode(const Tensor& x, int64_t channels, int64_t one_hot) {
    // match the eager chain of the device the caller's data lives on: CUDA eager multiplies by the reciprocal of a host
    // scalar, CPU eager divides (mu-law)
    const bool reciprocal = x.is_cuda();
    Tensor xd = dev(x).to(at::kFloat).contiguous();
    c10::cuda::CUDAGuard guard(xd.device());
    const int64_t L = xd.dim() ? xd.size(-1) : 1, outer = xd.numel() / std::max<int64_t>(L, 1);
    std::vector<int64_t> shape(xd.sizes().begin(), xd.sizes().end());
    if (one_hot == ACIDS_ONEHOT_CATEGORICAL) shape.push_back(channels);
    else if (one_hot == ACIDS_ONEHOT_CHANNEL) shape.insert(shape.end() - 1, channels);
    else one_hot = ACIDS_ONEHOT_NONE;
    Tensor out = at::empty(shape, xd.options().dtype(at::kLong));
    if (out.numel())
        check(acids_mulaw_encode(xd.data_ptr<float>(), outer, L, (int)channels, log1p_mu(channels), reciprocal ? 1 : 0, (int)one_hot,
                                 out.data_ptr<int64_t>(), stream_of(xd)));
    return ret(out, x);
}

Tensor mulaw_decode(const Tensor& q, int64_t channels) {
    const bool reciprocal = q.is_cuda();
    Tensor qd = dev(q).to(at::kLong).contiguous();
    c10::cuda::CUDAGuard guard(qd.device());
    Tensor out = at::empty(qd.sizes(), qd.options().dtype(at::kFloat));
    if (qd.numel())
        check(acids_mulaw_decode(qd.data_ptr<int64_t>(), qd.numel(), (int)channels, log1p_mu(channels), reciprocal ? 1 : 0, out.data_ptr<float>(),
                                 stream_of(qd)));
    return ret(out, q);
}

Tensor one_hot(const Tensor& q, int64_t n_classes) {
    Tensor qd = dev(q);
    TORCH_CHECK(qd.scalar_type() == at::kLong, "one_hot is only applicable to index tensor of type LongTensor.");
    TORCH_CHECK(n_classes >= 1, "one_hot: n_classes is not set (call scale_data first)");
    qd = qd.contiguous();
    c10::cuda::CUDAGuard guard(qd.device());
    std::vector<int64_t> shape(qd.sizes().begin(), qd.sizes().end());
    shape.push_back(n_classes);
    Tensor out = at::empty(shape, qd.options());
    if (out.numel()) check(acids_one_hot(qd.data_ptr<int64_t>(), qd.numel(), (int)n_classes, out.data_ptr<int64_t>(), stream_of(qd)));
    return ret(out, q);
}

// ---- statistics (norm.py:26-38, spectral_repr.py:242-245), raw-domain prologues (raw.py:34-49, :145-180) ----
Tensor stats(const Tensor& x, int64_t contrast, double eps, bool abs_contrast) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    int kind;
    if (xd.is_complex()) {
        xd = xd.to(at::kComplexFloat).resolve_conj().contiguous();
        kind = ACIDS_STATS_CABS_CONTRAST;
    } else {
        xd = xd.to(at::kFloat).contiguous();
        kind = abs_contrast ? ACIDS_STATS_ABS_CONTRAST : ACIDS_STATS_REAL;
    }
    TORCH_CHECK(xd.numel() > 0, "min(): Expected reduction dim to be specified for input.numel() == 0.");
    Tensor scratch = at::empty({acids_stats_scratch_bytes()}, xd.options().dtype(at::kByte));
    Tensor out = at::empty({4}, xd.options().dtype(at::kDouble));
    check(acids_stats(reinterpret_cast<const float*>(xd.data_ptr()), xd.numel(), kind, (int)contrast, (float)eps, scratch.data_ptr(),
                      out.data_ptr<double>(), stream_of(xd)));
    return ret(out, x);
}

Tensor stft_stats(const Tensor& x, const Tensor& window, int64_t n_fft, int64_t hop, int64_t contrast, double eps) {
    Tensor xd = dev(x);
    c10::cuda::CUDAGuard guard(xd.device());
    check_stft_input(xd, n_fft);
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 1, batch);
    const int64_t B = xf.size(0), L = xf.size(1), T = 1 + L / hop;
    TORCH_CHECK(B > 0, "min(): Expected reduction dim to be specified for input.numel() == 0.");
    Tensor w = dev(window).to(xf.device(), at::kFloat).contiguous();
    Tensor scratch = at::empty({acids_stats_scratch_bytes()}, xf.options().dtype(at::kByte));
    Tensor out = at::empty({4}, xf.options().dtype(at::kDouble));
    check(acids_stft_stats(xf.data_ptr<float>(), B, L, L, w.data_ptr<float>(), (int)n_fft, (int)hop, T, (int)contrast, (float)eps, scratch.data_ptr(),
                           out.data_ptr<double>(), stream_of(xf)));
    return ret(out, x);
}

Tensor mono_mix(const Tensor& x) {
    Tensor xd = dev(x).to(at::kFloat);
    c10::cuda::CUDAGuard guard(xd.device());
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 2, batch);
    TORCH_CHECK(xf.size(1) == 2, "mono_mix expects [..., 2, L]");
    const int64_t B = xf.size(0), L = xf.size(2);
    Tensor out = at::empty({B, L}, f32(xf));
    if (out.numel()) check(acids_mono_mix(xf.data_ptr<float>(), B, L, out.data_ptr<float>(), stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {L})), x);
}

Tensor midside(const Tensor& x, bool pad_mid, bool inverse) {
    Tensor xd = dev(x).to(at::kFloat);
    c10::cuda::CUDAGuard guard(xd.device());
    std::vector<int64_t> batch;
    Tensor xf = flat(xd, 2, batch);
    TORCH_CHECK(xf.size(1) == 2, "midside expects [..., 2, L]");
    const int64_t B = xf.size(0), L = xf.size(2);
    Tensor out = at::empty({B, 2, L}, f32(xf));
    if (out.numel()) check(acids_midside(xf.data_ptr<float>(), B, L, pad_mid ? 1 : 0, inverse ? 1 : 0, out.data_ptr<float>(), stream_of(xf)));
    return ret(out.reshape(with_batch(batch, {2, L})), x);
}

}  // namespace

TORCH_LIBRARY(acids_b200, m) {
    m.def("stft_fwd(Tensor x, Tensor window, int n_fft, int hop, bool center) -> Tensor", &stft_fwd);
    m.def("midside_stft_fwd(Tensor x, Tensor window, int n_fft, int hop, int midside) -> Tensor", &midside_stft_fwd);
    m.def("stft_mag_fwd(Tensor x, Tensor window, int n_fft, int hop, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, "
          "Tensor? offset, Tensor? scale, bool drop_first) -> Tensor", &stft_mag_fwd);
    m.def("stft_polar_fwd(Tensor x, Tensor window, int n_fft, int hop, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, "
          "Tensor? mag_offset, Tensor? mag_scale, int phase_mode, int method, bool weighted, Tensor? ph_offset, Tensor? ph_scale, "
          "bool drop_first, int midside=0) -> Tensor", &stft_polar_fwd);
    m.def("mag_epilogue(Tensor X, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? offset, Tensor? scale, "
          "bool drop_first) -> Tensor", &mag_epilogue);
    m.def("mag_invert(Tensor y, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? offset, Tensor? scale, "
          "bool pad_last) -> Tensor", &mag_invert);
    m.def("melspec_fwd(Tensor x, Tensor window, int n_fft, int hop, Tensor band_meta, Tensor band_coef, float power, Tensor? offset, "
          "Tensor? scale) -> Tensor", &melspec_fwd);
    m.def("mfcc_dct(Tensor mel, Tensor dct, float top_db) -> Tensor", &mfcc_dct);
    m.def("phase_fwd(Tensor X, int mode, int method, bool weighted, Tensor? offset, Tensor? scale, bool drop_first) -> Tensor", &phase_fwd);
    m.def("phase_inv(Tensor y, int mode, int method, Tensor? offset, Tensor? scale, bool pad_last) -> Tensor", &phase_inv);
    m.def("polar_fwd(Tensor X, Tensor? band_meta, Tensor? band_coef, int contrast, float eps, Tensor? mag_offset, Tensor? mag_scale, "
          "int phase_mode, int method, bool weighted, Tensor? ph_offset, Tensor? ph_scale, bool drop_first) -> Tensor", &polar_fwd);
    m.def("phase_inv_polar(Tensor y, Tensor mag, int mode, int method, Tensor? offset, Tensor? scale, bool pad_last) -> Tensor", &phase_inv_polar);
    m.def("polar_to_complex(Tensor mag, Tensor phase) -> Tensor", &polar_to_complex);
    m.def("griffinlim_update(Tensor rebuilt, Tensor tprev, Tensor mag, float momentum) -> Tensor", &griffinlim_update);
    m.def("istft_ola(Tensor X, Tensor window, int n_fft, int hop) -> Tensor", &istft_ola);
    m.def("irfft_frames(Tensor X, Tensor window, int n_fft) -> Tensor", &irfft_frames);
    m.def("ola_stream(Tensor frames, int hop, int keep, Tensor? carry_in, float gain) -> (Tensor, Tensor)", &ola_stream);
    m.def("mulaw_encode(Tensor x, int channels, int one_hot) -> Tensor", &mulaw_encode);
    m.def("mulaw_decode(Tensor q, int channels) -> Tensor", &mulaw_decode);
    m.def("one_hot(Tensor q, int n_classes) -> Tensor", &one_hot);
    m.def("stats(Tensor x, int contrast, float eps, bool abs_contrast=False) -> Tensor", &stats);
    m.def("stft_stats(Tensor x, Tensor window, int n_fft, int hop, int contrast, float eps) -> Tensor", &stft_stats);
    m.def("mono_mix(Tensor x) -> Tensor", &mono_mix);
    m.def("midside(Tensor x, bool pad_mid, bool inverse) -> Tensor", &midside);
}
