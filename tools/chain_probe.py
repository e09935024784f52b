"""Module-level timing of the BASELINE configurations 3 and 4 (not the bench): CUDA events around whole chains."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acids_transforms_b200 import transforms as Tr


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    res = {}
    # cfg 3: MFCC module (reference semantics: mel-spectrogram) and the opt-in 40-coefficient variant
    B3, L3 = 512, 441000
    x3 = 0.5 * (2 * torch.rand((B3, L3), device="cuda") - 1)
    m = Tr.MFCC(n_fft=2048, hop_length=512, n_mels=128).cuda()
    ms = timeit(lambda: m(x3))
    res["cfg3_MFCC_module_128mel"] = dict(ms=ms, audio_s_per_s=B3 * 10 / (ms / 1e3))
    try:
        m40 = Tr.MFCC(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40).cuda()
        ms = timeit(lambda: m40(x3))
        res["cfg3_MFCC_module_40coef"] = dict(ms=ms, audio_s_per_s=B3 * 10 / (ms / 1e3))
    except TypeError as exc:
        res["cfg3_MFCC_module_40coef"] = dict(error=str(exc))
    del x3
    torch.cuda.empty_cache()
    # cfg 4: MidSide + STFT(4096, 1024) + PolarIF, forward and inverse, stereo 4 s clips
    B4, L4 = 256, 176400
    x4 = 0.5 * (2 * torch.rand((B4, 2, L4), device="cuda") - 1)
    ch = (Tr.MidSide() + Tr.STFT(n_fft=4096, hop_length=1024) + Tr.PolarIF(
        magnitude_args={"mode": "bipolar", "n_fft": 4096}, phase_args={"mode": "bipolar"})).cuda()
    ch.scale_data(x4[:8])
    y = ch(x4)
    ms = timeit(lambda: ch(x4))
    res["cfg4_forward_chain"] = dict(ms=ms, audio_s_per_s=B4 * 4 / (ms / 1e3), out_shape=list(y.shape))
    ms = timeit(lambda: ch.invert(y))
    res["cfg4_inverse_chain"] = dict(ms=ms, audio_s_per_s=B4 * 4 / (ms / 1e3))
    del x4, y
    torch.cuda.empty_cache()
    # Griffin-Lim (the STFT default inversion mode): 30 iterations of ISTFT + STFT + one update kernel
    Bg = 256
    xg = 0.5 * (2 * torch.rand((Bg, 176400), device="cuda") - 1)
    st = Tr.STFT(n_fft=1024, hop_length=256).cuda()
    mag = st(xg).abs()
    ms = timeit(lambda: st.invert(mag), iters=3, warm=1)
    res["griffin_lim_256x4s"] = dict(ms=ms, audio_s_per_s=Bg * 4 / (ms / 1e3))
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()
