import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from acids_transforms_b200 import transforms as T
from conftest import if_mask
host = lambda t: t.detach().cpu().numpy()
torch.manual_seed(7)
x = torch.randn(5, 2, 20000, device="cuda")
for pad_mid in (True, False):
    for keep in (True, False):
        for mel in (True, False):
            rep = T.PolarIF(magnitude_args={"mode": "bipolar", "n_fft": 1024, "mel": mel}, keep_nyquist=keep)
            ch = (T.MidSide(pad_mid=pad_mid) + T.STFT(n_fft=1024, hop_length=256) + rep).cuda()
            ch.scale_data(x)
            ch._plan[0].one_kernel = True
            y, ref = host(ch(x)), host(ch.forward_unfused(x))
            X = host(ch[1](ch[0](x)))[..., (0 if keep else 1):]
            ok = if_mask(X)
            absX = np.abs(X)
            sc = float(ch[2].phase.norm.scale) * np.pi
            d = np.where(ok, np.abs(y[..., 1, :] - ref[..., 1, :]) * sc, 0)
            i = np.unravel_index(d.argmax(), d.shape)
            print(pad_mid, keep, mel, "mag err %.2e" % np.abs(y[..., 0, :] - ref[..., 0, :]).max(), "ph max %.3e at" % d.max(), i,
                  "|X| %.3e prev %.3e peak %.3e  y %.6f ref %.6f  n(>1e-4)=%d" % (absX[i], absX[i[0], i[1], max(i[2]-1,0), i[3]], absX.max(), y[i[0], i[1], i[2], 1, i[3]], ref[i[0], i[1], i[2], 1, i[3]], (d > 1e-4).sum()))
