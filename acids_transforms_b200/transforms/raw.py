"""Raw-domain transforms — host-side mirror of acids_transforms/transforms/raw.py.

Mono (mix) and MidSide are elementwise prologues of the spectral chains and MuLaw is kernel (5): each is
one launch.  Stereo and Window are views / concatenations for which torch is already optimal
(out of scope); they are provided as plain torch so chains keep working.
"""
import math
from typing import Dict, List, Optional

import torch

from .base import AudioTransform, frame_times
from ..utils.misc import frame
from .. import _torch_ops  # noqa: F401

__all__ = ["Mono", "Stereo", "MidSide", "Window", "MuLaw"]


class _Raw(AudioTransform):
    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return True

    @property
    def needs_scaling(self):
        return False


class Mono(_Raw):
    def __init__(self, mode: str = "mix", normalize: bool = False, squeeze: bool = True, inversion_mode: str = "mono"):
        super().__init__()
        self.mode = mode
        self.squeeze = squeeze
        self.normalize = normalize
        self.inversion_mode = inversion_mode

    def __repr__(self):
        return "Mono(mode=%s, normalize=%s squeeze=%s, inversion_mode=%s)" % (self.mode, self.normalize, self.squeeze,
                                                                              self.inversion_mode)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """raw.py:34-49."""
        if x.size(-2) == 2:
            if self.mode == "mix":
                x = torch.ops.acids_b200.mono_mix(x).unsqueeze(-2)
            elif self.mode == "right":
                x = x[..., 1:2, :]
            elif self.mode == "left":
                x = x[..., 0:1, :]
        if self.normalize:
            x = x / x.max()
        if self.squeeze:
            x = x.squeeze(-2)
        return x

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        time = time[..., 0] if self.squeeze else time[..., 0].unsqueeze(-1)
        return self.forward(x), time

    def get_inversion_modes(self):
        return ["mono", "stereo"]

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 0.0) -> torch.Tensor:
        if self.squeeze:
            x = x.unsqueeze(-2)
        if x.size(-2) == 1 and self.inversion_mode == "stereo":
            x = torch.cat([x, x], dim=-2)
        return x

    def test_inversion(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        y = self.forward(x)
        return {m: self.invert(y, inversion_mode=m) for m in self.get_inversion_modes()}


class Stereo(_Raw):
    def __init__(self, normalize: bool = False, sr: int = 44100):
        super().__init__()
        self.normalize = normalize

    def __repr__(self):
        return "Stereo(normalize=%s)" % self.normalize

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim == 1:
            x = torch.stack([x, x], dim=0)
        elif x.size(-2) == 1:
            x = torch.cat([x, x], dim=-2)
        elif x.size(-2) > 2:
            raise Exception("Stereo only works with 1/2 channels")
        if self.normalize:
            x = x / x.max()
        return x

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        if x.ndim == 1:
            return torch.stack([x, x], dim=0)
        if x.size(-2) == 1:
            return torch.cat([x, x], dim=-2)
        if x.size(-2) > 2:
            return x[..., :2, :]
        return x


class MidSide(_Raw):
    def __init__(self, sr: int = 44100, normalize: bool = False, pad_mid: bool = True):
        super().__init__(sr=sr)
        self.pad_mid = pad_mid
        self.normalize = normalize

    def __repr__(self):
        return "MidSide(normalize=%s)" % self.normalize

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """M = (L + R) / 2 (/ sqrt 2 when pad_mid), S = (L - R) / 2  (raw.py:145-162)."""
        if x.ndim == 1:
            x = torch.stack([x, torch.zeros_like(x)], dim=0)
        elif x.size(-2) == 1:
            x = torch.cat([x, torch.zeros_like(x)], dim=-2)
        elif x.size(-2) > 2:
            raise Exception("MidSide only works with 1 or 2 channels")
        else:
            x = torch.ops.acids_b200.midside(x, self.pad_mid, False)
        if self.normalize:
            x = x / x.max()
        return x

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        """L = M sqrt 2 + S, R = M sqrt 2 - S  (raw.py:164-180)."""
        if x.ndim == 1:
            return torch.stack([x, x], dim=0)
        if x.size(-2) == 1:
            return torch.cat([x, x], dim=-2)
        return torch.ops.acids_b200.midside(x[..., :2, :], self.pad_mid, True)


class Window(_Raw):
    def __init__(self, sr: int = 44100, window_size: int = 1024, hop_size: int = 256, dim: int = -1, batch_dim: int = 0,
                 inversion_mode: str = "crop"):
        super().__init__()
        self.sr = sr
        self.window_size = window_size
        self.hop_size = hop_size or window_size
        assert self.window_size >= self.hop_size
        self.dim = dim
        self.batch_dim = batch_dim
        self.inversion_mode = inversion_mode

    def __repr__(self):
        return "Window(ws=%s, hs=%s, dim=%s, inversion=%s)" % (self.window_size, self.hop_size, self.dim, self.inversion_mode)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return frame(x, self.window_size, self.hop_size, self.dim)

    @property
    def ratio(self):
        return self.hop_size

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        y = self.forward(x)
        return y, frame_times(y.size(-2), self.hop_size, self.sr, time)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        """raw.py:236-262: concatenate the first hop of every chunk plus the tail of the last one ("crop")."""
        dim = self.dim if self.dim >= 0 else x.ndim + self.dim
        if self.window_size == self.hop_size:
            shape = list(x.shape)
            return x.reshape(shape[:dim - 1] + [shape[dim - 1] * shape[dim]] + shape[dim + 1:])
        head = x.narrow(dim, 0, self.hop_size)
        head = head.reshape(list(head.shape[:dim - 1]) + [head.size(dim - 1) * self.hop_size] + list(head.shape[dim + 1:]))
        tail = x.select(dim - 1, x.size(dim - 1) - 1).narrow(dim - 1, self.hop_size, x.size(dim) - self.hop_size)
        return torch.cat([head, tail], dim - 1)


class MuLaw(_Raw):
    def __init__(self, channels: int = 256, one_hot: str = "none", **kwargs):
        super().__init__()
        self.channels = channels
        self.one_hot = one_hot

    def __repr__(self):
        return "MuLaw(channels=%s, one_hot=%s)" % (self.channels, self.one_hot)

    def _one_hot_id(self) -> int:
        if self.one_hot == "categorical":
            return 1
        if self.one_hot == "channel":
            return 2
        return 0

    @torch.jit.export
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """mu-law quantisation (+ one-hot) in one kernel: float32 -> int64 (raw.py:280-292)."""
        return torch.ops.acids_b200.mulaw_encode(x, self.channels, self._one_hot_id())

    @torch.jit.export
    def decode(self, x: torch.Tensor) -> torch.Tensor:
        """Undo the one-hot layout (argmax over the class axis), then expand (raw.py:294-308; the reference's
        'categorical' branch indexes the wrong column of nonzero() — argmax is what it means)."""
        x = x.long()
        if self.one_hot == "channel":
            x = x.argmax(-2)
        elif self.one_hot == "categorical":
            x = x.argmax(-1)
        return torch.ops.acids_b200.mulaw_decode(x, self.channels)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.encode(x)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        return torch.ops.acids_b200.mulaw_decode(x, self.channels)       # raw.py:314-316: decoding only
