"""Shape utilities — host-side mirror of acids_transforms/transforms/misc.py.

Unsqueeze / Squeeze / Transpose are metadata-only (out of scope, kept so chains
compose); OneHot.forward is the int64 one-hot writer that follows MuLaw (kernel 5, bit-exact).
"""
from typing import List, Optional

import torch

from .base import AudioTransform, NotInvertibleError
from .. import _torch_ops  # noqa: F401

__all__ = ["Unsqueeze", "Squeeze", "Transpose", "OneHot"]


class Unsqueeze(AudioTransform):
    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return self.dim is not None

    @property
    def needs_scaling(self):
        return False

    def __repr__(self):
        return "Unsqueeze(dim=%s)" % self.dim

    def __init__(self, sr: int = 44100, dim: int = 1):
        super().__init__(sr)
        self.dim = dim

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x.unsqueeze(self.dim)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        return x.squeeze(self.dim)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        fake = torch.zeros(2, 512)
        assert self(fake).shape == (2, 1, 512)
        return fake if time is None else (fake, time)

    def test_inversion(self, x: torch.Tensor):
        assert self.invert(self.forward(torch.zeros(2, 512))).shape == (2, 512)
        return {}


class Squeeze(AudioTransform):
    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return self.dim is not None

    @property
    def needs_scaling(self):
        return False

    def __init__(self, sr: int = 44100, dim: Optional[int] = None):
        super().__init__(sr)
        self.dim = dim

    def __repr__(self):
        return "Squeeze(dim=%s)" % self.dim

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dim = self.dim
        if dim is None:
            return x.squeeze()
        return x.squeeze(dim)

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        dim = self.dim
        if dim is None:
            raise NotInvertibleError
        return x.unsqueeze(dim)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        fake = torch.zeros(2, 1, 512, 1)
        self.dim = None
        assert self(fake).shape == (2, 512)
        self.dim = 1
        assert self(fake).shape == (2, 512, 1)
        return fake if time is None else (fake, time)

    def test_inversion(self, x: torch.Tensor):
        self.dim = 1
        fake = torch.zeros(2, 1, 512, 1)
        assert self.invert(self.forward(fake)).shape == fake.shape
        return {}


class Transpose(AudioTransform):
    invertible = True

    @property
    def scriptable(self):
        return True

    def __repr__(self):
        return "Transpose(dims=%s, contiguous=%s)" % (self.dims, self.contiguous)

    def __init__(self, dims=(-2, -1), contiguous: bool = True):
        super().__init__()
        self.dims: List[int] = list(dims)
        self.contiguous = bool(contiguous)

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = x.transpose(self.dims[0], self.dims[1])
        return y.contiguous() if self.contiguous else y

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        return self.forward(x)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        y = self(torch.zeros(2, 128, 512))
        assert y.shape == (2, 512, 128)
        return y if time is None else (y, time)

    def test_inversion(self, x: torch.Tensor):
        assert self.invert(self.test_forward(x)).shape == (2, 128, 512)
        return {}


class OneHot(AudioTransform):
    @property
    def scriptable(self):
        return True

    @property
    def invertible(self):
        return True

    @property
    def needs_scaling(self):
        return self.n_classes == -1

    def __init__(self, sr: int = 44100, dtype=torch.long, n_classes: int = -1):
        super().__init__(sr)
        self.n_classes = n_classes

    def __repr__(self):
        return "OneHot(n_classes=%s)" % self.n_classes

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        self.n_classes = int(x.max()) + 1

    @torch.jit.export
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.acids_b200.one_hot(x, self.n_classes)        # F.one_hot int64, misc.py:176-179

    @torch.jit.export
    def invert(self, x_onehot: torch.Tensor, inversion_mode: Optional[str] = None, tolerance: float = 1.e-4) -> torch.Tensor:
        return x_onehot.argmax(-1)

    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        q = torch.randint(0, 256, (2, 44100))
        self.scale_data(q)
        return self(q) if time is None else self.forward_with_time(q, time)

    def test_inversion(self, x: torch.Tensor):
        self.invert(torch.randint_like(x, 0, 256))
        return {}

    @classmethod
    def test_scripted_transform(cls, transform, invert: bool = True):
        q = torch.randint(0, 256, (2, 44100))
        transform.scale_data(q)
        y = transform(q)
        if invert:
            transform.invert(y)
