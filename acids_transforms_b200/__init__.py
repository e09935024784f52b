"""acids_transforms_b200 — the spectral hot path of domkirke/acids_transforms on NVIDIA B200 (sm_100a).

Same nn.Module API as the reference (`from acids_transforms_b200.transforms import *`); every hot op is a
hand-written CUDA kernel behind the C ABI of include/acids_b200.h.  There is no CPU or eager fallback.
"""
from .utils import *        # noqa: F401,F403
from .transforms import *   # noqa: F401,F403

__version__ = "0.1.0"
