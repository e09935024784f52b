"""Shape helpers of the reference's utils (acids_transforms/utils/misc.py:138-178): views and padding only."""
from typing import List, Tuple

import torch


def pad(tensor: torch.Tensor, target_size: int, dim: int) -> torch.Tensor:
    """Zero-pad `dim` up to target_size (utils/misc.py:138-145)."""
    if tensor.size(dim) > target_size:
        return tensor
    shape = list(tensor.shape)
    shape[dim] = target_size - tensor.shape[dim]
    return torch.cat([tensor, torch.zeros(shape, dtype=tensor.dtype, device=tensor.device)], dim=dim)


def frame(tensor: torch.Tensor, wsize: int, hsize: int, dim: int) -> torch.Tensor:
    """Overlapping frames as a strided VIEW: [..., L] -> [..., n, wsize] (utils/misc.py:148-165)."""
    if dim < 0:
        dim = tensor.ndim + dim
    if not tensor.is_contiguous():
        tensor = tensor.contiguous()
    n = (tensor.shape[dim] - wsize) // hsize
    if tensor.shape[dim] >= n * hsize + wsize:
        n += 1
    tensor = pad(tensor, n * hsize + wsize, dim)
    shape = list(tensor.shape)
    shape[dim] = n
    shape.insert(dim + 1, wsize)
    strides = [tensor.stride(i) for i in range(tensor.ndim)]
    strides.insert(dim, hsize * tensor.stride(dim))
    return torch.as_strided(tensor, shape, strides)


def reshape_batches(x: torch.Tensor, dim: int, allow_clone: bool = True) -> Tuple[torch.Tensor, List[int]]:
    """Flatten the leading dims, keep the last -dim (utils/misc.py:168-178)."""
    batch = list(x.shape[:dim])
    event = list(x.shape[dim:])
    if not x.is_contiguous() and not allow_clone:
        raise ValueError("found non contiguous tensor of size : %s" % str(x.shape))
    return x.reshape([-1] + event), batch


def import_data(path: str, sr: int = 44100):
    """WAV loader used by the reference's test fixture (utils/misc.py:29-59).  File I/O is outside the
    hot path; scipy reads the files because torchaudio.load needs torchcodec, absent from this image."""
    import os
    import numpy as np
    from scipy.io import wavfile
    if os.path.isfile(path):
        rate, data = wavfile.read(path)
        if data.dtype.kind == "i":
            data = data.astype(np.float64) / float(2 ** (8 * data.dtype.itemsize - 1))
        x = torch.from_numpy(np.atleast_2d(np.asarray(data, np.float32).T)).contiguous()
        if rate != sr:
            import torchaudio
            x = torchaudio.functional.resample(x, rate, sr)
        return x, os.path.basename(path)
    if os.path.isdir(path):
        data, names = [], []
        for f in sorted(os.listdir(path)):
            try:
                x, n = import_data(os.path.join(path, f), sr)
            except Exception:
                continue
            data.append(x)
            names.append(os.path.splitext(n)[0])
        longest = max(d.shape[1] for d in data)
        stereo = any(d.shape[0] == 2 for d in data)
        out = []
        for d in data:
            if d.shape[0] > 1:
                d = d if stereo else d[:1]
            else:
                d = torch.cat([d, d]) if stereo else d
            out.append(torch.cat([d, torch.zeros(d.shape[0], longest - d.shape[1], dtype=d.dtype)], 1))
        return torch.stack(out), names
    raise FileNotFoundError(path)
