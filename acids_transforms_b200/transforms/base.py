"""Protocol and `+` composition of the audio transforms — the drop-in boundary of SURVEY.md §8(b).

Mirrors the reference's interface (acids_transforms/transforms/base.py:13-180): the class flags
`invertible` / `scriptable` / `needs_scaling`, `forward` / `invert` / `scale_data` /
`forward_with_time`, `realtime()`, `ratio`, the `test_*` hooks its reflection-driven tests call, and
`ComposeAudioTransform` whose flags are AND/OR over its children.  No arithmetic lives here.
"""
from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn


class NotInvertibleError(Exception):
    pass


InversionEnumType = Union[str, None]


class AudioTransform(nn.Module):
    invertible = True
    scriptable = False
    needs_scaling = False

    def __init__(self, sr: int = 44100):
        super().__init__()
        self.sr = sr

    def __repr__(self):
        return "AudioTransform()"

    def __add__(self, other):
        if isinstance(other, ComposeAudioTransform):
            return ComposeAudioTransform(transforms=[self] + list(other.transforms))
        if isinstance(other, AudioTransform):
            return ComposeAudioTransform(transforms=[self, other])
        raise TypeError("AudioTransform cannot be added to type: %s" % type(other))

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        pass

    @torch.jit.export
    def forward(self, x):
        return x

    def get_inversion_modes(self):
        return None

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None) -> torch.Tensor:
        return x

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        return self.forward(x), time

    def realtime(self):
        return self

    @property
    def ratio(self):
        return 1

    # ---- hooks called by the reference's test-suite (base.py:60-80) ----
    def test_forward(self, x: torch.Tensor, time: Optional[torch.Tensor] = None):
        if time is None:
            return self.forward(x)
        return self.forward_with_time(x, time)

    def test_inversion(self, x: torch.Tensor):
        if not self.invertible:
            raise NotImplementedError
        return {"inverted": self.invert(self.forward(x))}

    @classmethod
    def test_scripted_transform(cls, transform, batch_size=(2, 2), invert=True):
        x = torch.zeros(*batch_size, 44100)
        time = torch.zeros(*batch_size)
        transform.forward(x)
        x_t, _ = transform.forward_with_time(x, time)
        if invert:
            transform.invert(x_t)


def frame_times(n_frames: int, hop: int, sr: int, time: torch.Tensor) -> torch.Tensor:
    """Frame time stamps t * hop / sr + time[..., None]  (stft.py:106-117), built on `time`'s device."""
    shifts = torch.arange(n_frames, device=time.device, dtype=torch.float32) * float(hop) / float(sr)
    return shifts + time.unsqueeze(-1)


class ComposeAudioTransform(AudioTransform):
    """Chain built by `a + b + c` (base.py:83-180)."""

    def __init__(self, transforms=[], sr: int = 44100):
        super().__init__(sr=sr)
        self.transforms = nn.ModuleList(list(transforms))
        # Execution plan: the same child modules, with fusable neighbours (STFT|DGT + Magnitude) replaced by a
        # stage that runs them as ONE kernel.  `transforms` stays the public, reference-compatible view
        # (indexing, invert order, state_dict keys); `_plan` shares its modules and is hidden from state_dict.
        from .fused import build_plan
        self._plan = nn.ModuleList(build_plan(self.transforms))
        self._register_state_dict_hook(_drop_plan_keys)
        self._register_load_state_dict_pre_hook(_alias_plan_keys, with_module=True)

    @property
    def invertible(self):
        return all(t.invertible for t in self.transforms)

    @property
    def needs_scaling(self):
        return any(t.needs_scaling for t in self.transforms)

    @property
    def scriptable(self):
        return all(t.scriptable for t in self.transforms)

    def __getitem__(self, item):
        return self.transforms[item]

    def __len__(self):
        return len(self.transforms)

    def __repr__(self) -> str:
        return "ComposeAudioTransform(%s)" % [repr(t) + "\n" for t in self.transforms]

    def __add__(self, other):
        if not isinstance(other, AudioTransform):
            raise TypeError("ComposeAudioTransform can only be added to other AudioTransforms")
        if isinstance(other, ComposeAudioTransform):
            return ComposeAudioTransform(list(self.transforms) + list(other.transforms))
        return ComposeAudioTransform(list(self.transforms) + [other])

    def __radd__(self, other):
        if not isinstance(other, AudioTransform):
            raise TypeError("ComposeAudioTransform can only be added to other AudioTransforms")
        if isinstance(other, ComposeAudioTransform):
            return ComposeAudioTransform(list(other.transforms) + list(self.transforms))
        return ComposeAudioTransform([other] + list(self.transforms))

    def realtime(self):
        return ComposeAudioTransform(transforms=[t.realtime() for t in self.transforms], sr=self.sr)

    @property
    def ratio(self):
        r = 1
        for t in self.transforms:
            r = r * t.ratio
        return r

    @torch.jit.export
    def scale_data(self, x: torch.Tensor) -> None:
        # each stage is fitted on the output of the previous ones (base.py:144-148)
        for t in self._plan:
            t.scale_data(x)
            x = t(x)

    @torch.jit.export
    def forward(self, x: torch.Tensor):
        for t in self._plan:
            x = t(x)
        return x

    @torch.jit.export
    def forward_with_time(self, x: torch.Tensor, time: torch.Tensor):
        for t in self._plan:
            x, time = t.forward_with_time(x, time)
        return x, time

    def forward_unfused(self, x: torch.Tensor):
        """The children one after the other, exactly like the reference's loop (base.py:150-154)."""
        for t in self.transforms:
            x = t(x)
        return x

    @torch.jit.export
    def invert(self, x: torch.Tensor, inversion_mode: Optional[str] = None):
        # through the plan as well: torch.jit.script copies shared submodules, so a scripted chain must use
        # ONE set of children for scale_data / forward / invert (a fused stage inverts its children in turn)
        for t in self._plan[::-1]:
            x = t.invert(x, inversion_mode=inversion_mode)
        return x

    def get_inversion_modes(self, idx):
        return type(self.transforms[idx]).get_inversion_modes()

    def test_inversion(self, x: torch.Tensor):
        if not self.invertible:
            raise NotImplementedError
        y = self.forward(x)
        return {"inverted": self.invert(y)}


def _drop_plan_keys(module, state_dict, prefix, local_metadata):
    for k in [k for k in state_dict if k.startswith(prefix + "_plan.")]:
        del state_dict[k]
    return state_dict


def _alias_plan_keys(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """The plan shares its modules with `transforms`: give the loader the keys it will look up under `_plan.`."""
    own = module.state_dict(prefix=prefix)                       # reference-compatible keys
    index = {id(v): k for k, v in module.transforms.state_dict(prefix=prefix + "transforms.", keep_vars=True).items()}
    for name, buf in list(module._plan.named_buffers(prefix=prefix + "_plan", remove_duplicate=False)) + \
            list(module._plan.named_parameters(prefix=prefix + "_plan", remove_duplicate=False)):
        src = index.get(id(buf))
        if src is not None and src in state_dict and name not in state_dict:
            state_dict[name] = state_dict[src]


def apply_transform_to_list(transform, data, time=None, **kwargs):
    if time is None:
        return [transform(d, **kwargs) for d in data]
    outs = [transform(d, time=t, **kwargs) for d, t in zip(data, time)]
    return [o[0] for o in outs], [o[1] for o in outs]


def apply_invert_transform_to_list(transform, data, time=None, **kwargs):
    if time is None:
        return [transform.invert(d, **kwargs) for d in data]
    outs = [transform.invert(d, time=t, **kwargs) for d, t in zip(data, time)]
    return [o[0] for o in outs], [o[1] for o in outs]
