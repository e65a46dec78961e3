"""Binding between the nn.Module shims (which hold the parameters under the reference's state-dict keys)
and the SlimeEngine that executes them.  The root model owns ONE binding shared by all of its
sub-modules; a module built stand-alone through build_vision_tower / build_vision_projector /
build_vision_sampler lazily gets a private binding that registers only its own weight group."""
from __future__ import annotations

import weakref
from typing import Optional, Sequence

import torch

from ..config import SlimeConfig
from ..engine import SlimeEngine


class EngineBinding:
    def __init__(self, owner: torch.nn.Module, cfg: SlimeConfig, key_prefix: str, groups: Sequence[str]):
        self._owner = weakref.ref(owner)
        self._cfg = cfg                   # SlimeConfig or a zero-argument callable producing one (lazy)
        self.key_prefix = key_prefix      # prepended to owner.state_dict() keys to obtain reference keys
        self.groups = tuple(groups)
        self._engine: Optional[SlimeEngine] = None
        self._stamp = None

    @property
    def cfg(self) -> SlimeConfig:
        if callable(self._cfg):
            self._cfg = self._cfg()
        return self._cfg

    def _current_stamp(self, owner):
        # cheap change detector: (device, data_ptr, version) of every parameter
        return tuple((p.device, p.data_ptr(), p._version) for p in owner.parameters())

    def mark_dirty(self):
        self._stamp = None

    def engine(self, device: Optional[torch.device] = None) -> SlimeEngine:
        owner = self._owner()
        if owner is None:
            raise RuntimeError("the module that owns this engine binding is gone")
        params = list(owner.parameters())
        dev = device
        if dev is None or dev.type != "cuda":
            dev = next((p.device for p in params if p.device.type == "cuda"), None)
        if dev is None or dev.type != "cuda":
            raise RuntimeError("slime_b200 modules run on a CUDA (B200) device only: move the model with .to('cuda') "
                               "or pass CUDA inputs; there is no CPU/PyTorch fallback path")
        stamp = self._current_stamp(owner)
        # compute dtype follows the parameters, as in the reference: model.half() / torch_dtype=float16
        # (llava/model/builder.py:43) -> the float16 build; anything else (bf16, fp32 masters) -> bfloat16
        big = max(params, key=lambda p: p.numel(), default=None)
        dtype = torch.float16 if big is not None and big.dtype == torch.float16 else torch.bfloat16
        if self._engine is None or self._engine.device != dev or self._engine.dtype != dtype:
            self._engine = SlimeEngine(self.cfg, dev, max_pos=max(self.cfg.max_position_embeddings, 4096),
                                       dtype=dtype)
            self._stamp = None
        if stamp != self._stamp:
            sd = owner.state_dict()
            pre = self.key_prefix
            self._engine.load_weights(lambda name: sd[name[len(pre):]] if pre and name.startswith(pre) else sd[name],
                                      self.groups)
            self._stamp = stamp
        return self._engine


def bind(module: torch.nn.Module, binding: EngineBinding) -> None:
    """Attach a binding to `module` and all of its sub-modules (not registered as a sub-module/buffer)."""
    for m in module.modules():
        object.__setattr__(m, "_slime_binding", binding)


def binding_of(module: torch.nn.Module) -> Optional[EngineBinding]:
    return getattr(module, "_slime_binding", None)
