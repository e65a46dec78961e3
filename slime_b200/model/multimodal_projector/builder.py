"""Drop-in for the reference's llava/model/multimodal_projector/builder.py: GatedBlock (the gated
MLP (+) Resampler global projector, also used as the plain projector of compressed local tokens) and
build_vision_projector.  Only mm_projector_type='gated' - the setting of every SliME release - is built."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...config import SlimeConfig
from ..multimodal_resampler.sampler import Resampler
from .._runtime import EngineBinding, bind, binding_of


class GatedBlock(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.target_sequence_length = 576
        grid_size = int(math.sqrt(self.target_sequence_length))
        self.attn = Resampler(grid_size=grid_size, embed_dim=config.mm_hidden_size,
                              num_heads=config.mm_hidden_size // 128, kv_dim=config.mm_hidden_size,
                              llm_hidden_size=config.hidden_size, use_post_proj=False)
        self.projection = nn.Sequential(nn.Linear(config.mm_hidden_size, config.hidden_size), nn.GELU(),
                                        nn.Linear(config.hidden_size, config.hidden_size))
        self.expert_ffn = [self.projection, self.attn]
        self.num_experts = len(self.expert_ffn)
        self.w_gate = nn.Parameter(torch.zeros(config.mm_hidden_size, self.num_experts, dtype=torch.bfloat16))
        self.w_noise = nn.Parameter(torch.zeros(config.mm_hidden_size, self.num_experts, dtype=torch.bfloat16))
        self.register_buffer("mean", torch.tensor([0.0], dtype=torch.bfloat16))
        self.register_buffer("std", torch.tensor([1.0], dtype=torch.bfloat16))
        self.learnable_gated = getattr(config, "mm_learnable_gated", -1)
        self.k = 2
        self._config = config

    def _engine(self, device):
        b = binding_of(self)
        if b is None:
            cfg = SlimeConfig.from_hf_config(self._config).replace(vit_hidden=self._config.mm_hidden_size)
            b = EngineBinding(self, cfg, "model.mm_projector.", ("rs_global", "proj"))
            bind(self, b)
        return b.engine(device)

    def forward(self, x, text_embedding=None, attn_mask=None):
        if self.training:
            raise NotImplementedError("the B200 path is inference-only (noisy gating / load losses are training code)")
        eng = self._engine(x.device)
        T = self.target_sequence_length
        if x.shape[0] != T and (x.dim() < 2 or x.shape[1] != T):
            # reference builder.py:180-181: anything that is not a 576-token global view is just projected
            out = eng.projector(x.reshape(-1, x.shape[-1]))
            return out.view(*x.shape[:-1], out.shape[-1]).to(x.dtype)
        mark = x.dim() <= 2
        xs = x.unsqueeze(0) if mark else x
        out = eng.gated_projector(xs).to(x.dtype)
        return out.squeeze(0) if mark else out


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


def build_vision_projector(config, delay_load=False, **kwargs):
    projector_type = getattr(config, "mm_projector_type", "linear")
    if projector_type == "gated":
        return GatedBlock(config)
    if projector_type == "identity":
        return IdentityMap()
    raise NotImplementedError(
        f"mm_projector_type={projector_type!r}: the B200 path builds the SliME release projector ('gated') only; "
        "linear / mlpNx_gelu / qformer / qformer_text are not used by any shipped SliME script (SURVEY.md section 2)")
