"""Drop-in for the reference's llava/model/multimodal_resampler/builder.py: TextGuidedSampler
(local compression layer `post_qformer` + cosine text-guided router + top-p selection) and
build_vision_sampler.  Only mm_resampler_type='cosine' - the setting of every SliME release - is built."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...config import SlimeConfig
from .sampler import Resampler
from .._runtime import EngineBinding, bind, binding_of


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_resampler_type": "identity"}


class TextGuidedRouterCosine(nn.Module):
    """Marker for the cosine selector; its arithmetic runs inside slime_router_fwd*."""

    def __init__(self, pad_token_id, temp=1.0, embed_dim=4096):
        super().__init__()
        self.pad_token_id = pad_token_id
        self.temp = temp


class TextGuidedSampler(nn.Module):
    def __init__(self, projector_type, config):
        super().__init__()
        if projector_type != "cosine":
            raise NotImplementedError("only the 'cosine' text-guided router (the SliME release setting) is built")
        self.num_queries = config.mm_resampler_dim
        self.topp = config.mm_resampler_topp
        self.temp = config.mm_resampler_temp
        self.grid_size = int(math.sqrt(self.num_queries))
        self.selector = TextGuidedRouterCosine(pad_token_id=getattr(config, "pad_token_id", 0), temp=self.temp,
                                               embed_dim=config.hidden_size)
        self.post_qformer = Resampler(grid_size=self.grid_size, embed_dim=config.mm_hidden_size,
                                      num_heads=config.mm_hidden_size // 128, kv_dim=config.mm_hidden_size,
                                      llm_hidden_size=config.hidden_size)
        self._config = config

    def _engine(self, device):
        b = binding_of(self)
        if b is None:
            cfg = SlimeConfig.from_hf_config(self._config).replace(vit_hidden=self._config.mm_hidden_size)
            b = EngineBinding(self, cfg, "model.sampler.", ("rs_local",))
            bind(self, b)
        return b.engine(device)

    def forward(self, local_f, text_embedding, attn_mask=None):
        """local_f [N, H], text_embedding [T, H], attn_mask [T] -> the kept rows of local_f, in order."""
        if self.training:
            raise NotImplementedError("the B200 path is inference-only (Gumbel noise is training code)")
        if local_f.shape[0] == 0:
            return local_f
        eng = self._engine(local_f.device)
        sel_idx, sel_count, _ = eng.router_embeds(local_f.unsqueeze(0), text_embedding.unsqueeze(0),
                                                  None if attn_mask is None else attn_mask.unsqueeze(0))
        k = int(sel_count[0])  # the reference syncs here too (nonzero / numel, builder.py:266-269)
        return local_f[sel_idx[0, :k].long()]


def build_vision_sampler(config, delay_load=False, **kwargs):
    mm_resampler_type = getattr(config, "mm_resampler_type", None)
    if mm_resampler_type == "identity" or mm_resampler_type is None:
        return IdentityMap()
    return TextGuidedSampler(mm_resampler_type, config)
