"""Drop-in for the reference's llava/model/multimodal_resampler/builder.py: TextGuidedSampler
(local compression layer `post_qformer` + text-guided router + top-p selection) and build_vision_sampler.
mm_resampler_type='cosine' (the setting of every SliME release) and 'qformer' (the cross-attention router
TextGuidedRouterAttention, reference builder.py:94-162) are built."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...config import SlimeConfig
from .sampler import Resampler, _MHAParams
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


class TextGuidedRouterAttention(nn.Module):
    """Parameter holder of the reference's cross-attention router (builder.py:94-162): same parameter names and
    shapes (`query` and `self_attn` exist in the reference but are never used in its forward); the arithmetic runs
    inside slime_router_fwd* when the ctx is created with SLIME_FLAG_ROUTER_QFORMER."""

    def __init__(self, grid_size, embed_dim, num_heads, kv_dim=None, temp=1.0, norm_layer=nn.LayerNorm):
        super().__init__()
        if kv_dim is not None and kv_dim != embed_dim:
            raise NotImplementedError("kv_proj != Identity is not reachable from TextGuidedSampler")
        if num_heads != embed_dim // 128:
            raise NotImplementedError("the reference always builds heads of 128 (num_heads = embed_dim // 128)")
        self.num_queries = 1
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.temp = temp
        self.query = nn.Parameter(torch.zeros(self.num_queries, embed_dim))
        nn.init.trunc_normal_(self.query, std=.02)
        self.kv_proj = nn.Identity()
        self.self_attn = _MHAParams(embed_dim)
        self.cross_attn = _MHAParams(embed_dim)
        self.ln_q = norm_layer(embed_dim)
        self.ln_kv = norm_layer(embed_dim)
        self.ln_post = norm_layer(embed_dim)
        self.prob_proj = nn.Sequential(nn.Linear(embed_dim, embed_dim // 4), nn.ReLU(), nn.Linear(embed_dim // 4, 1))


class TextGuidedSampler(nn.Module):
    def __init__(self, projector_type, config):
        super().__init__()
        if projector_type not in ("cosine", "qformer"):
            raise NotImplementedError("text-guided router types built: 'cosine' (the SliME release setting), 'qformer'")
        self.num_queries = config.mm_resampler_dim
        self.topp = config.mm_resampler_topp
        self.temp = config.mm_resampler_temp
        self.grid_size = int(math.sqrt(self.num_queries))
        self.router_type = projector_type
        if projector_type == "cosine":
            self.selector = TextGuidedRouterCosine(pad_token_id=getattr(config, "pad_token_id", 0), temp=self.temp,
                                                   embed_dim=config.hidden_size)
        else:
            self.selector = TextGuidedRouterAttention(grid_size=1, embed_dim=config.hidden_size,
                                                      num_heads=config.hidden_size // 128, temp=self.temp)
        self.post_qformer = Resampler(grid_size=self.grid_size, embed_dim=config.mm_hidden_size,
                                      num_heads=config.mm_hidden_size // 128, kv_dim=config.mm_hidden_size,
                                      llm_hidden_size=config.hidden_size)
        self._config = config

    def _engine(self, device):
        b = binding_of(self)
        if b is None:
            cfg = SlimeConfig.from_hf_config(self._config).replace(vit_hidden=self._config.mm_hidden_size,
                                                                   mm_resampler_type=self.router_type)
            b = EngineBinding(self, cfg, "model.sampler.", ("rs_local", "router"))
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
