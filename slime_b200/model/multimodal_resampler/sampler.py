"""Drop-in for the reference's llava/model/multimodal_resampler/sampler.py::Resampler (one cross-attention
layer, learned queries, fixed 2-D sincos positions).  Parameters keep the reference names (query, pos_embed,
attn.in_proj_weight/bias, attn.out_proj.*, ln_q/ln_kv/ln_post.*); forward = slime_resampler_fwd."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...config import SlimeConfig
from ...synth import sincos_2d
from .._runtime import EngineBinding, bind, binding_of


class _MHAParams(nn.Module):
    """Holds nn.MultiheadAttention's parameter names without its (never used) forward."""

    def __init__(self, embed_dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)


class Resampler(nn.Module):
    def __init__(self, grid_size, embed_dim, num_heads, kv_dim=None, llm_hidden_size=4096, norm_layer=None,
                 use_post_proj=False):
        super().__init__()
        if kv_dim is not None and kv_dim != embed_dim:
            raise NotImplementedError("kv_proj != Identity is not used by any SliME configuration")
        if use_post_proj:
            raise NotImplementedError("use_post_proj is not used by any SliME configuration")
        if num_heads != embed_dim // 128:
            raise NotImplementedError("the reference always builds heads of 128 (num_heads = embed_dim // 128)")
        self.num_queries = grid_size ** 2
        self.grid_size = grid_size
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.llm_hidden_size = llm_hidden_size
        self.pos_embed = nn.Parameter(sincos_2d(embed_dim, grid_size).to(torch.float16), requires_grad=False)
        self.query = nn.Parameter(torch.zeros(self.num_queries, embed_dim))
        nn.init.trunc_normal_(self.query, std=.02)
        self.kv_proj = nn.Identity()
        self.attn = _MHAParams(embed_dim)
        self.ln_q = nn.LayerNorm(embed_dim, eps=1e-6)
        self.ln_kv = nn.LayerNorm(embed_dim, eps=1e-6)
        self.ln_post = nn.LayerNorm(embed_dim, eps=1e-6)
        self.proj = nn.Identity()

    @property
    def _which(self):
        return 0 if self.num_queries != 576 else 1

    def _engine(self, device):
        b = binding_of(self)
        if b is None:
            cfg = SlimeConfig(vit_hidden=self.embed_dim, vit_heads=self.embed_dim // 64, hidden_size=self.llm_hidden_size,
                              mm_resampler_dim=self.num_queries if self._which == 0 else 144)
            prefix = "model.sampler.post_qformer." if self._which == 0 else "model.mm_projector.attn."
            b = EngineBinding(self, cfg, prefix, ("rs_local",) if self._which == 0 else ("rs_global",))
            bind(self, b)
        return b.engine(device)

    def forward(self, x, tgt_size=(24, 24), text=None, attn_mask=None):
        mark = x.dim() <= 2
        if mark:
            x = x.unsqueeze(0)
        if x.shape[1] != 576:
            raise NotImplementedError("the B200 path is built for the 24x24 CLIP-L/14-336 patch grid")
        out = self._engine(x.device).resampler(self._which, x).to(x.dtype)
        return out.squeeze(0) if mark else out
