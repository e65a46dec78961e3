"""Packs a reference-keyed state dict into the layouts the sm_100a kernels consume and
registers them with a slime_ctx by canonical name (include/slime_b200.h: slime_ctx_set_weight).

Packing = one-time, load-time layout work done with torch tensor ops (plumbing):
  * CLIP q/k/v -> one [3D, D] weight + [3D] bias;  Llama q/k/v -> one [(h+2kv)*hd, H] weight
  * Llama gate/up -> rows interleaved (g0,u0,g1,u1,...) so the SwiGLU epilogue sees both halves of
    a pair in one accumulator tile
  * Llama q/k (rope_interleaved): inside every head, rows (i, i + hd/2) made adjacent so the RoPE pair of
    HF's rotate_half (llama/modeling_llama.py:152-176) sits in adjacent accumulator columns and the rotation
    runs in the QKV GEMM's epilogue (SLIME_FLAG_ROPE_INTERLEAVED); q.k is invariant under the common permutation
  * Llama RMSNorm weights (norm_folded): gamma of input_layernorm / post_attention_layernorm multiplied into the COLUMNS
    of the qkv / gate-up weights (in fp32, rounded once); the prefill GEMMs then run on the un-normalised residual stream
    and scale their output rows by 1/rms (SLIME_FLAG_NORM_FOLDED, csrc/gemm.h), the registered norm weights become ones
    (the decode step's fused residual + RMSNorm multiplies by them)
  * CLIP patch conv [D,3,14,14] -> [D, 588] zero-padded to K = 640 (TMA needs 16-byte row strides)
  * Resampler key positions: the bicubic 12x12 -> 24x24 resize of the constant sincos table
    (reference multimodal_resampler/sampler.py:27-36,149-155) is input independent -> done here once
  * buffers for the input-independent projected queries / key-position table, filled by the
    library itself in slime_ctx_finalize_weights
Only the ViT layers that feed hidden_states[mm_vision_select_layer] are packed (23 of 24).
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import torch
import torch.nn.functional as F

from .config import SlimeConfig
from .synth import CLIP_PREFIX


def resize_pos_table(pos: torch.Tensor, tgt: int) -> torch.Tensor:
    """get_abs_pos of the reference (sampler.py:27-36): bicubic, align_corners=False, computed in fp32."""
    src = int(math.sqrt(pos.shape[0]))
    if src == tgt:
        return pos
    out = F.interpolate(pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2), size=(tgt, tgt), mode="bicubic",
                        align_corners=False)
    return out.permute(0, 2, 3, 1).flatten(0, 2).to(pos.dtype)


ALL_GROUPS = ("vit", "rs_local", "rs_global", "proj", "llm", "router")  # "router": only packed for mm_resampler_type="qformer"


def rope_interleave_rows(w: torch.Tensor, head_dim: int) -> torch.Tensor:
    """[heads*hd, H] -> same shape with rows (i, i + hd/2) of every head adjacent: new row 2i = old i, 2i+1 = old i + hd/2."""
    heads = w.shape[0] // head_dim
    return w.reshape(heads, 2, head_dim // 2, w.shape[1]).transpose(1, 2).reshape(w.shape)


def pack_weights(cfg: SlimeConfig, get: Callable[[str], torch.Tensor], device,
                 groups=ALL_GROUPS, dtype: torch.dtype = torch.bfloat16,
                 rope_interleaved: bool = False, norm_folded: bool = False) -> Dict[str, torch.Tensor]:
    """get(name) returns the reference tensor `name` (any dtype/device); returns canonical-name ->
    contiguous CUDA tensor (2-D) in the engine's 16-bit element type.  Tensors are pulled one at a time so a lazy source (e.g. the
    on-GPU synthetic generator) never holds two copies of the model."""
    bf = dtype

    def g(name: str) -> torch.Tensor:
        return get(name).to(device=device, dtype=bf)

    out: Dict[str, torch.Tensor] = {}
    D, H = cfg.vit_hidden, cfg.hidden_size
    v = CLIP_PREFIX
    if "vit" in groups:
        _pack_vit(cfg, g, out)
    _pack_adapter(cfg, g, out, device, groups, bf)
    if "llm" in groups:
        _pack_llm(cfg, g, out, rope_interleaved, norm_folded)
    if "router" in groups and cfg.mm_resampler_type == "qformer":
        _pack_router(cfg, g, out)
    for k, t in out.items():
        assert t.dim() == 2 and t.is_contiguous() and t.dtype == bf, k
    return out


def _pack_vit(cfg, g, out):
    D = cfg.vit_hidden
    v = CLIP_PREFIX
    pw = g(v + "embeddings.patch_embedding.weight").reshape(D, -1)
    out["vit.patch_w"] = F.pad(pw, (0, cfg.vit_kpad - pw.shape[1])).contiguous()
    out["vit.cls"] = g(v + "embeddings.class_embedding").reshape(1, D).contiguous()
    out["vit.pos"] = g(v + "embeddings.position_embedding.weight").contiguous()
    out["vit.pre_ln_w"] = g(v + "pre_layrnorm.weight").reshape(1, D)
    out["vit.pre_ln_b"] = g(v + "pre_layrnorm.bias").reshape(1, D)
    for l in range(cfg.vit_layers_used):
        p, c = f"{v}encoder.layers.{l}.", f"vit.layers.{l}."
        out[c + "qkv_w"] = torch.cat([g(p + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")]).contiguous()
        out[c + "qkv_b"] = torch.cat([g(p + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")]).reshape(1, -1)
        out[c + "o_w"] = g(p + "self_attn.out_proj.weight").contiguous()
        out[c + "o_b"] = g(p + "self_attn.out_proj.bias").reshape(1, -1)
        out[c + "ln1_w"] = g(p + "layer_norm1.weight").reshape(1, -1)
        out[c + "ln1_b"] = g(p + "layer_norm1.bias").reshape(1, -1)
        out[c + "ln2_w"] = g(p + "layer_norm2.weight").reshape(1, -1)
        out[c + "ln2_b"] = g(p + "layer_norm2.bias").reshape(1, -1)
        out[c + "fc1_w"] = g(p + "mlp.fc1.weight").contiguous()
        out[c + "fc1_b"] = g(p + "mlp.fc1.bias").reshape(1, -1)
        out[c + "fc2_w"] = g(p + "mlp.fc2.weight").contiguous()
        out[c + "fc2_b"] = g(p + "mlp.fc2.bias").reshape(1, -1)


def _pack_adapter(cfg, g, out, device, groups, bf):
    D = cfg.vit_hidden
    side = int(math.sqrt(cfg.vit_patches))
    for ref_prefix, c, nq in (("model.sampler.post_qformer.", "rs_local.", cfg.mm_resampler_dim),
                              ("model.mm_projector.attn.", "rs_global.", 576)):
        if c[:-1] not in groups:
            continue
        pos = g(ref_prefix + "pos_embed")
        out[c + "query"] = g(ref_prefix + "query").contiguous()
        out[c + "pos_q"] = pos.contiguous()
        out[c + "pos_k"] = resize_pos_table(pos, side).contiguous()
        out[c + "in_proj_w"] = g(ref_prefix + "attn.in_proj_weight").contiguous()
        out[c + "in_proj_b"] = g(ref_prefix + "attn.in_proj_bias").reshape(1, -1)
        out[c + "out_w"] = g(ref_prefix + "attn.out_proj.weight").contiguous()
        out[c + "out_b"] = g(ref_prefix + "attn.out_proj.bias").reshape(1, -1)
        for ln in ("ln_q", "ln_kv", "ln_post"):
            out[c + ln + "_w"] = g(ref_prefix + ln + ".weight").reshape(1, -1)
            out[c + ln + "_b"] = g(ref_prefix + ln + ".bias").reshape(1, -1)
        out[c + "derived_q"] = torch.zeros(nq, D, device=device, dtype=bf)
        out[c + "derived_kvbias"] = torch.zeros(cfg.vit_patches, 2 * D, device=device, dtype=bf)
    if "proj" not in groups:
        return
    m = "model.mm_projector."
    out["proj.fc1_w"] = g(m + "projection.0.weight").contiguous()
    out["proj.fc1_b"] = g(m + "projection.0.bias").reshape(1, -1)
    out["proj.fc2_w"] = g(m + "projection.2.weight").contiguous()
    out["proj.fc2_b"] = g(m + "projection.2.bias").reshape(1, -1)
    out["proj.w_gate"] = g(m + "w_gate").contiguous()


def _pack_router(cfg, g, out):
    """TextGuidedRouterAttention (reference multimodal_resampler/builder.py:101-137); `query` and `self_attn` are
    parameters the reference creates but never uses in forward - they are not sent to the device."""
    r = "model.sampler.selector."
    out["router.in_proj_w"] = g(r + "cross_attn.in_proj_weight").contiguous()
    out["router.in_proj_b"] = g(r + "cross_attn.in_proj_bias").reshape(1, -1)
    out["router.out_w"] = g(r + "cross_attn.out_proj.weight").contiguous()
    out["router.out_b"] = g(r + "cross_attn.out_proj.bias").reshape(1, -1)
    for ln in ("ln_q", "ln_kv", "ln_post"):
        out[f"router.{ln}_w"] = g(r + ln + ".weight").reshape(1, -1)
        out[f"router.{ln}_b"] = g(r + ln + ".bias").reshape(1, -1)
    out["router.fc1_w"] = g(r + "prob_proj.0.weight").contiguous()
    out["router.fc1_b"] = g(r + "prob_proj.0.bias").reshape(1, -1)
    out["router.fc2_w"] = g(r + "prob_proj.2.weight").reshape(1, -1).contiguous()
    out["router.fc2_b"] = g(r + "prob_proj.2.bias").reshape(1, 1)


def _fold_gamma(w: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """W[:, k] * gamma[k] in fp32, rounded once to W's dtype (norm folding)."""
    return (w.float() * gamma.float().reshape(1, -1)).to(w.dtype).contiguous()


def _pack_llm(cfg, g, out, rope_interleaved=False, norm_folded=False):
    H = cfg.hidden_size
    ones = None
    perm = (lambda w: rope_interleave_rows(w, cfg.head_dim)) if rope_interleaved else (lambda w: w)
    out["llm.embed"] = g("model.embed_tokens.weight").contiguous()
    out["llm.norm_w"] = g("model.norm.weight").reshape(1, -1)
    out["llm.lm_head"] = g("lm_head.weight").contiguous()
    for l in range(cfg.num_hidden_layers):
        p, c = f"model.layers.{l}.", f"llm.layers.{l}."
        out[c + "qkv_w"] = torch.cat([perm(g(p + "self_attn.q_proj.weight")), perm(g(p + "self_attn.k_proj.weight")),
                                      g(p + "self_attn.v_proj.weight")]).contiguous()
        out[c + "o_w"] = g(p + "self_attn.o_proj.weight").contiguous()
        gate, up = g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")
        out[c + "gate_up_w"] = torch.stack([gate, up], dim=1).reshape(2 * cfg.intermediate_size, H).contiguous()
        del gate, up
        out[c + "down_w"] = g(p + "mlp.down_proj.weight").contiguous()
        out[c + "in_norm_w"] = g(p + "input_layernorm.weight").reshape(1, -1)
        out[c + "post_norm_w"] = g(p + "post_attention_layernorm.weight").reshape(1, -1)
        if norm_folded:
            out[c + "qkv_w"] = _fold_gamma(out[c + "qkv_w"], out[c + "in_norm_w"])
            out[c + "gate_up_w"] = _fold_gamma(out[c + "gate_up_w"], out[c + "post_norm_w"])
            if ones is None:
                ones = torch.ones_like(out[c + "in_norm_w"])
            out[c + "in_norm_w"] = ones
            out[c + "post_norm_w"] = ones
