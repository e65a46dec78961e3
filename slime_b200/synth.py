"""Seeded synthetic weights and inputs for the SliME prefill path (no network: there are no
checkpoints or datasets to load).

`synth_state_dict` produces a state-dict keyed EXACTLY like the reference's `LlavaLlamaForCausalLM`
(the golden-vector generator oracle/gen_golden.py loads it into the real reference
with strict=True, which pins the key names and shapes), with every value representable in bf16 so
the fp32 oracle and the bf16 CUDA path start from identical numbers.  Each tensor is drawn from its
own generator seeded by crc32(name) ^ seed, so the dict can be produced lazily, in any order, and
layer by layer for the 8B/13B shapes.
"""
from __future__ import annotations

import math
import zlib
from typing import Callable, Dict, Iterator, Tuple

import numpy as np
import torch

from .config import IMAGE_TOKEN_INDEX, SlimeConfig

CLIP_PREFIX = "model.vision_tower.vision_tower.vision_model."


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(dim // 2, dtype=np.float32)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(dim: int, grid: int) -> torch.Tensor:
    """The Resampler's fixed 2-D sin/cos table (reference multimodal_resampler/sampler.py:39-88):
    first half of the channels encodes the w coordinate, second half the h coordinate."""
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    g = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([_sincos_1d(dim // 2, g[0]), _sincos_1d(dim // 2, g[1])], axis=1)
    # the reference stores the table as fp16 (sampler.py:115-117) and .to(dtype)s it with the model
    return torch.from_numpy(emb).to(torch.float16).float()


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def weight_specs(cfg: SlimeConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """(reference state-dict key, shape, kind) for every tensor of the model."""
    D, I, H = cfg.vit_hidden, cfg.vit_mlp, cfg.hidden_size
    v = CLIP_PREFIX
    yield v + "embeddings.class_embedding", (D,), "embed"
    yield v + "embeddings.patch_embedding.weight", (D, 3, cfg.vit_patch, cfg.vit_patch), "linear"
    yield v + "embeddings.position_embedding.weight", (cfg.vit_tokens, D), "embed"
    yield v + "pre_layrnorm.weight", (D,), "ln_w"
    yield v + "pre_layrnorm.bias", (D,), "ln_b"
    for l in range(cfg.vit_layers):
        p = f"{v}encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{n}.weight", (D, D), "linear"
            yield p + f"self_attn.{n}.bias", (D,), "bias"
        yield p + "layer_norm1.weight", (D,), "ln_w"
        yield p + "layer_norm1.bias", (D,), "ln_b"
        yield p + "mlp.fc1.weight", (I, D), "linear"
        yield p + "mlp.fc1.bias", (I,), "bias"
        yield p + "mlp.fc2.weight", (D, I), "linear"
        yield p + "mlp.fc2.bias", (D,), "bias"
        yield p + "layer_norm2.weight", (D,), "ln_w"
        yield p + "layer_norm2.bias", (D,), "ln_b"
    yield v + "post_layernorm.weight", (D,), "ln_w"
    yield v + "post_layernorm.bias", (D,), "ln_b"
    for prefix, nq in (("model.mm_projector.attn.", 576), ("model.sampler.post_qformer.", cfg.mm_resampler_dim)):
        yield prefix + "pos_embed", (nq, D), "sincos"
        yield prefix + "query", (nq, D), "embed"
        yield prefix + "attn.in_proj_weight", (3 * D, D), "linear"
        yield prefix + "attn.in_proj_bias", (3 * D,), "bias"
        yield prefix + "attn.out_proj.weight", (D, D), "linear"
        yield prefix + "attn.out_proj.bias", (D,), "bias"
        for ln in ("ln_q", "ln_kv", "ln_post"):
            yield prefix + ln + ".weight", (D,), "ln_w"
            yield prefix + ln + ".bias", (D,), "ln_b"
    if cfg.mm_resampler_type == "qformer":  # TextGuidedRouterAttention (reference multimodal_resampler/builder.py:101-137)
        r = "model.sampler.selector."
        yield r + "query", (1, H), "embed"
        for att in ("self_attn", "cross_attn"):  # self_attn is built by the reference but never used in forward
            yield r + att + ".in_proj_weight", (3 * H, H), "linear"
            yield r + att + ".in_proj_bias", (3 * H,), "bias"
            yield r + att + ".out_proj.weight", (H, H), "linear"
            yield r + att + ".out_proj.bias", (H,), "bias"
        for ln in ("ln_q", "ln_kv", "ln_post"):
            yield r + ln + ".weight", (H,), "ln_w"
            yield r + ln + ".bias", (H,), "ln_b"
        yield r + "prob_proj.0.weight", (H // 4, H), "linear"
        yield r + "prob_proj.0.bias", (H // 4,), "bias"
        yield r + "prob_proj.2.weight", (1, H // 4), "router_out"
        yield r + "prob_proj.2.bias", (1,), "bias"
    m = "model.mm_projector."
    yield m + "w_gate", (D, 2), "gate"
    yield m + "w_noise", (D, 2), "zeros"
    yield m + "mean", (1,), "zeros"
    yield m + "std", (1,), "ones"
    yield m + "projection.0.weight", (H, D), "linear"
    yield m + "projection.0.bias", (H,), "bias"
    yield m + "projection.2.weight", (H, H), "linear"
    yield m + "projection.2.bias", (H,), "bias"
    yield "model.embed_tokens.weight", (cfg.vocab_size, H), "embed"
    qd, kd = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        yield p + "self_attn.q_proj.weight", (qd, H), "linear"
        yield p + "self_attn.k_proj.weight", (kd, H), "linear"
        yield p + "self_attn.v_proj.weight", (kd, H), "linear"
        yield p + "self_attn.o_proj.weight", (H, qd), "linear"
        yield p + "mlp.gate_proj.weight", (cfg.intermediate_size, H), "linear"
        yield p + "mlp.up_proj.weight", (cfg.intermediate_size, H), "linear"
        yield p + "mlp.down_proj.weight", (H, cfg.intermediate_size), "linear"
        yield p + "input_layernorm.weight", (H,), "ln_w"
        yield p + "post_attention_layernorm.weight", (H,), "ln_w"
    yield "model.norm.weight", (H,), "ln_w"
    yield "lm_head.weight", (cfg.vocab_size, H), "linear"


def synth_tensor(name: str, shape: Tuple[int, ...], kind: str, seed: int, device="cpu",
                 dtype=torch.float32) -> torch.Tensor:
    """One synthetic tensor; values are bf16-representable.  Linear weights ~ N(0, 1/fan_in) keep
    activations O(1) through the stack so every stage is exercised with non-trivial numerics."""
    if kind == "zeros":
        return torch.zeros(shape, dtype=dtype, device=device)
    if kind == "ones":
        return torch.ones(shape, dtype=dtype, device=device)
    if kind == "sincos":
        return sincos_2d(shape[1], int(round(math.sqrt(shape[0])))).to(device=device, dtype=dtype)
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        g = torch.Generator(device=device)
        g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    else:
        x = torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32)
    if kind == "linear":
        fan_in = int(np.prod(shape[1:]))
        x = x * (1.0 / math.sqrt(fan_in))
    elif kind == "gate":
        x = x * (2.0 / math.sqrt(shape[0]))
    elif kind == "router_out":  # last layer of the qformer router: logits of O(3) so the inner softmax is not flat
        x = x * (6.0 / math.sqrt(shape[1]))
    elif kind == "bias":
        x = x * 0.1
    elif kind == "ln_w":
        x = 1.0 + 0.1 * x
    elif kind == "ln_b":
        x = 0.1 * x
    elif kind == "embed":
        x = x * 0.5
    else:
        raise ValueError(kind)
    return x.to(torch.bfloat16).to(dtype)


def synth_state_dict(cfg: SlimeConfig, seed: int = 3407, device="cpu", dtype=torch.float32,
                     filter_fn: Callable[[str], bool] | None = None) -> Dict[str, torch.Tensor]:
    """Full reference-keyed state dict.  NOTE: CPU and CUDA generators give different streams; the
    parity tests always generate on CPU (the same values on both sides), the benchmark generates on
    the GPU (8B parameters of CPU randn would take minutes)."""
    out = {}
    for name, shape, kind in weight_specs(cfg):
        if filter_fn is not None and not filter_fn(name):
            continue
        out[name] = synth_tensor(name, shape, kind, seed, device=device, dtype=dtype)
    return out


def synth_inputs(cfg: SlimeConfig, batch: int, n_crops: int, prompt_len: int, seed: int = 3407,
                 image_pos: int = 35, ragged: bool = False, device="cpu"):
    """Synthetic request batch: pixels N(0,1) [B,n,3,S,S] (bf16-representable),
    input_ids uniform in [3, V-1000) (or [3, V) for tiny vocabularies) with one IMAGE_TOKEN_INDEX at
    `image_pos`, attention_mask all ones or right-padded with lengths T - (i mod 17) when ragged."""
    g = _gen(f"inputs/{batch}/{n_crops}/{prompt_len}", seed)
    S = cfg.vit_image
    pixels = _bf16_round(torch.randn((batch, n_crops, 3, S, S), generator=g))
    hi = cfg.vocab_size - 1000 if cfg.vocab_size > 2000 else cfg.vocab_size
    ids = torch.randint(3, hi, (batch, prompt_len), generator=g, dtype=torch.long)
    ipos = min(image_pos, prompt_len - 1)
    ids[:, ipos] = IMAGE_TOKEN_INDEX
    mask = torch.ones((batch, prompt_len), dtype=torch.long)
    if ragged:
        for i in range(batch):
            keep = max(ipos + 1, prompt_len - (i % 17))
            mask[i, keep:] = 0
            ids[i, keep:] = cfg.pad_token_id
    return pixels.to(device), ids.to(device), mask.to(device)


def grid_for_crops(n_local: int) -> Tuple[int, int]:
    """(num_patch_width, num_patch_height) used for the spatial merge of n_local crops in synthetic
    runs: the most square factorisation (2x2 for the 672 px headline shape)."""
    h = int(math.sqrt(n_local))
    while h > 1 and n_local % h:
        h -= 1
    h = max(h, 1)
    return n_local // h, h
