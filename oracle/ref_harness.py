"""TEST INFRASTRUCTURE - builds the UNMODIFIED reference model (/root/reference, read-only) from a
SlimeConfig and the synthetic state-dict, for generating golden vectors (oracle/gen_golden.py) and
for the CPU baseline arm of bench.py when the reference tree is present.

Only usable where /root/reference exists (this container); nothing on the GPU box imports it.
Follows the construction gotchas recorded in SURVEY.md 8(c).
"""
from __future__ import annotations

import os
import sys
import tempfile

import torch

REFERENCE_ROOT = os.environ.get("SLIME_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llava"))


def build_reference(cfg, dtype=torch.float32, seed: int = 3407, state_dict=None, attn_implementation="eager"):
    """Returns the reference `LlavaLlamaForCausalLM` (eval mode) holding the synthetic weights."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel

    from llava.model.language_model.llava_llama import LlavaConfig, LlavaLlamaForCausalLM  # type: ignore

    from slime_b200.synth import synth_state_dict

    tmp = tempfile.mkdtemp(prefix="slime_clip_")
    clip_cfg = CLIPVisionConfig(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp,
                                num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                                image_size=cfg.vit_image, patch_size=cfg.vit_patch,
                                layer_norm_eps=cfg.vit_ln_eps, hidden_act="quick_gelu", projection_dim=cfg.vit_hidden)
    CLIPVisionModel(clip_cfg).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": cfg.vit_image}, crop_size={"height": cfg.vit_image, "width": cfg.vit_image}
                       ).save_pretrained(tmp)

    hf = LlavaConfig(
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps,
        rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_position_embeddings, pad_token_id=cfg.pad_token_id,
        head_dim=cfg.head_dim, attention_bias=False, mlp_bias=False, tie_word_embeddings=False,
        attn_implementation=attn_implementation,
    )
    hf.pretraining_tp = 1
    hf.mm_vision_tower = tmp
    hf.mm_vision_select_layer = cfg.mm_vision_select_layer
    hf.mm_vision_select_feature = cfg.mm_vision_select_feature
    hf.mm_projector_type = cfg.mm_projector_type
    hf.mm_hidden_size = cfg.vit_hidden
    hf.mm_resampler_type = cfg.mm_resampler_type
    hf.mm_resampler_dim = cfg.mm_resampler_dim
    hf.mm_resampler_topp = cfg.mm_resampler_topp
    hf.mm_resampler_temp = cfg.mm_resampler_temp
    hf.mm_learnable_gated = cfg.mm_learnable_gated
    hf.mm_patch_merge_type = cfg.mm_patch_merge_type
    hf.image_aspect_ratio = cfg.image_aspect_ratio
    hf.image_grid_pinpoints = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]
    hf.use_local_only = cfg.use_local_only
    hf.use_global_only = cfg.use_global_only
    hf.seperator = cfg.seperator
    hf.tokenizer_padding_side = cfg.tokenizer_padding_side
    hf.tokenizer_model_max_length = cfg.tokenizer_model_max_length
    hf.mm_use_im_start_end = False
    hf.mm_use_im_patch_token = False
    if hasattr(hf, "rope_parameters") and isinstance(hf.rope_parameters, dict):
        hf.rope_parameters["rope_theta"] = cfg.rope_theta

    model = LlavaLlamaForCausalLM(hf)
    model.get_vision_tower().load_model()
    sd = state_dict if state_dict is not None else synth_state_dict(cfg, seed=seed)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys, res
    model = model.to(dtype).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
