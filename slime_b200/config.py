"""Model description of the SliME prefill path.

`SlimeConfig` carries exactly the attributes the reference reads on the hot path:
the CLIP tower dims (HF CLIPVisionConfig), the Llama dims (HF LlamaConfig) and the SliME
hyper-parameters set by the training scripts (reference scripts/llama/llama3_8b_sft.sh:14-48).
`from_hf_config` accepts the reference's own `LlavaConfig` object, so the shims can be built from a
released checkpoint's config.json unchanged.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

IGNORE_INDEX = -100        # reference llava/constants.py:8
IMAGE_TOKEN_INDEX = -200   # reference llava/constants.py:9


@dataclass
class SlimeConfig:
    # --- CLIP ViT (openai/clip-vit-large-patch14-336) ---
    vit_hidden: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    vit_image: int = 336
    vit_patch: int = 14
    vit_ln_eps: float = 1e-5
    mm_vision_select_layer: int = -2
    mm_vision_select_feature: str = "patch"
    # --- SliME adapter ---
    mm_projector_type: str = "gated"
    mm_resampler_type: str = "cosine"
    mm_resampler_dim: int = 144
    mm_resampler_topp: float = 0.95
    mm_resampler_temp: float = 1.0
    mm_learnable_gated: int = -1
    mm_patch_merge_type: str = "spatial"
    image_aspect_ratio: str = "anyres"
    use_local_only: bool = False
    use_global_only: bool = False
    seperator: int = 1919          # (sic) reference spelling, llava_arch.py:219
    tokenizer_padding_side: str = "right"
    tokenizer_model_max_length: int | None = None
    pad_token_id: int = 0
    # --- Llama / Vicuna decoder ---
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 32
    head_dim: int = 128
    intermediate_size: int = 11008
    vocab_size: int = 32000
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-5
    max_position_embeddings: int = 4096
    name: str = "custom"

    # ---- derived ----
    @property
    def vit_layers_used(self) -> int:
        """hidden_states[select_layer] is the output of this many encoder layers
        (hidden_states has num_layers + 1 entries; reference clip_encoder.py:37)."""
        sl = self.mm_vision_select_layer
        n = sl if sl >= 0 else self.vit_layers + 1 + sl
        if not 0 <= n <= self.vit_layers:
            raise ValueError(f"mm_vision_select_layer={sl} out of range")
        return n

    @property
    def vit_patches(self) -> int:
        return (self.vit_image // self.vit_patch) ** 2

    @property
    def vit_tokens(self) -> int:
        return self.vit_patches + 1

    @property
    def vit_kpad(self) -> int:
        k = 3 * self.vit_patch * self.vit_patch
        return (k + 63) // 64 * 64

    @property
    def resampler_grid(self) -> int:
        return int(round(self.mm_resampler_dim ** 0.5))

    @property
    def qkv_dim(self) -> int:
        return (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim

    def validate(self) -> None:
        if self.mm_projector_type != "gated":
            raise NotImplementedError("only mm_projector_type='gated' (the SliME release setting) is built")
        if self.mm_resampler_type not in ("cosine", "qformer"):
            raise NotImplementedError("mm_resampler_type must be 'cosine' (the SliME release setting) or 'qformer' "
                                      "(the cross-attention router, reference multimodal_resampler/builder.py:94-162)")
        if self.mm_resampler_type == "qformer" and self.hidden_size % 128:
            raise ValueError("the 'qformer' router needs hidden_size % 128 == 0 (heads = hidden_size // 128)")
        if self.mm_vision_select_feature != "patch":
            raise NotImplementedError("only mm_vision_select_feature='patch' is built")
        if self.vit_hidden % 128:
            raise ValueError("mm_hidden_size must be a multiple of 128 (Resampler heads = D // 128)")
        if self.mm_patch_merge_type not in ("flat", "spatial"):
            raise NotImplementedError(f"mm_patch_merge_type={self.mm_patch_merge_type!r}")

    def replace(self, **kw) -> "SlimeConfig":
        return dataclasses.replace(self, **kw)

    @classmethod
    def from_hf_config(cls, cfg, clip_cfg=None) -> "SlimeConfig":
        def g(k, d=None):
            v = getattr(cfg, k, d)
            return d if v is None else v

        heads = g("num_attention_heads", 32)
        hidden = g("hidden_size", 4096)
        out = cls(
            hidden_size=hidden, num_hidden_layers=g("num_hidden_layers", 32),
            num_attention_heads=heads,
            num_key_value_heads=g("num_key_value_heads", heads),
            head_dim=g("head_dim", None) or hidden // heads,
            intermediate_size=g("intermediate_size", 11008), vocab_size=g("vocab_size", 32000),
            rope_theta=float(_rope_theta(cfg)), rms_norm_eps=g("rms_norm_eps", 1e-5),
            max_position_embeddings=g("max_position_embeddings", 4096),
            mm_vision_select_layer=g("mm_vision_select_layer", -2),
            mm_vision_select_feature=g("mm_vision_select_feature", "patch"),
            mm_projector_type=g("mm_projector_type", "gated"), mm_resampler_type=g("mm_resampler_type", "cosine"),
            mm_resampler_dim=g("mm_resampler_dim", 144), mm_resampler_topp=g("mm_resampler_topp", 0.95),
            mm_resampler_temp=g("mm_resampler_temp", 1.0), mm_learnable_gated=g("mm_learnable_gated", -1),
            mm_patch_merge_type=g("mm_patch_merge_type", "flat"), image_aspect_ratio=g("image_aspect_ratio", "anyres"),
            use_local_only=g("use_local_only", False), use_global_only=g("use_global_only", False),
            seperator=g("seperator", 1919), tokenizer_padding_side=g("tokenizer_padding_side", "right"),
            tokenizer_model_max_length=g("tokenizer_model_max_length", None), pad_token_id=g("pad_token_id", 0) or 0,
            name=g("_name_or_path", "custom") or "custom",
        )
        if clip_cfg is not None:
            out = out.replace(vit_hidden=clip_cfg.hidden_size, vit_layers=clip_cfg.num_hidden_layers,
                              vit_heads=clip_cfg.num_attention_heads, vit_mlp=clip_cfg.intermediate_size,
                              vit_image=clip_cfg.image_size, vit_patch=clip_cfg.patch_size,
                              vit_ln_eps=clip_cfg.layer_norm_eps)
        return out


def _rope_theta(cfg):
    if getattr(cfg, "rope_theta", None) is not None:
        return cfg.rope_theta
    rp = getattr(cfg, "rope_parameters", None)
    if isinstance(rp, dict) and "rope_theta" in rp:
        return rp["rope_theta"]
    return 10000.0


_CLIP_L = dict(vit_hidden=1024, vit_layers=24, vit_heads=16, vit_mlp=4096, vit_image=336, vit_patch=14)

PRESETS = {
    # model dimensions hard-coded from the HF hub configs the reference downloads
    "vicuna-7b": SlimeConfig(**_CLIP_L, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                             num_key_value_heads=32, intermediate_size=11008, vocab_size=32000, rope_theta=1e4,
                             max_position_embeddings=4096, name="SliME-Vicuna-7B"),
    "llama3-8b": SlimeConfig(**_CLIP_L, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                             num_key_value_heads=8, intermediate_size=14336, vocab_size=128256, rope_theta=5e5,
                             max_position_embeddings=8192, name="SliME-Llama3-8B"),
    "vicuna-13b": SlimeConfig(**_CLIP_L, hidden_size=5120, num_hidden_layers=40, num_attention_heads=40,
                              num_key_value_heads=40, intermediate_size=13824, vocab_size=32000, rope_theta=1e4,
                              max_position_embeddings=4096, name="SliME-Vicuna-13B"),
    # small shapes for parity tests (the oracle finishes in seconds on CPU)
    "tiny": SlimeConfig(vit_hidden=128, vit_layers=3, vit_heads=2, vit_mlp=256, vit_image=336, vit_patch=14,
                        hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                        head_dim=128, intermediate_size=512, vocab_size=1024, rope_theta=1e4,
                        max_position_embeddings=4096, seperator=19, name="tiny"),
    "small": SlimeConfig(vit_hidden=256, vit_layers=4, vit_heads=4, vit_mlp=512, vit_image=336, vit_patch=14,
                         hidden_size=512, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                         head_dim=128, intermediate_size=1024, vocab_size=2048, rope_theta=5e5,
                         max_position_embeddings=4096, seperator=19, name="small"),
}


def preset(name: str, **overrides) -> SlimeConfig:
    return PRESETS[name].replace(**overrides)
