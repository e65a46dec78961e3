"""Host-side weight packing (slime_b200/weights.py) and configuration checks that need no GPU: canonical names and
shapes of every weight group, the optional cross-attention router group, and the layout transformations the kernels
rely on (packed QKV, interleaved gate/up rows, RoPE row interleave)."""
import pytest
import torch

from slime_b200.config import preset
from slime_b200.synth import synth_state_dict, weight_specs
from slime_b200.weights import ALL_GROUPS, pack_weights, rope_interleave_rows


def pack(cfg, groups=ALL_GROUPS, **kw):
    sd = synth_state_dict(cfg)
    return sd, pack_weights(cfg, lambda n: sd[n], "cpu", groups, torch.bfloat16, **kw)


def test_cosine_config_packs_no_router_group():
    cfg = preset("tiny")
    sd, w = pack(cfg)
    assert not any(k.startswith("router.") for k in w)
    assert not any("sampler.selector" in n for n, _, _ in weight_specs(cfg))
    H, hd = cfg.hidden_size, cfg.head_dim
    assert w["llm.layers.0.qkv_w"].shape == ((cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * hd, H)
    assert w["llm.layers.0.gate_up_w"].shape == (2 * cfg.intermediate_size, H)
    # gate / up rows interleaved (g0, u0, g1, u1, ...): the SwiGLU epilogue sees a pair in adjacent accumulator columns
    gu = w["llm.layers.0.gate_up_w"]
    assert torch.equal(gu[0::2], sd["model.layers.0.mlp.gate_proj.weight"].to(torch.bfloat16))
    assert torch.equal(gu[1::2], sd["model.layers.0.mlp.up_proj.weight"].to(torch.bfloat16))
    # only the ViT layers that feed hidden_states[-2] are packed
    assert f"vit.layers.{cfg.vit_layers_used - 1}.qkv_w" in w and f"vit.layers.{cfg.vit_layers_used}.qkv_w" not in w
    assert w["vit.patch_w"].shape == (cfg.vit_hidden, cfg.vit_kpad)


def test_qformer_config_packs_the_router_group():
    """mm_resampler_type='qformer' (reference multimodal_resampler/builder.py:94-162): cross_attn, the three LayerNorms
    and prob_proj go to the device; `query` and `self_attn` exist in the reference's state dict but are never used."""
    cfg = preset("tiny", mm_resampler_type="qformer")
    cfg.validate()
    sd, w = pack(cfg)
    H = cfg.hidden_size
    expect = {"router.in_proj_w": (3 * H, H), "router.in_proj_b": (1, 3 * H), "router.out_w": (H, H), "router.out_b": (1, H),
              "router.ln_q_w": (1, H), "router.ln_q_b": (1, H), "router.ln_kv_w": (1, H), "router.ln_kv_b": (1, H),
              "router.ln_post_w": (1, H), "router.ln_post_b": (1, H), "router.fc1_w": (H // 4, H),
              "router.fc1_b": (1, H // 4), "router.fc2_w": (1, H // 4), "router.fc2_b": (1, 1)}
    got = {k: tuple(v.shape) for k, v in w.items() if k.startswith("router.")}
    assert got == expect
    assert torch.equal(w["router.in_proj_w"], sd["model.sampler.selector.cross_attn.in_proj_weight"].to(torch.bfloat16))
    assert "model.sampler.selector.self_attn.in_proj_weight" in sd and "model.sampler.selector.query" in sd
    # a binding that registers only some groups does not pull the router in
    _, w2 = pack(cfg, groups=("rs_local",))
    assert not any(k.startswith("router.") for k in w2)


def test_config_validation():
    preset("tiny", mm_resampler_type="cosine").validate()
    with pytest.raises(NotImplementedError):
        preset("tiny", mm_resampler_type="perceiver").validate()
    with pytest.raises(ValueError):
        preset("tiny", mm_resampler_type="qformer", hidden_size=192).validate()
    with pytest.raises(NotImplementedError):
        preset("tiny", mm_projector_type="mlp2x_gelu").validate()


def test_rope_interleave_is_a_row_permutation_that_pairs_the_rotation_partners():
    hd, heads, H = 128, 3, 64
    w = torch.arange(heads * hd * H, dtype=torch.float32).reshape(heads * hd, H)
    p = rope_interleave_rows(w, hd)
    for h in range(heads):
        for i in range(hd // 2):
            assert torch.equal(p[h * hd + 2 * i], w[h * hd + i])
            assert torch.equal(p[h * hd + 2 * i + 1], w[h * hd + i + hd // 2])
    cfg = preset("tiny")
    _, plain = pack(cfg, groups=("llm",))
    _, inter = pack(cfg, groups=("llm",), rope_interleaved=True)
    q_rows = cfg.num_attention_heads * cfg.head_dim
    assert torch.equal(inter["llm.layers.0.qkv_w"][:q_rows], rope_interleave_rows(plain["llm.layers.0.qkv_w"][:q_rows], cfg.head_dim))
    kv = cfg.num_key_value_heads * cfg.head_dim
    assert torch.equal(inter["llm.layers.0.qkv_w"][q_rows + kv:], plain["llm.layers.0.qkv_w"][q_rows + kv:])  # v untouched
