"""Checkpoint formats of the reference's load_pretrained_model (llava/model/builder.py:93-127): a synthetic checkpoint
is written in each on-disk layout and read back into the drop-in model on the CPU (no forward pass)."""
import json
import os
import tempfile

import pytest
import torch

from slime_b200.config import preset
from slime_b200.synth import synth_state_dict


def _write_config(d, cfg, clip_dir):
    hf = dict(model_type="llava_llama", hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
              num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
              num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, vocab_size=cfg.vocab_size,
              rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, mm_vision_tower=clip_dir,
              mm_vision_select_layer=-2, mm_projector_type="gated", mm_hidden_size=cfg.vit_hidden,
              mm_resampler_type="cosine", mm_resampler_dim=144, mm_resampler_topp=0.95, mm_resampler_temp=1.0,
              mm_patch_merge_type="spatial", image_aspect_ratio="anyres", seperator=cfg.seperator)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(hf, f)


def _write_clip(cfg, sd):
    from safetensors.torch import save_file

    d = tempfile.mkdtemp(prefix="clip_")
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(dict(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                       num_attention_heads=cfg.vit_heads, image_size=336, patch_size=14, layer_norm_eps=1e-5), f)
    pre = "model.vision_tower.vision_tower."
    save_file({k[len(pre):]: v.contiguous() for k, v in sd.items() if k.startswith(pre)},
              os.path.join(d, "model.safetensors"))
    return d


@pytest.mark.parametrize("layout", ["sharded_safetensors", "single_bin", "base_plus_adapters"])
def test_checkpoint_roundtrip(layout):
    from safetensors.torch import save_file

    from slime_b200.checkpoint import load_model

    cfg = preset("tiny")
    sd = {k: v.to(torch.bfloat16) for k, v in synth_state_dict(cfg).items()}
    clip_dir = _write_clip(cfg, sd)
    llm = {k: v for k, v in sd.items() if not k.startswith("model.vision_tower.")}
    d = tempfile.mkdtemp(prefix="ckpt_")
    _write_config(d, cfg, clip_dir)
    base = None
    if layout == "sharded_safetensors":
        keys = sorted(llm)
        half = len(keys) // 2
        shards = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
        wm = {}
        for fn, ks in shards.items():
            save_file({k: llm[k].contiguous() for k in ks}, os.path.join(d, fn))
            wm.update({k: fn for k in ks})
        with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
            json.dump({"weight_map": wm}, f)
    elif layout == "single_bin":
        torch.save(llm, os.path.join(d, "pytorch_model.bin"))
    else:  # LLM weights in a base dir, adapters (projector + sampler) in the model dir
        base = tempfile.mkdtemp(prefix="base_")
        _write_config(base, cfg, clip_dir)
        adapters = {k: v for k, v in llm.items() if "mm_projector" in k or "sampler" in k}
        torch.save({k: v for k, v in llm.items() if k not in adapters}, os.path.join(base, "pytorch_model.bin"))
        torch.save({k: v for k, v in adapters.items() if "mm_projector" in k}, os.path.join(d, "mm_projector.bin"))
        torch.save({k: v for k, v in adapters.items() if "sampler" in k}, os.path.join(d, "sampler.bin"))
    model = load_model(d, base, device="cpu", dtype=torch.bfloat16)
    got = model.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert torch.equal(got[k].to(torch.bfloat16), v), k


def test_missing_tensor_is_reported():
    from slime_b200.checkpoint import load_model

    cfg = preset("tiny")
    sd = {k: v.to(torch.bfloat16) for k, v in synth_state_dict(cfg).items()}
    clip_dir = _write_clip(cfg, sd)
    d = tempfile.mkdtemp(prefix="ckpt_")
    _write_config(d, cfg, clip_dir)
    llm = {k: v for k, v in sd.items() if not k.startswith("model.vision_tower.") and "w_gate" not in k}
    torch.save(llm, os.path.join(d, "pytorch_model.bin"))
    with pytest.raises(RuntimeError, match="lacks"):
        load_model(d, None, device="cpu")
