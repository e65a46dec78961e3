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


def _lora_adapter(d, sd, targets, r=4, alpha=8.0, seed=0, named=False, rslora=False):
    """A PEFT-style LoRA checkpoint in `d`: adapter_config.json + adapter_model.safetensors with lora_A [r, in] /
    lora_B [out, r] for every `targets` leaf of every decoder layer.  Returns {state-dict key: (A, B)}."""
    from safetensors.torch import save_file

    g = torch.Generator().manual_seed(seed)
    ad, pairs = {}, {}
    for k, w in sd.items():
        if not k.startswith("model.layers.") or not k.endswith(".weight"):
            continue
        leaf = k.split(".")[-2]
        if leaf not in targets:
            continue
        A = (torch.randn(r, w.shape[1], generator=g) * 0.05).to(torch.bfloat16)
        B = (torch.randn(w.shape[0], r, generator=g) * 0.05).to(torch.bfloat16)
        mod = "base_model.model." + k[: -len(".weight")]
        mid = ".default" if named else ""
        ad[f"{mod}.lora_A{mid}.weight"] = A
        ad[f"{mod}.lora_B{mid}.weight"] = B
        pairs[k] = (A, B)
    save_file(ad, os.path.join(d, "adapter_model.safetensors"))
    with open(os.path.join(d, "adapter_config.json"), "w") as f:
        json.dump(dict(peft_type="LORA", r=r, lora_alpha=alpha, target_modules=sorted(targets), use_rslora=rslora,
                       fan_in_fan_out=False), f)
    return pairs


@pytest.mark.parametrize("named,rslora", [(False, False), (True, True)])
def test_lora_checkpoint_is_merged_at_load(named, rslora):
    """reference llava/model/builder.py:52-91: base weights + non_lora_trainables.bin + PeftModel.merge_and_unload().
    The merged weight must equal W + scaling * B @ A (peft tuners/lora/layer.py get_delta_weight; scaling = alpha / r,
    alpha / sqrt(r) with rslora), everything else must be untouched."""
    from slime_b200.checkpoint import load_pretrained_model

    cfg = preset("tiny")
    sd = {k: v.to(torch.bfloat16) for k, v in synth_state_dict(cfg).items()}
    clip_dir = _write_clip(cfg, sd)
    llm = {k: v for k, v in sd.items() if not k.startswith("model.vision_tower.")}
    base = tempfile.mkdtemp(prefix="base_")
    d = tempfile.mkdtemp(prefix="lora_")
    _write_config(base, cfg, clip_dir)
    _write_config(d, cfg, clip_dir)
    adapters = {k: v for k, v in llm.items() if "mm_projector" in k or "sampler" in k}
    torch.save({k: v for k, v in llm.items() if k not in adapters}, os.path.join(base, "pytorch_model.bin"))
    # the trainer saves the non-LoRA trainables from the peft-wrapped model: 'base_model.model.' + key
    torch.save({"base_model.model." + k: v for k, v in adapters.items()}, os.path.join(d, "non_lora_trainables.bin"))
    r, alpha = 4, 8.0
    pairs = _lora_adapter(d, llm, {"q_proj", "v_proj", "down_proj"}, r=r, alpha=alpha, named=named, rslora=rslora)
    assert len(pairs) == 3 * cfg.num_hidden_layers
    tok, model, proc, ctx_len = load_pretrained_model(d, base, "slime-tiny-lora", device="cpu", torch_dtype=torch.bfloat16)
    got = model.state_dict()
    scaling = alpha / r ** 0.5 if rslora else alpha / r
    for k, v in sd.items():
        if k in pairs:
            A, B = pairs[k]
            want = (v.float() + scaling * (B.float() @ A.float())).to(torch.bfloat16)
            assert torch.equal(got[k].to(torch.bfloat16), want), k
            assert not torch.equal(want, v)
        else:
            assert torch.equal(got[k].to(torch.bfloat16), v), k
    assert ctx_len == 2048
    with pytest.raises(ValueError, match="model_base"):
        load_pretrained_model(d, None, "slime-tiny-lora", device="cpu")


def test_vision_tower_weights_must_come_from_somewhere():
    """with mm_vision_tower = a hub id that is not in the local HF cache and no tower tensors in the checkpoint
    the loader must raise instead of returning a model with a default-initialised CLIP; with tower tensors in the
    checkpoint it loads them; with the tower's own files present those win unless unfreeze_mm_vision_tower is set
    (reference clip_encoder.py:25-34, builder.py:160-166)."""
    from slime_b200.checkpoint import load_model

    cfg = preset("tiny")
    sd = {k: v.to(torch.bfloat16) for k, v in synth_state_dict(cfg).items()}
    clip_dir = _write_clip(cfg, sd)
    llm = {k: v for k, v in sd.items() if not k.startswith("model.vision_tower.")}

    def ckpt(tower_name, with_tower, **extra):
        d = tempfile.mkdtemp(prefix="ckpt_")
        _write_config(d, cfg, tower_name)
        if extra:
            with open(os.path.join(d, "config.json")) as f:
                raw = json.load(f)
            raw.update(extra)
            with open(os.path.join(d, "config.json"), "w") as f:
                json.dump(raw, f)
        torch.save(dict(sd) if with_tower else llm, os.path.join(d, "pytorch_model.bin"))
        return d

    # (1) a config-only tower directory (dimensions known, no weights) and no tower tensors in the checkpoint
    cfg_only = tempfile.mkdtemp(prefix="clip_cfg_")
    with open(os.path.join(cfg_only, "config.json"), "w") as f:
        json.dump(dict(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                       num_attention_heads=cfg.vit_heads, image_size=336, patch_size=14, layer_norm_eps=1e-5), f)
    with pytest.raises(RuntimeError, match="randomly initialised CLIP"):
        load_model(ckpt(cfg_only, False), None, device="cpu")
    # (2) the same tower name, but the checkpoint carries model.vision_tower.*: loaded from the checkpoint
    m = load_model(ckpt(cfg_only, True), None, device="cpu")
    assert m.get_vision_tower().weights_source == "checkpoint"
    k = "model.vision_tower.vision_tower.vision_model.encoder.layers.0.mlp.fc1.weight"
    assert torch.equal(m.state_dict()[k].to(torch.bfloat16), sd[k])
    # (3) tower directory with weights + DIFFERENT tower tensors in the checkpoint: the tower's own files win ...
    other = {kk: (v + 1 if kk.startswith("model.vision_tower.") else v) for kk, v in sd.items()}
    d3 = tempfile.mkdtemp(prefix="ckpt_")
    _write_config(d3, cfg, clip_dir)
    torch.save(other, os.path.join(d3, "pytorch_model.bin"))
    m = load_model(d3, None, device="cpu")
    assert m.get_vision_tower().weights_source == "directory"
    assert torch.equal(m.state_dict()[k].to(torch.bfloat16), sd[k])
    # ... unless the tower was fine-tuned with the model (unfreeze_mm_vision_tower): then the checkpoint's copy wins
    with open(os.path.join(d3, "config.json")) as f:
        raw = json.load(f)
    raw["unfreeze_mm_vision_tower"] = True
    with open(os.path.join(d3, "config.json"), "w") as f:
        json.dump(raw, f)
    m = load_model(d3, None, device="cpu")
    assert torch.equal(m.state_dict()[k].to(torch.bfloat16), other[k])
