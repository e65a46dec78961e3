"""Checkpoint loading for the drop-in model: the on-disk formats the reference's
llava/model/builder.py::load_pretrained_model consumes (:52-127) - HF (sharded) safetensors / pytorch_model*.bin
with a config.json, the adapter-only files the SliME training stages save (mm_projector.bin, sampler.bin,
non_lora_trainables.bin; reference llava/train/train.py:240-272), and LoRA checkpoints (adapter_config.json +
adapter_model.*, merged into the base weights at load like the reference's PeftModel.merge_and_unload, :52-91).
No network, no `transformers` model classes: tensors are read straight into the reference-keyed state dict and loaded
into slime_b200.model.LlavaLlamaForCausalLM.

Vision tower weights (reference clip_encoder.py:25-34, builder.py:160-166): the reference loads CLIP from
`config.mm_vision_tower` AFTER the checkpoint, so the tower's own pretrained files win over `model.vision_tower.*`
keys in the checkpoint - unless `unfreeze_mm_vision_tower` is set (the tower is then built before from_pretrained and
the checkpoint's fine-tuned tower weights overwrite it).  The same precedence is kept here, the tower directory / hub
id is resolved through the local HF cache when it is not a directory, and loading FAILS when neither source provides
the tower (a model whose CLIP is default-initialised noise must never be returned silently).
"""
from __future__ import annotations

import glob
import json
import math
import os
from typing import Dict, Optional

import torch

ADAPTER_FILES = ("mm_projector.bin", "sampler.bin", "non_lora_trainables.bin")
TOWER_PREFIX = "model.vision_tower."


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def _strip_wrappers(part: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Adapter files are saved from (possibly peft-wrapped) trainers: strip the wrappers like the reference does
    (builder.py:77-80: 'base_model.' first, then one 'model.' level when keys start with 'model.model.')."""
    part = {(k[11:] if k.startswith("base_model.") else k): v for k, v in part.items()}
    if any(k.startswith("model.model.") for k in part):
        part = {(k[6:] if k.startswith("model.") else k): v for k, v in part.items()}
    return part


def read_state_dict(model_path: str) -> Dict[str, torch.Tensor]:
    """All tensors of an HF-style checkpoint directory (index + shards, or a single file), adapter files last."""
    sd: Dict[str, torch.Tensor] = {}
    files = []
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        ip = os.path.join(model_path, index)
        if os.path.exists(ip):
            with open(ip) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
            break
    if not files:
        for single in ("model.safetensors", "pytorch_model.bin"):
            if os.path.exists(os.path.join(model_path, single)):
                files = [single]
                break
    if not files:
        files = [os.path.basename(p) for p in sorted(glob.glob(os.path.join(model_path, "model-*.safetensors")))]
    for fn in files:
        sd.update(_load_file(os.path.join(model_path, fn)))
    for fn in ADAPTER_FILES:
        p = os.path.join(model_path, fn)
        if os.path.exists(p):
            sd.update(_strip_wrappers(_load_file(p)))
    return sd


# ------------------------------------------------------------------------------------------------------------------
# LoRA (reference builder.py:52-91: PeftModel.from_pretrained(model, model_path).merge_and_unload())
# ------------------------------------------------------------------------------------------------------------------
def has_lora_adapter(model_path: str) -> bool:
    return os.path.exists(os.path.join(model_path, "adapter_config.json"))


def merge_lora(sd: Dict[str, torch.Tensor], adapter_dir: str) -> int:
    """Merge a PEFT LoRA adapter into the reference-keyed state dict `sd` in place:
        W <- W + scaling * (B @ A),   scaling = lora_alpha / r   (lora_alpha / sqrt(r) with use_rslora)
    (peft 0.x, tuners/lora/layer.py::Linear.get_delta_weight; `fan_in_fan_out` transposes the delta).  The product is
    formed in fp32 and rounded once to W's dtype.  Adapter keys look like
    `base_model.model.<module path>.lora_A.weight` ([r, in]) / `.lora_B.weight` ([out, r]), optionally with an adapter
    name (`.lora_A.default.weight`).  Returns the number of merged matrices; raises if a target is absent from `sd`."""
    with open(os.path.join(adapter_dir, "adapter_config.json")) as f:
        ac = json.load(f)
    r, alpha = int(ac["r"]), float(ac.get("lora_alpha", ac["r"]))
    scaling = alpha / math.sqrt(r) if ac.get("use_rslora") else alpha / r
    fan_in_fan_out = bool(ac.get("fan_in_fan_out", False))
    rank_pattern, alpha_pattern = ac.get("rank_pattern") or {}, ac.get("alpha_pattern") or {}
    path = next((os.path.join(adapter_dir, fn) for fn in ("adapter_model.safetensors", "adapter_model.bin")
                 if os.path.exists(os.path.join(adapter_dir, fn))), None)
    if path is None:
        raise RuntimeError(f"{adapter_dir} has an adapter_config.json but no adapter_model.safetensors / adapter_model.bin")
    ad = _load_file(path)
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in ad.items():
        for tag in ("lora_A", "lora_B"):
            marker = f".{tag}."
            if marker in k:
                mod = k[: k.index(marker)]
                mod = mod[len("base_model.model."):] if mod.startswith("base_model.model.") else mod
                pairs.setdefault(mod, {})[tag] = v
    merged = 0
    for mod, ab in sorted(pairs.items()):
        if "lora_A" not in ab or "lora_B" not in ab:
            raise RuntimeError(f"LoRA adapter: module {mod} lacks its lora_A or lora_B matrix")
        key = mod + ".weight"
        if key not in sd:
            raise RuntimeError(f"LoRA adapter targets {key}, which the base checkpoint does not contain")
        leaf = mod.rsplit(".", 1)[-1]
        r_m = int(next((v for p, v in rank_pattern.items() if mod.endswith(p) or leaf == p), r))
        a_m = float(next((v for p, v in alpha_pattern.items() if mod.endswith(p) or leaf == p), alpha))
        sc = (a_m / math.sqrt(r_m) if ac.get("use_rslora") else a_m / r_m) if (r_m, a_m) != (r, alpha) else scaling
        delta = (ab["lora_B"].float() @ ab["lora_A"].float()) * sc
        if fan_in_fan_out:
            delta = delta.t()
        w = sd[key]
        if delta.shape != w.shape:
            raise RuntimeError(f"LoRA delta of {key} has shape {tuple(delta.shape)}, weight is {tuple(w.shape)}")
        sd[key] = (w.float() + delta).to(w.dtype)
        merged += 1
    return merged


# ------------------------------------------------------------------------------------------------------------------
def load_model(model_path: str, model_base: Optional[str] = None, device: str = "cuda",
               dtype: torch.dtype = torch.bfloat16, lora: Optional[bool] = None):
    """LlavaLlamaForCausalLM from a checkpoint directory.  With `model_base`, the LLM weights come from the base
    directory and `model_path` supplies config.json + the adapter files (the reference's "projector only" case) and,
    when it holds a LoRA adapter (`lora` None = auto-detect), the low-rank deltas merged into the base weights."""
    from .model import LlavaConfig, LlavaLlamaForCausalLM

    cfg = LlavaConfig.from_pretrained(model_path)
    model = LlavaLlamaForCausalLM(cfg)
    tower = model.get_vision_tower()
    if tower is not None:
        tower.load_model()  # creates the CLIP parameter tree and loads mm_vision_tower's own weights if it can find them
    if model_base is not None:
        sd = read_state_dict(model_base)
        for fn in ADAPTER_FILES:
            p = os.path.join(model_path, fn)
            if os.path.exists(p):
                sd.update(_strip_wrappers(_load_file(p)))
    else:
        sd = read_state_dict(model_path)
    sd = {k: v for k, v in sd.items() if "rotary_emb.inv_freq" not in k and not k.endswith("position_ids")}
    if lora is None:
        lora = has_lora_adapter(model_path)
    if lora:
        if not has_lora_adapter(model_path):
            raise RuntimeError(f"{model_path} was named a LoRA checkpoint but has no adapter_config.json")
        merge_lora(sd, model_path)
    # vision tower precedence (module docstring)
    ckpt_has_tower = any(k.startswith(TOWER_PREFIX) for k in sd)
    tower_src = getattr(tower, "weights_source", None) if tower is not None else None
    if tower is not None and tower_src is not None and not getattr(cfg, "unfreeze_mm_vision_tower", False):
        sd = {k: v for k, v in sd.items() if not k.startswith(TOWER_PREFIX)}  # the tower's own pretrained files win
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not k.startswith(TOWER_PREFIX)]
    if missing:
        raise RuntimeError(f"checkpoint at {model_path} lacks {len(missing)} tensors, e.g. {missing[:5]}")
    if res.unexpected_keys:
        raise RuntimeError(f"checkpoint at {model_path} has unexpected tensors, e.g. {res.unexpected_keys[:5]}")
    if tower is not None:
        missing_tower = [k for k in res.missing_keys if k.startswith(TOWER_PREFIX)]
        if tower_src is None and (not ckpt_has_tower or missing_tower):
            raise RuntimeError(
                f"no weights for the vision tower {tower.vision_tower_name!r}: it is neither a local directory nor in the "
                f"local HF cache, and the checkpoint holds {'only part of' if ckpt_has_tower else 'no'} "
                f"'{TOWER_PREFIX}*' tensors - refusing to return a model with a randomly initialised CLIP")
        if tower_src is None:
            tower.weights_source = "checkpoint"
    model = model.to(device=device, dtype=dtype).eval()
    return model


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda", use_flash_attn=False, **kwargs):
    """Same signature and return tuple as the reference's llava/model/builder.py::load_pretrained_model (:26-173):
    (tokenizer, model, image_processor, context_len).  A name containing 'lora' with a `model_base` merges the LoRA
    adapter at load (:52-91); 8/4-bit bitsandbytes loading is not part of the B200 path."""
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes quantised loading is not part of the B200 path")
    lora = None
    if "lora" in model_name.lower():
        if model_base is None:
            raise ValueError("There is `lora` in model name but no `model_base` is provided (reference builder.py:50-51)")
        lora = True
    # the reference's inference dtype is fp16 (builder.py:43); a caller may override it
    model = load_model(model_path, model_base, device=device, dtype=kwargs.get("torch_dtype", torch.float16), lora=lora)
    tokenizer = None
    try:
        from transformers import AutoTokenizer

        tokenizer = AutoTokenizer.from_pretrained(model_base or model_path, use_fast=False)
    except Exception:
        tokenizer = None
    tower = model.get_vision_tower()
    image_processor = tower.image_processor if tower is not None else None
    context_len = getattr(model.config, "max_sequence_length", 2048)
    return tokenizer, model, image_processor, context_len
