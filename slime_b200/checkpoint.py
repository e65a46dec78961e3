"""Checkpoint loading for the drop-in model: the on-disk formats the reference's
llava/model/builder.py::load_pretrained_model consumes (:93-127) - HF (sharded) safetensors / pytorch_model*.bin
with a config.json, plus the adapter-only files the SliME training stages save (mm_projector.bin, sampler.bin,
non_lora_trainables.bin; reference llava/train/train.py:240-272).  No network, no `transformers` model classes:
tensors are read straight into the reference-keyed state dict and loaded into slime_b200.model.LlavaLlamaForCausalLM.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict, Optional

import torch

ADAPTER_FILES = ("mm_projector.bin", "sampler.bin", "non_lora_trainables.bin")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


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
            part = _load_file(p)
            # adapter files are saved from (possibly peft-wrapped) trainers: strip the wrappers like the reference does
            part = {(k[11:] if k.startswith("base_model.") else k): v for k, v in part.items()}
            part = {(k[6:] if k.startswith("model.model.") else k): v for k, v in part.items()}
            sd.update(part)
    return sd


def load_model(model_path: str, model_base: Optional[str] = None, device: str = "cuda",
               dtype: torch.dtype = torch.bfloat16):
    """LlavaLlamaForCausalLM from a checkpoint directory.  With `model_base`, the LLM weights come from the base
    directory and `model_path` supplies config.json + the adapter files (the reference's "projector only" case)."""
    from .model import LlavaConfig, LlavaLlamaForCausalLM

    cfg = LlavaConfig.from_pretrained(model_path)
    model = LlavaLlamaForCausalLM(cfg)
    tower = model.get_vision_tower()
    if tower is not None:
        tower.load_model()  # creates the CLIP parameter tree (and loads mm_vision_tower's own weights if present)
    sd = read_state_dict(model_base) if model_base is not None else {}
    if model_base is not None:
        for fn in ADAPTER_FILES:
            p = os.path.join(model_path, fn)
            if os.path.exists(p):
                sd.update(_load_file(p))
    else:
        sd = read_state_dict(model_path)
    sd = {k: v for k, v in sd.items() if "rotary_emb.inv_freq" not in k and not k.endswith("position_ids")}
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not k.startswith("model.vision_tower.")]
    if missing:
        raise RuntimeError(f"checkpoint at {model_path} lacks {len(missing)} tensors, e.g. {missing[:5]}")
    if res.unexpected_keys:
        raise RuntimeError(f"checkpoint at {model_path} has unexpected tensors, e.g. {res.unexpected_keys[:5]}")
    model = model.to(device=device, dtype=dtype).eval()
    return model


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto",
                          device="cuda", use_flash_attn=False, **kwargs):
    """Same signature and return tuple as the reference's llava/model/builder.py::load_pretrained_model (:26-173):
    (tokenizer, model, image_processor, context_len).  8/4-bit loading and LoRA merging are not built."""
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes quantised loading is not part of the B200 path")
    if "lora" in model_name.lower():
        raise NotImplementedError("LoRA merging is an offline tool in the reference (scripts/merge_lora_weights.py)")
    model = load_model(model_path, model_base, device=device, dtype=kwargs.get("torch_dtype", torch.bfloat16))
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
