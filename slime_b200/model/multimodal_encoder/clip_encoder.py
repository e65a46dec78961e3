"""Drop-in for the reference's llava/model/multimodal_encoder/clip_encoder.py::CLIPVisionTower.

Same constructor, attributes and state-dict keys (`vision_tower.vision_model.*`, HF CLIP naming), but the
forward pass is slime_vision_tower_fwd (hand-written sm_100a kernels): hidden_states[select_layer] without
the CLS token, computing only the encoder layers that feed it (23 of 24 for select_layer = -2).
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from ...config import SlimeConfig
from .._runtime import EngineBinding, bind, binding_of

_CLIP_L_336 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                   image_size=336, patch_size=14, layer_norm_eps=1e-5)


def load_clip_config(name_or_path: str) -> SimpleNamespace:
    """CLIPVisionConfig.from_pretrained without the hub: a local directory's config.json, or the known
    dimensions of openai/clip-vit-large-patch14-336 (the tower every SliME release uses)."""
    cfg_file = os.path.join(name_or_path, "config.json") if os.path.isdir(name_or_path) else None
    if cfg_file and os.path.exists(cfg_file):
        with open(cfg_file) as f:
            raw = json.load(f)
        raw = raw.get("vision_config", raw)
        vals = {k: raw.get(k, v) for k, v in _CLIP_L_336.items()}
        return SimpleNamespace(**vals)
    if "clip-vit-large-patch14-336" in name_or_path or "ShareGPT4V" in name_or_path:
        return SimpleNamespace(**_CLIP_L_336)
    raise ValueError(f"cannot resolve the CLIP config of {name_or_path!r} offline")


class _Params(nn.Module):
    """Parameter container (never called): gives the tensors their HF CLIP names."""


def _linear(i, o):
    return nn.Linear(i, o)


def build_clip_param_tree(c: SimpleNamespace) -> nn.Module:
    D, I = c.hidden_size, c.intermediate_size
    root = _Params()
    vm = _Params()
    root.vision_model = vm
    emb = _Params()
    emb.class_embedding = nn.Parameter(torch.zeros(D))
    emb.patch_embedding = nn.Conv2d(3, D, kernel_size=c.patch_size, stride=c.patch_size, bias=False)
    emb.position_embedding = nn.Embedding((c.image_size // c.patch_size) ** 2 + 1, D)
    vm.embeddings = emb
    vm.pre_layrnorm = nn.LayerNorm(D, eps=c.layer_norm_eps)
    enc = _Params()
    layers = []
    for _ in range(c.num_hidden_layers):
        l = _Params()
        sa = _Params()
        sa.q_proj, sa.k_proj, sa.v_proj, sa.out_proj = _linear(D, D), _linear(D, D), _linear(D, D), _linear(D, D)
        l.self_attn = sa
        l.layer_norm1 = nn.LayerNorm(D, eps=c.layer_norm_eps)
        mlp = _Params()
        mlp.fc1, mlp.fc2 = _linear(D, I), _linear(I, D)
        l.mlp = mlp
        l.layer_norm2 = nn.LayerNorm(D, eps=c.layer_norm_eps)
        layers.append(l)
    enc.layers = nn.ModuleList(layers)
    vm.encoder = enc
    vm.post_layernorm = nn.LayerNorm(D, eps=c.layer_norm_eps)
    return root


class CLIPVisionTower(nn.Module):
    def __init__(self, vision_tower, args, delay_load=False):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = vision_tower
        self.select_layer = args.mm_vision_select_layer
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        self._args = args
        self.cfg_only = load_clip_config(vision_tower)
        self.image_processor = None
        self.weights_source = None  # where the tower's weights came from (set by load_model / the checkpoint loader)
        if not delay_load or getattr(args, "unfreeze_mm_vision_tower", False):
            self.load_model()

    def load_model(self, device_map=None):
        if self.is_loaded:
            print("{} is already loaded, `load_model` called again, skipping.".format(self.vision_tower_name))
            return
        try:  # host-side preprocessing object, only if transformers + the files are present
            from transformers import CLIPImageProcessor

            self.image_processor = CLIPImageProcessor.from_pretrained(self.vision_tower_name)
        except Exception:
            self.image_processor = None
        self.vision_tower = build_clip_param_tree(self.cfg_only)
        self._maybe_load_checkpoint()
        self.vision_tower.requires_grad_(False)
        self.is_loaded = True

    def _resolve_dir(self):
        """Local directory holding the tower's own files: `mm_vision_tower` itself, or - for a hub id such as
        'openai/clip-vit-large-patch14-336' (every released SliME config) - its snapshot in the local HF cache
        (no network: local_files_only).  None when neither exists."""
        d = self.vision_tower_name
        if os.path.isdir(d):
            return d, "directory"
        try:
            from huggingface_hub import snapshot_download

            return snapshot_download(d, local_files_only=True), "hf-cache"
        except Exception:
            return None, None

    def _maybe_load_checkpoint(self):
        """Loads the tower's own pretrained weights when they can be found and records where they came from in
        `self.weights_source` ('directory' / 'hf-cache' / None).  None means the parameters are still
        default-initialised: slime_b200.checkpoint.load_model then requires the model checkpoint to carry the
        `model.vision_tower.*` tensors and raises otherwise (the reference always loads CLIP from this name after
        from_pretrained, llava/model/builder.py:160-162)."""
        self.weights_source = None
        d, how = self._resolve_dir()
        if d is None:
            return
        sd = None
        st = os.path.join(d, "model.safetensors")
        pt = os.path.join(d, "pytorch_model.bin")
        if os.path.exists(st):
            try:
                from safetensors.torch import load_file

                sd = load_file(st)
            except ImportError:
                sd = None
        if sd is None and os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        if sd is not None:
            sd = {k: v for k, v in sd.items() if k.startswith("vision_model.") and "position_ids" not in k}
            self.vision_tower.load_state_dict(sd, strict=True)
            self.weights_source = how

    # ------------------------------------------------------------------ engine
    def _engine(self, device):
        b = binding_of(self)
        if b is None:  # stand-alone tower: private engine holding only the ViT group
            cfg = SlimeConfig.from_hf_config(self._args, self.config)
            b = EngineBinding(self, cfg, "model.vision_tower.", ("vit",))
            bind(self, b)
        return b.engine(device)

    @torch.no_grad()
    def forward(self, images):
        if self.select_feature != "patch":
            raise ValueError(f"Unexpected select feature: {self.select_feature}")
        if type(images) is list:
            return [self._engine(im.device).vision_tower(im.unsqueeze(0)).to(im.dtype) for im in images]
        return self._engine(images.device).vision_tower(images).to(images.dtype)

    # ------------------------------------------------------------------ reference properties (clip_encoder.py:60-89)
    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.vision_tower.vision_model.embeddings.class_embedding.dtype

    @property
    def device(self):
        return self.vision_tower.vision_model.embeddings.class_embedding.device

    @property
    def config(self):
        return self.cfg_only

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2
