"""Drop-in for the reference's llava/model/llava_arch.py: LlavaMetaModel / LlavaMetaForCausalLM with the same
method names, argument order and return tuples, executed by libslime_b200 through one shared SlimeEngine.

  encode_images(images, input_ids, split_sizes, attention_mask, images_mask, image_sizes, labels)
      -> (list of [1, 577 + K_b, H] features, split_sizes)                       (reference :212-269)
  get_pure_text_embedding(input_ids, attention_mask, labels) -> ([B,T,H], [B,T])   (reference :162-210)
  prepare_inputs_labels_for_multimodal(input_ids, position_ids, attention_mask, past_key_values, labels,
      images, image_sizes, images_mask) -> (None, position_ids, attention_mask, past_key_values,
      inputs_embeds, labels)                                                       (reference :274-459)

The multimodal work is always done for the whole batch at once (vision tower over all crops, adapter,
router, splice kernels); the per-sample Python loops and host syncs of the reference are gone except for the
one unavoidable read of the spliced lengths.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional

import torch
import torch.nn as nn

from ..config import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from ..mm_utils import get_anyres_image_grid_shape
from .multimodal_encoder.builder import build_vision_tower
from .multimodal_projector.builder import build_vision_projector
from .multimodal_resampler.builder import build_vision_sampler


class LlavaMetaModel:
    def __init__(self, config):
        super(LlavaMetaModel, self).__init__(config)
        if hasattr(config, "mm_vision_tower"):
            self.vision_tower = build_vision_tower(config, delay_load=True)
            self.mm_projector = build_vision_projector(config)
            self.sampler = build_vision_sampler(config)
            t = getattr(config, "mm_resampler_type", None)
            self.has_sampler = t != "identity" and t is not None and t != "spatial"
            if "unpad" in getattr(config, "mm_patch_merge_type", ""):
                raise NotImplementedError("mm_patch_merge_type '*unpad*' is not used by SliME and is not built")

    def get_vision_tower(self):
        vision_tower = getattr(self, "vision_tower", None)
        if type(vision_tower) is list:
            vision_tower = vision_tower[0]
        return vision_tower

    def initialize_vision_modules(self, model_args, fsdp=None):
        """Reference :52-119 (training-script entry): records the SliME settings on the config and builds
        the tower / projector / sampler if they do not exist yet.  Loading of pretrain_mm_mlp_adapter /
        pretrain_mm_re_sampler checkpoints is supported; fsdp wrapping is training-only and rejected."""
        if fsdp:
            raise NotImplementedError("fsdp is a training option; the B200 path is inference-only")
        self.config.mm_vision_tower = model_args.vision_tower
        if self.get_vision_tower() is None:
            self.vision_tower = build_vision_tower(model_args)
        else:
            self.vision_tower.load_model()
        vt = self.get_vision_tower()
        cfg = self.config
        cfg.use_mm_proj = True
        cfg.use_local_only = getattr(model_args, "use_local_only", False)
        cfg.use_global_only = getattr(model_args, "use_global_only", False)
        cfg.mm_projector_type = getattr(model_args, "mm_projector_type", "linear")
        cfg.mm_hidden_size = vt.hidden_size
        for k in ("mm_vision_select_layer", "mm_vision_select_feature", "mm_patch_merge_type", "mm_resampler_type",
                  "mm_resampler_topp", "mm_resampler_dim", "mm_resampler_temp"):
            setattr(cfg, k, getattr(model_args, k))
        cfg.seperator = getattr(model_args, "seperator", 1919)
        cfg.mm_learnable_gated = getattr(model_args, "mm_learnable_gated", -1)
        if getattr(self, "mm_projector", None) is None:
            self.mm_projector = build_vision_projector(cfg)
            self.sampler = build_vision_sampler(cfg)
            t = getattr(cfg, "mm_resampler_type", None)
            self.has_sampler = t != "identity" and t is not None and t != "spatial"

        def get_w(weights, keyword):
            return {k.split(keyword + ".")[1]: v for k, v in weights.items() if keyword in k}

        if getattr(model_args, "pretrain_mm_mlp_adapter", None) is not None:
            w = torch.load(model_args.pretrain_mm_mlp_adapter, map_location="cpu")
            self.mm_projector.load_state_dict(get_w(w, "mm_projector"), strict=False)
        if getattr(model_args, "pretrain_mm_re_sampler", None) is not None:
            w = torch.load(model_args.pretrain_mm_re_sampler, map_location="cpu")
            self.sampler.load_state_dict(get_w(w, "sampler"))


class LlavaMetaForCausalLM(ABC):
    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    # ------------------------------------------------------------------ helpers
    def _grids(self, image_sizes, n_local: List[int]):
        if getattr(self.config, "mm_patch_merge_type", "flat") != "spatial":
            return None
        if image_sizes is None:
            raise ValueError("mm_patch_merge_type='spatial' needs image_sizes")
        size = self.get_vision_tower().config.image_size
        pin = getattr(self.config, "image_grid_pinpoints", None)
        return [get_anyres_image_grid_shape(tuple(int(v) for v in s), pin, size) if n > 0 else (0, 0)
                for s, n in zip(image_sizes, n_local)]

    def _split_images(self, images, images_mask=None):
        """list / 5-D tensor -> list of per-sample [n_b,3,S,S] (padded crops removed when images_mask is given,
        reference :228-231, :299-302)."""
        if type(images) is list:
            per = [x.unsqueeze(0) if x.ndim == 3 else x for x in images]
        else:
            per = [images[b] for b in range(images.shape[0])]
        if images_mask is not None:
            per = [p[m.to(p.device).bool()] if m.shape[0] == p.shape[0] else p for p, m in zip(per, images_mask)]
        return per

    # ------------------------------------------------------------------ reference API
    def get_pure_text_embedding(self, input_ids, attention_mask=None, labels=None):
        """Prompt embeddings with the image placeholder slots removed and the row count kept at T by zero rows
        at the padded end (front for left padding); mask zero there.  Pure gather/indexing (torch plumbing) -
        the fused router (slime_router_fwd) consumes ids directly and never materialises this tensor."""
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        eng = self._engine(input_ids.device)
        table = eng.weights["llm.embed"]
        left = getattr(self.config, "tokenizer_padding_side", "right") == "left"
        keep = input_ids != IMAGE_TOKEN_INDEX
        B, T = input_ids.shape
        # stable partition: kept tokens first (right padding) or last (left padding), original order preserved
        order = torch.argsort((~keep if not left else keep).to(torch.int8), dim=1, stable=True)
        ids_sorted = torch.gather(input_ids, 1, order)
        keep_sorted = torch.gather(keep, 1, order)
        emb = table[ids_sorted.clamp_min(0).to(table.device)] * keep_sorted.unsqueeze(-1).to(table.device, table.dtype)
        mask = torch.gather(attention_mask, 1, order) * keep_sorted.to(attention_mask.dtype)
        max_len = getattr(self.config, "tokenizer_model_max_length", None)
        if max_len is not None:
            emb, mask = emb[:, :max_len], mask[:, :max_len]
        return emb, mask

    def encode_images(self, images, input_ids=None, split_sizes=None, attention_mask=None, images_mask=None,
                      image_sizes=None, labels=None):
        """Sampler branch of the reference (:217-255): returns per-sample [1, n_tokens, H] features."""
        if not self.get_model().has_sampler or split_sizes is None:
            raise NotImplementedError("only the SliME branch (text-guided sampler, per-sample crop stacks) is built")
        per = list(torch.split(images, split_sizes, dim=0))
        if images_mask is not None:
            per = self._split_images(per, images_mask)
        eng = self._engine(images.device)
        n_local = [p.shape[0] - 1 for p in per]
        res = eng.prefill_features(per, input_ids, attention_mask, grids=self._grids(image_sizes, n_local))
        return res, split_sizes

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                             images, image_sizes=None, images_mask=None):
        vision_tower = self.get_vision_tower()
        if vision_tower is None or images is None or input_ids.shape[1] == 1:
            return input_ids, position_ids, attention_mask, past_key_values, None, labels
        if not (type(images) is list or images.ndim == 5):
            raise NotImplementedError("SliME feeds a stack of crops per sample (5-D tensor or list)")
        out = self._spliced(input_ids, attention_mask, labels, images, image_sizes, images_mask, padded=True)
        sp = out["splice"]
        new_labels = sp["labels"] if labels is not None else None
        new_mask = None if attention_mask is None else sp["attention_mask"].to(dtype=attention_mask.dtype)
        new_pos = None if position_ids is None else sp["position_ids"]
        return None, new_pos, new_mask, past_key_values, sp["inputs_embeds"], new_labels

    # ------------------------------------------------------------------ shared implementation
    def _spliced(self, input_ids, attention_mask, labels, images, image_sizes, images_mask, padded):
        per = self._split_images(images, images_mask)
        eng = self._engine(input_ids.device if input_ids.is_cuda else per[0].device)
        n_local = [p.shape[0] - 1 for p in per]
        return eng.prefill_splice(per, input_ids, attention_mask, grids=self._grids(image_sizes, n_local),
                                  labels=labels, padded=padded)

    def initialize_vision_tokenizer(self, model_args, tokenizer):
        """Reference :461-503.  SliME runs with mm_use_im_start_end=False and mm_use_im_patch_token=False, for
        which the reference function is a no-op; the token-adding variants resize the embedding table and are
        training-time options."""
        if getattr(model_args, "mm_use_im_patch_token", False) or getattr(model_args, "mm_use_im_start_end", False):
            raise NotImplementedError("mm_use_im_start_end / mm_use_im_patch_token are not used by SliME")
