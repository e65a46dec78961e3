"""Host-side integer math of the hot path (reference llava/mm_utils.py:41-97,156-174): which crop grid
the reference's slicer produced for an image, needed to raster-merge the local tokens."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple


def _factor_pairs(n: int) -> List[Tuple[int, int]]:
    return [(i, n // i) for i in range(1, n + 1) if n % i == 0]


def select_best_resolution_uhd(original_size: Tuple[int, int], processor_size: Tuple[int, int] = (336, 336)
                               ) -> Tuple[int, int]:
    """Best (width, height) canvas, a multiple of the 336 px tile in both directions, among the
    factorisations of ceil(area ratio) and its neighbours (at most 7 tiles).  Same preference order
    as the reference: larger effective resolution first, then less wasted canvas; the first
    candidate wins ties."""
    tw, th = processor_size
    ow, oh = original_size
    scale = math.ceil(ow * oh / (tw * th))
    if scale > 6:
        scale = 6
    elif scale == 1:
        scale = 2
    if scale <= 2:
        candidates = _factor_pairs(scale) + _factor_pairs(scale + 1)
    else:
        candidates = _factor_pairs(scale - 1) + _factor_pairs(scale) + _factor_pairs(scale + 1)
    best, best_eff, best_waste = None, 0, float("inf")
    for w_tiles, h_tiles in candidates:
        w, h = w_tiles * tw, h_tiles * th
        s = min(w / ow, h / oh)
        eff = min(int(ow * s) * int(oh * s), ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (w, h), eff, waste
    return best


def get_anyres_image_grid_shape(image_size: Sequence[int], grid_pinpoints=None, patch_size: int = 336
                                ) -> Tuple[int, int]:
    """(num_patch_width, num_patch_height).  `grid_pinpoints` is accepted for signature parity; the
    reference computes a pinpoint-based resolution and then overwrites it with the UHD choice
    (mm_utils.py:171-173), so only the UHD rule matters.  The tile is hard-coded to 336 there."""
    w, h = select_best_resolution_uhd(tuple(image_size), (336, 336))
    return w // patch_size, h // patch_size


def _processor_geometry(image_processor):
    """(crop, shortest_edge, mean, std) from a CLIPImageProcessor-like object (dict or SizeDict attributes)."""
    def pick(obj, key, default):
        if obj is None:
            return default
        if isinstance(obj, dict):
            return obj.get(key, default)
        return getattr(obj, key, None) or default

    crop = int(pick(getattr(image_processor, "crop_size", None), "height", 336))
    short = int(pick(getattr(image_processor, "size", None), "shortest_edge", crop))
    mean = tuple(getattr(image_processor, "image_mean", None) or (0.48145466, 0.4578275, 0.40821073))
    std = tuple(getattr(image_processor, "image_std", None) or (0.26862954, 0.26130258, 0.27577711))
    return crop, short, mean, std


def process_anyres_image(image, processor, grid_pinpoints=None, dtype=None, device=None):
    """reference mm_utils.py:177-210, on the GPU: -> [1 + tiles, 3, 336, 336] CUDA tensor (float32 by default,
    bit-identical to the reference's CPU tensor)."""
    import torch

    from .preprocess import preprocess_images

    crop, short, mean, std = _processor_geometry(processor)
    outs, _ = preprocess_images([image], "anyres", crop, short, mean, std, dtype or torch.float32, device)
    return outs[0]


def process_images(images, image_processor, model_cfg, dtype=None, device=None):
    """reference mm_utils.py:231-259 with the same return convention (a stacked tensor when every image produced
    the same number of crops, else a list), computed by one batched GPU call.  `dtype` lets the caller get the
    model dtype directly instead of float32 + `.to(dtype)`."""
    import torch

    from .preprocess import preprocess_images

    mode = getattr(model_cfg, "image_aspect_ratio", None)
    crop, short, mean, std = _processor_geometry(image_processor)
    outs, plans = preprocess_images(list(images), mode, crop, short, mean, std, dtype or torch.float32, device)
    if mode != "anyres":
        return torch.stack([o[0] for o in outs], dim=0)  # [B, 3, 336, 336] like image_processor(images)
    if all(o.shape == outs[0].shape for o in outs):
        return torch.stack(outs, dim=0)
    return outs
