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
