"""GPU image pre-processing in front of the hot path: the reference's `process_images`
(llava/mm_utils.py:231-259) = PIL `Image.resize` + pad / tile + `CLIPImageProcessor.preprocess`, done by
`slime_preprocess_fwd` (csrc/preprocess.cu) for a whole batch of images in two launches, bit-exact with the
PIL / numpy path.

Host side here: the integer planning (which canvas, which paste offset - reference mm_utils.py:41-132 and the HF
processor's resize / centre-crop sizes), the 3 x 256 normalisation table, one pinned staging buffer and one
host->device copy of the raw RGB bytes."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .mm_utils import select_best_resolution_uhd

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_DTYPE_CODE = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}


@dataclass
class ImagePlan:
    """Jobs (as keyword dicts of slime_resize_job fields, without src_offset / first_crop) and crop count."""
    jobs: List[dict]
    n_crops: int
    grid: Optional[Tuple[int, int]]  # (tiles_w, tiles_h) of the local crops for 'anyres'


def resize_and_pad_plan(ow: int, oh: int, tw: int, th: int) -> Tuple[int, int, int, int]:
    """(new_w, new_h, paste_x, paste_y) of resize_and_pad_image (reference mm_utils.py:99-132)."""
    scale_w = tw / ow
    scale_h = th / oh
    if scale_w < scale_h:
        new_w = tw
        new_h = min(math.ceil(oh * scale_w), th)
    else:
        new_h = th
        new_w = min(math.ceil(ow * scale_h), tw)
    return new_w, new_h, (tw - new_w) // 2, (th - new_h) // 2


def plan_image(w: int, h: int, mode: Optional[str], crop: int = 336, shortest_edge: int = 336,
               mean: Sequence[float] = CLIP_MEAN) -> ImagePlan:
    plain = dict(src_w=w, src_h=h, virt_w=w, virt_h=h, virt_x=0, virt_y=0, fill=(0, 0, 0))
    if mode == "anyres":
        # process_anyres_image (mm_utils.py:177-210): crop 0 = whole image squashed to shortest_edge^2, then the
        # tiles of the aspect-preserving resize pasted centred on the best canvas
        tw, th = select_best_resolution_uhd((w, h), (crop, crop))
        new_w, new_h, px, py = resize_and_pad_plan(w, h, tw, th)
        if shortest_edge != crop:
            raise ValueError("anyres needs processor.size['shortest_edge'] == crop_size (336 / 336 in every SliME config)")
        jobs = [dict(plain, out_w=crop, out_h=crop, canvas_w=crop, canvas_h=crop, paste_x=0, paste_y=0, rel_crop=0),
                dict(plain, out_w=new_w, out_h=new_h, canvas_w=tw, canvas_h=th, paste_x=px, paste_y=py, rel_crop=1)]
        return ImagePlan(jobs, 1 + (tw // crop) * (th // crop), (tw // crop, th // crop))
    if mode == "pad":
        # expand2square with the mean colour (mm_utils.py:214-228), then the processor (square -> no crop)
        side = max(w, h)
        fill = tuple(int(x * 255) for x in mean)
        vx, vy = ((0, (w - h) // 2) if w > h else ((h - w) // 2, 0)) if w != h else (0, 0)
        job = dict(src_w=w, src_h=h, virt_w=side, virt_h=side, virt_x=vx, virt_y=vy, fill=fill)
        w, h = side, side
    elif mode in (None, "square", "default"):
        job = dict(plain)
    else:
        raise NotImplementedError(f"image_aspect_ratio={mode!r}: only 'anyres' (SliME), 'pad' and the plain processor "
                                  "are built ('pad_then_devide' / 'any_res' are unused by the released configs)")
    # CLIPImageProcessor: resize so the SHORT side is shortest_edge (long side int(short_edge * long / short)),
    # then centre-crop crop x crop (HF image_transforms.get_resize_output_image_size / center_crop)
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    if new_w < crop or new_h < crop:
        raise ValueError("centre crop larger than the resized image is not supported")
    job.update(out_w=new_w, out_h=new_h, canvas_w=crop, canvas_h=crop, paste_x=-((new_w - crop) // 2),
               paste_y=-((new_h - crop) // 2), rel_crop=0)
    return ImagePlan([job], 1, None)


def normalise_lut(mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD,
                  rescale: float = 1 / 255) -> np.ndarray:
    """[3, 256] float32 table of rescale + normalize for every byte value, in the operation order of
    transformers' numpy transforms (rescale: float64 product -> float32; normalize: float32 (x - mean) / std)."""
    u = np.arange(256, dtype=np.uint8)
    x = (u.astype(np.float64) * rescale).astype(np.float32)
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return np.ascontiguousarray(((x[None, :] - m[:, None]) / s[:, None]).astype(np.float32))


def _as_rgb_array(image) -> np.ndarray:
    if isinstance(image, np.ndarray):
        arr = image
    elif torch.is_tensor(image):
        arr = image.cpu().numpy()
    else:  # PIL.Image
        if getattr(image, "mode", "RGB") != "RGB":
            image = image.convert("RGB")  # CLIPImageProcessor do_convert_rgb
        arr = np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError(f"expected an RGB uint8 image [H, W, 3], got {arr.dtype} {arr.shape}")
    return np.ascontiguousarray(arr)


def preprocess_images(images: Sequence, mode: Optional[str] = "anyres", crop: int = 336, shortest_edge: int = 336,
                      mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD,
                      dtype: torch.dtype = torch.float32, device=None) -> Tuple[List[torch.Tensor], List[ImagePlan]]:
    """-> ([n_i, 3, crop, crop] CUDA tensor per image (views of one allocation), the plans).  Raises without a GPU /
    the CUDA library: there is no CPU path."""
    if dtype not in _DTYPE_CODE:
        raise ValueError(f"unsupported dtype {dtype}")
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    arrays = [_as_rgb_array(im) for im in images]
    plans = [plan_image(a.shape[1], a.shape[0], mode, crop, shortest_edge, mean) for a in arrays]
    n_jobs = sum(len(p.jobs) for p in plans)
    jobs = (L.ResizeJob * n_jobs)()
    offsets, total_bytes = [], 0
    for a in arrays:
        offsets.append(total_bytes)
        total_bytes += (a.size + 255) // 256 * 256
    j, first = 0, 0
    crop_starts = []
    for img_i, p in enumerate(plans):
        crop_starts.append(first)
        for jd in p.jobs:
            q = jobs[j]
            q.src_offset = offsets[img_i]
            for k in ("src_w", "src_h", "virt_w", "virt_h", "virt_x", "virt_y", "out_w", "out_h", "canvas_w",
                      "canvas_h", "paste_x", "paste_y"):
                setattr(q, k, int(jd[k]))
            q.first_crop = first + jd["rel_crop"]
            for c in range(3):
                q.fill[c] = int(jd["fill"][c])
            j += 1
        first += p.n_crops
    staging = torch.empty(total_bytes, dtype=torch.uint8, pin_memory=True)
    stage_np = staging.numpy()
    for a, off in zip(arrays, offsets):
        stage_np[off:off + a.size] = a.reshape(-1)
    with torch.cuda.device(device):
        src = staging.to(device, non_blocking=True)
        out = torch.empty(first, 3, crop, crop, dtype=dtype, device=device)
        ws_bytes = lib.slime_preprocess_workspace_bytes(C.cast(jobs, C.c_void_p), n_jobs)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        lut = normalise_lut(mean, std)
        rc = lib.slime_preprocess_fwd(L.ptr(src), C.cast(jobs, C.c_void_p), n_jobs, crop,
                                      lut.ctypes.data_as(C.c_void_p), L.ptr(out), _DTYPE_CODE[dtype], L.ptr(ws),
                                      ws_bytes, L.stream_ptr())
        L.check(rc, "preprocess_fwd")
        # src / ws are only read by work already enqueued on this stream; the caching allocator keeps them
        # stream-ordered, so dropping the references here is safe
    return [out[s:s + p.n_crops] for s, p in zip(crop_starts, plans)], plans
