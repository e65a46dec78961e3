"""Timing of the GPU image pre-processing against PIL + numpy on the host (the reference's way), same images.
Algorithmic bytes = RGB source bytes + crop bytes written (bf16)."""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from slime_b200.mm_utils import process_images
from slime_b200.preprocess import normalise_lut, plan_image, preprocess_images, resize_and_pad_plan
from slime_b200.mm_utils import select_best_resolution_uhd


def pil_anyres(img):
    from PIL import Image

    pil = Image.fromarray(img)
    w, h = pil.size
    tw, th = select_best_resolution_uhd((w, h), (336, 336))
    nw, nh, px, py = resize_and_pad_plan(w, h, tw, th)
    canvas = Image.new("RGB", (tw, th), (0, 0, 0))
    canvas.paste(pil.resize((nw, nh)), (px, py))
    crops = [pil.resize((336, 336))]
    for i in range(0, th, 336):
        for j in range(0, tw, 336):
            crops.append(canvas.crop((j, i, j + 336, i + 336)))
    mean = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
    std = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
    out = []
    for c in crops:
        x = (np.asarray(c).astype(np.float64) * (1 / 255)).astype(np.float32)
        out.append(((x - mean) / std).transpose(2, 0, 1))
    return torch.from_numpy(np.stack(out))


def main():
    rng = np.random.default_rng(0)
    proc = types.SimpleNamespace(crop_size={"height": 336, "width": 336}, size={"shortest_edge": 336},
                                 image_mean=None, image_std=None)
    cfg = types.SimpleNamespace(image_aspect_ratio="anyres")
    for name, (w, h), B in [("672x672 x16", (672, 672), 16), ("1024x768 x16", (1024, 768), 16),
                            ("4032x3024 x4", (4032, 3024), 4), ("640x480 x64", (640, 480), 64)]:
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(B)]
        for _ in range(3):
            out = process_images(imgs, proc, cfg, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        # whole call (host planning + coefficient tables + pinned staging + H2D + 2 kernels)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            out = process_images(imgs, proc, cfg, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        call_ms = (time.perf_counter() - t0) / reps * 1e3
        # kernels only: CUDA events around the library call with everything already on the device
        from slime_b200 import _lib as L
        import ctypes as C
        lib = L.load()
        lib.slime_profile_enable(1)
        out = process_images(imgs, proc, cfg, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        ms3 = (C.c_double * 3)(); work3 = (C.c_double * 3)(); n3 = (C.c_longlong * 3)()
        lib.slime_profile_collect(ms3, work3, n3)
        lib.slime_profile_enable(0)
        kern_ms, bytes_ = ms3[2], work3[2]
        t0 = time.perf_counter()
        ref = [pil_anyres(im) for im in imgs[:max(1, B // 4)]]
        cpu_ms = (time.perf_counter() - t0) * 1e3 * B / max(1, B // 4)
        n_crops = sum(o.shape[0] for o in (out if isinstance(out, list) else list(out)))
        print(f"{name}: {n_crops} crops | GPU call {call_ms:.2f} ms ({B / call_ms * 1e3:.0f} img/s), kernels "
              f"{kern_ms:.3f} ms = {bytes_ / kern_ms / 1e6:.0f} GB/s algorithmic | PIL+numpy 1 core {cpu_ms:.1f} ms "
              f"({B / cpu_ms * 1e3:.0f} img/s) | x{cpu_ms / call_ms:.0f}")
        got = out[0] if isinstance(out, list) else out[0]
        assert torch.equal(got.cpu(), ref[0].to(torch.bfloat16)), "GPU != PIL"


if __name__ == "__main__":
    main()
