"""Generates tests/golden/preprocess_*.npz by running the REFERENCE's own `process_images`
(/root/reference/llava/mm_utils.py:231-259, unmodified, imported here) on small seeded images.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONPATH=/root/reference:/root/repo python oracle/gen_golden_preprocess.py

The image processor handed to the reference is transformers' PIL/numpy-backend CLIP processor (the arithmetic
of the pinned transformers 4.37.2: float64 rescale -> float32, float32 normalise) behind a thin adapter that
exposes `crop_size` / `size` as plain dicts, as 4.37.2 did (5.x returns a SizeDict without `.values()`, which
mm_utils.py:194 calls)."""
import os
import sys
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [  # name, (w, h), image_aspect_ratio
    ("anyres_500x700", (500, 700), "anyres"),
    ("anyres_800x330", (800, 330), "anyres"),
    ("anyres_97x61", (97, 61), "anyres"),
    ("pad_211x400", (211, 400), "pad"),
    ("plain_640x480", (640, 480), None),
    ("plain_350x900", (350, 900), None),
]


def smooth_image(rng, w, h):
    """low-frequency content + noise + saturated patches (exercises the clipping of over/undershoot)"""
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(xx / (7.0 + c) + yy / 11.0) for c in range(3)], axis=-1)
    img += rng.normal(0, 8, img.shape)
    img[h // 3: h // 3 + max(h // 8, 1), w // 4: w // 4 + max(w // 6, 1)] = 255
    img[h // 2: h // 2 + max(h // 9, 1), w // 2: w // 2 + max(w // 7, 1)] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil

    from llava.mm_utils import process_images

    hf = CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    proc = types.SimpleNamespace(crop_size={"height": 336, "width": 336}, size={"shortest_edge": 336},
                                 image_mean=list(hf.image_mean), image_std=list(hf.image_std),
                                 preprocess=hf.preprocess, __call__=hf.__call__)

    class Proc:  # callable adapter for the `else` branch of process_images (image_processor(images, ...))
        crop_size, size, image_mean, image_std = proc.crop_size, proc.size, proc.image_mean, proc.image_std
        preprocess = staticmethod(hf.preprocess)

        def __call__(self, images, return_tensors=None):
            return hf(images, return_tensors=return_tensors)

    rng = np.random.default_rng(3407)
    for name, (w, h), mode in CASES:
        img = smooth_image(rng, w, h)
        cfg = types.SimpleNamespace(image_aspect_ratio=mode, image_grid_pinpoints="[(336, 672)]")
        out = process_images([Image.fromarray(img)], Proc(), cfg)
        out = out[0] if mode == "anyres" else out
        out = out.numpy()
        # lossless compact encoding: every output value is one of 256 floats per channel (a function of the
        # byte), so store the byte codes + the table observed in the reference's output; the decode
        # lut[c][codes] is asserted identical to the reference tensor before saving
        from oracle.preprocess_oracle import clip_normalise_lut
        lut = clip_normalise_lut(hf.image_mean, hf.image_std)
        codes = np.empty(out.shape, dtype=np.uint8)
        for c in range(3):
            order = np.argsort(lut[c], kind="stable")
            pos = np.searchsorted(lut[c][order], out[:, c])
            codes[:, c] = order[np.clip(pos, 0, 255)]
            assert np.array_equal(lut[c][codes[:, c]], out[:, c]), "reference output is not in the 4.37.2 table"
        np.savez_compressed(os.path.join(GOLDEN, f"preprocess_{name}.npz"), image=img, codes=codes, lut=lut,
                            mode=np.array(mode if mode else "none"))
        print(name, img.shape, "->", out.shape)


if __name__ == "__main__":
    sys.exit(main())
