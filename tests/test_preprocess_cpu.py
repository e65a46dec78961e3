"""Pins the CPU restatement of the image pre-processing (oracle/preprocess_oracle.py): against Pillow itself,
against transformers' numpy/PIL CLIP processor, and against outputs of the reference's own process_images
(tests/golden/preprocess_*.npz, made by oracle/gen_golden_preprocess.py).  Bit-exact everywhere."""
import glob
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as O
from slime_b200.mm_utils import select_best_resolution_uhd
from slime_b200.preprocess import normalise_lut, plan_image, resize_and_pad_plan

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "preprocess_*.npz")))


def test_resize_restatement_matches_pillow():
    from PIL import Image

    rng = np.random.default_rng(7)
    cases = [(500, 700, 336, 336), (336, 336, 336, 336), (640, 480, 672, 504), (37, 53, 336, 336), (1, 1, 5, 7),
             (1500, 90, 336, 20), (200, 150, 672, 504), (3, 2, 100, 100)]
    for _ in range(12):
        cases.append(tuple(int(v) for v in (rng.integers(1, 600), rng.integers(1, 600), rng.integers(1, 700),
                                            rng.integers(1, 700))))
    for w, h, ow, oh in cases:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if rng.random() < 0.3:
            a = (a // 128 * 255).astype(np.uint8)  # hard edges: over/undershoot must clip like PIL
        ref = np.asarray(Image.fromarray(a).resize((ow, oh)))
        assert np.array_equal(O.resize_bicubic_u8(a, ow, oh), ref), (w, h, ow, oh)


def test_normalisation_table_matches_transformers_numpy_path():
    pil_clip = pytest.importorskip("transformers.models.clip.image_processing_pil_clip")
    from PIL import Image

    proc = pil_clip.CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    a = np.arange(336 * 336 * 3, dtype=np.int64).reshape(336, 336, 3)
    a = ((a * 2654435761) >> 7 & 255).astype(np.uint8)  # every byte value in every channel
    out = proc.preprocess(Image.fromarray(a), return_tensors="np")["pixel_values"][0]
    lut = O.clip_normalise_lut(proc.image_mean, proc.image_std)
    for c in range(3):
        assert np.array_equal(lut[c][a[..., c]], out[c])
    assert np.array_equal(lut, normalise_lut(proc.image_mean, proc.image_std))  # product table == oracle table


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[11:-4] for p in GOLDEN])
def test_oracle_matches_reference_process_images(path):
    g = np.load(path)
    img, mode = g["image"], str(g["mode"])
    expect = np.stack([g["lut"][c][g["codes"][:, c]] for c in range(3)], axis=1)
    h, w = img.shape[:2]
    if mode == "anyres":
        got = O.process_anyres(img, select_best_resolution_uhd((w, h), (336, 336)))
    else:
        got = O.process_single(img, mode)[None]
    assert got.shape == expect.shape
    assert np.array_equal(got, expect)


def test_host_plan_matches_oracle_geometry():
    rng = np.random.default_rng(11)
    for _ in range(300):
        w, h = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        p = plan_image(w, h, "anyres")
        tw, th = select_best_resolution_uhd((w, h), (336, 336))
        assert (p.jobs[1]["out_w"], p.jobs[1]["out_h"], p.jobs[1]["paste_x"], p.jobs[1]["paste_y"]) == \
            O.resize_and_pad_plan(w, h, tw, th) == resize_and_pad_plan(w, h, tw, th)
        assert p.n_crops == 1 + (tw // 336) * (th // 336) and p.grid == (tw // 336, th // 336)
        q = plan_image(w, h, "pad").jobs[0]
        assert q["virt_w"] == q["virt_h"] == max(w, h) and (q["out_w"], q["out_h"]) == (336, 336)
        if min(w, h) * 20 > max(w, h):
            r = plan_image(w, h, None).jobs[0]
            assert min(r["out_w"], r["out_h"]) == 336 and r["paste_x"] <= 0 and r["paste_y"] <= 0


def test_unsupported_modes_raise():
    with pytest.raises(NotImplementedError):
        plan_image(100, 100, "pad_then_devide")
