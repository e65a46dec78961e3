"""Pins the CPU oracle against the UNMODIFIED reference running live in this container (/root/reference, read-only;
skipped where the tree is absent, e.g. on the GPU box), on randomised inputs the committed golden fixtures do not
cover: arbitrary attention-mask patterns (gaps, left / right padding of the INPUT), both output padding sides,
truncation at tokenizer_model_max_length, labels, prompts without an image placeholder, 1 / 3 / 5 crops per sample,
flat and spatial merge, several top-p values.  Integer outputs (attention mask, labels, position ids, lengths) must be
exact; embeddings and last-token logits agree to fp32 summation order (reference llava_arch.py:162-459,
multimodal_resampler/builder.py:248-281, llava_llama.py:57-104)."""
import random

import pytest
import torch

from oracle import ref_harness
from oracle import slime_oracle as O
from slime_b200.config import preset
from slime_b200.synth import synth_inputs, synth_state_dict

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


def _build(router):
    cfg = preset("tiny", mm_patch_merge_type="flat", mm_resampler_type=router)
    torch.manual_seed(0)
    model = ref_harness.build_reference(cfg, dtype=torch.float32)
    return cfg, model, synth_state_dict(cfg)


@pytest.fixture(scope="module")
def live():
    return _build("cosine")


@pytest.fixture(scope="module")
def live_qformer():
    """the cross-attention router (TextGuidedRouterAttention, multimodal_resampler/builder.py:94-162)"""
    return _build("qformer")


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def make_trial(cfg, seed):
    rng = random.Random(seed)
    B = rng.choice([1, 2, 3])
    n = rng.choice([1, 3, 5])
    T = rng.randint(8, 36)
    ipos = rng.randint(0, T - 1)
    px, ids, mask = synth_inputs(cfg, B, n, T, seed=1000 + seed, image_pos=ipos, ragged=False)
    for b in range(B):  # arbitrary mask patterns: right padding, left padding, a hole - never on the placeholder
        kind = rng.choice(["full", "right", "left", "hole"])
        if kind == "right" and ipos < T - 1:
            mask[b, rng.randint(ipos + 1, T - 1):] = 0
        elif kind == "left" and ipos > 0:
            mask[b, : rng.randint(1, ipos)] = 0
        elif kind == "hole":
            j = rng.randint(0, T - 1)
            if j != ipos:
                mask[b, j] = 0
    if B > 1 and rng.random() < 0.3:  # a prompt without an image placeholder (llava_arch.py:369-376)
        ids[B - 1, ipos] = 7
    labels = None
    if rng.random() < 0.6:
        labels = ids.clone()
        labels[:, : rng.randint(0, T // 2)] = -100
        labels[labels == -200] = -100
    over = dict(tokenizer_padding_side=rng.choice(["right", "left"]),
                tokenizer_model_max_length=rng.choice([None, None, 200, 450]),
                mm_resampler_topp=rng.choice([0.3, 0.6, 0.95]),
                mm_patch_merge_type="spatial" if (n == 5 and rng.random() < 0.5) else "flat")
    return px, ids, mask, labels, over


@pytest.mark.parametrize("seed", range(14))
def test_oracle_equals_live_reference(live, seed):
    _check(live, seed)


@pytest.mark.parametrize("seed", [0, 1, 4, 9, 13])
def test_oracle_equals_live_reference_qformer_router(live_qformer, seed):
    _check(live_qformer, seed)


def _check(built, seed):
    cfg0, model, sd = built
    px, ids, mask, labels, over = make_trial(cfg0, seed)
    cfg = cfg0.replace(**over)
    B = px.shape[0]
    sizes = [(672, 672)] * B
    model.config.tokenizer_padding_side = cfg.tokenizer_padding_side
    model.config.tokenizer_model_max_length = cfg.tokenizer_model_max_length
    model.config.mm_patch_merge_type = cfg.mm_patch_merge_type
    model.get_model().sampler.topp = cfg.mm_resampler_topp
    with torch.no_grad():
        _, pos, attn, _, emb, lab = model.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, px,
                                                                               image_sizes=sizes)
        out = model(input_ids=ids, attention_mask=mask, images=px, image_sizes=sizes, labels=labels, use_cache=False)
        grids = [O.grid_shape(sizes[0], cfg.vit_image)] * B
        res = O.prefill(sd, cfg, px, ids, mask, grids, labels=labels)
    # ---- integer outputs: exact ----
    lens = attn.sum(1).tolist()
    assert res["lengths"] == lens, (over, res["lengths"], lens)
    assert torch.equal(res["attention_mask"], attn.bool())
    if lab is not None:
        assert torch.equal(res["labels"], lab)
    if pos is not None:
        assert torch.equal(res["position_ids"], pos)
    # ---- floats: same arithmetic, fp32 summation order only ----
    assert res["inputs_embeds"].shape == emb.shape
    assert rel(res["inputs_embeds"], emb) < 2e-5
    left = cfg.tokenizer_padding_side == "left"
    for b in range(B):
        ref_last = out.logits[b, -1] if left else out.logits[b, lens[b] - 1]
        assert rel(res["logits"][b][-1], ref_last) < 2e-4, (seed, b)


def test_grid_math_equals_live_reference():
    """select_best_resolution_uhd / get_anyres_image_grid_shape (reference llava/mm_utils.py:41-97,156-174) - the host
    integer math that decides the crop grid and therefore the raster order of the local tokens - on 6000 random image
    sizes plus the degenerate ones: the framework's function AND the oracle's restatement must return the reference's."""
    import sys

    if ref_harness.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    from llava import mm_utils as R  # type: ignore

    from slime_b200 import mm_utils as M

    rng = random.Random(7)
    sizes = [(rng.randint(1, 4000), rng.randint(1, 4000)) for _ in range(3000)]
    sizes += [(rng.randint(300, 1100), rng.randint(300, 1100)) for _ in range(3000)]  # around the tile multiples
    sizes += [(336, 336), (672, 672), (1008, 1008), (1344, 1344), (1, 1), (1, 5000), (5000, 1), (337, 336), (671, 673)]
    pin = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]
    for size in sizes:
        ref_res = R.select_best_resolution_uhd(size, (336, 336))
        assert tuple(M.select_best_resolution_uhd(size, (336, 336))) == tuple(ref_res), size
        ref_grid = tuple(R.get_anyres_image_grid_shape(size, pin, 336))
        assert tuple(M.get_anyres_image_grid_shape(size, pin, 336)) == ref_grid, size
        assert tuple(O.grid_shape(size, 336)) == ref_grid, size


def test_image_preprocessing_equals_live_reference():
    """The reference's own process_images (llava/mm_utils.py:231-259: anyres slicing incl. the bicubic global view,
    'pad' and plain CLIP modes) run live on random images of random sizes; the pre-processing oracle - which the CUDA
    kernels are compared with bit for bit on the GPU - must return exactly the same float32 tensors."""
    import sys
    import types

    import numpy as np

    pil_clip = pytest.importorskip("transformers.models.clip.image_processing_pil_clip")
    from PIL import Image

    if ref_harness.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    from llava.mm_utils import process_images  # type: ignore

    from oracle import preprocess_oracle as PO
    from slime_b200.mm_utils import select_best_resolution_uhd

    hf = pil_clip.CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})

    class Proc:  # transformers 5.x returns SizeDict objects; the reference (4.37.2) expects plain dicts
        crop_size, size = {"height": 336, "width": 336}, {"shortest_edge": 336}
        image_mean, image_std = list(hf.image_mean), list(hf.image_std)
        preprocess = staticmethod(hf.preprocess)

        def __call__(self, images, return_tensors=None):
            return hf(images, return_tensors=return_tensors)

    rng = np.random.default_rng(2024)
    trials = [("anyres", 640, 480), ("anyres", 97, 1301), ("pad", 500, 123), (None, 400, 700)]
    for _ in range(8):
        trials.append((["anyres", "anyres", "pad", None][int(rng.integers(0, 4))], int(rng.integers(40, 1400)),
                       int(rng.integers(40, 1400))))
    for mode, w, h in trials:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if rng.random() < 0.5:  # smooth content with saturated patches: bicubic over/undershoot must clip like PIL
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.clip(np.stack([127 + 125 * np.sin(xx / (5.0 + c) + yy / 9.0) for c in range(3)], -1), 0, 255).astype(np.uint8)
            img[h // 3: h // 3 + h // 6 + 1, w // 4: w // 4 + w // 5 + 1] = 255
        cfg = types.SimpleNamespace(image_aspect_ratio=mode, image_grid_pinpoints="[(336, 672)]")
        ref = process_images([Image.fromarray(img)], Proc(), cfg)
        ref = (ref[0] if mode == "anyres" else ref).numpy()
        if mode == "anyres":
            got = PO.process_anyres(img, select_best_resolution_uhd((w, h), (336, 336)))
        else:
            got = PO.process_single(img, mode)[None]
        assert got.shape == ref.shape, (mode, w, h, got.shape, ref.shape)
        assert np.array_equal(got, ref), (mode, w, h, float(np.abs(got - ref).max()))


@pytest.mark.parametrize("seed,n,T", [(0, 5, 20), (1, 3, 12), (2, 1, 9)])
def test_greedy_generation_equals_live_reference(live, seed, n, T):
    """The reference's generate() (llava_llama.py:106-144 -> HF greedy search with a KV cache on the spliced
    inputs_embeds) against the definition the decode-step tests use on the GPU: token t+1 = argmax of a fresh prefill of
    the sequence grown by token t (tests/test_decode_gpu.py compares the CUDA decode steps with exactly that)."""
    cfg, model, sd = live
    px, ids, mask = synth_inputs(cfg, 1, n, T, seed=500 + seed, image_pos=min(3, T - 1), ragged=False)
    model.config.tokenizer_padding_side = "right"
    model.config.tokenizer_model_max_length = None
    model.config.mm_patch_merge_type = "flat"
    model.get_model().sampler.topp = cfg.mm_resampler_topp
    steps = 6
    with torch.no_grad():
        ref = model.generate(ids, images=px, image_sizes=[(672, 672)], attention_mask=mask, max_new_tokens=steps,
                             do_sample=False, use_cache=True, pad_token_id=0)
        res = O.prefill(sd, cfg, px, ids, mask, [None])
        seq = res["inputs_embeds"][0, : res["lengths"][0]]
        embed, toks = sd["model.embed_tokens.weight"], []
        for _ in range(steps):
            t = int(O.llama_prefill(sd, cfg, seq[None], [seq.shape[0]])[0][-1].argmax())
            toks.append(t)
            seq = torch.cat([seq, embed[t][None]])
    assert ref.shape == (1, steps)
    assert ref[0].tolist() == toks
