"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, transformers 5.5.0) on the synthetic weights/inputs of slime_b200/synth.py.

Run here (the container that has /root/reference):   python -m oracle.gen_golden
The fixtures are committed; the GPU box never needs the reference tree.

Every case records the per-stage tensors of the reference's own modules (sub-sampled where large):
vision tower features, gated global projection, local compression, merged local projection, router
probabilities + the reference's selected indices, the spliced inputs_embeds / attention_mask /
position_ids / labels, and the logits - in fp32, plus the error of the reference's OWN bf16 run
against its fp32 run, which calibrates the end-to-end tolerance of the bf16 CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_harness import build_reference  # noqa: E402
from slime_b200.config import preset  # noqa: E402
from slime_b200.synth import synth_inputs  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (preset, overrides, batch, crops, prompt_len, image_pos, ragged, image_size, with_labels)
CASES = {
    "tiny_spatial_b2": ("tiny", {}, 2, 5, 24, 5, True, (672, 672), True),
    "tiny_global_only_crop": ("tiny", {}, 1, 1, 16, 3, False, (336, 336), False),
    "tiny_flat_left_trunc": ("tiny", dict(mm_patch_merge_type="flat", tokenizer_padding_side="left",
                                          tokenizer_model_max_length=900), 3, 4, 40, 7, True, (672, 672), True),
    "small_wide_topp50": ("small", dict(mm_resampler_topp=0.5), 1, 7, 32, 9, False, (1008, 672), False),
    # the cross-attention router (mm_resampler_type='qformer', reference multimodal_resampler/builder.py:94-162)
    "tiny_qformer_router_b2": ("tiny", dict(mm_resampler_type="qformer", mm_resampler_topp=0.6), 2, 5, 24, 5, True,
                               (672, 672), False),
}


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


@torch.no_grad()
def run_case(name):
    pname, over, B, n, T, ipos, ragged, isize, with_labels = CASES[name]
    cfg = preset(pname, **over)
    model = build_reference(cfg, dtype=torch.float32)
    px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=ipos, ragged=ragged)
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[:, : ipos + 2] = -100
        labels[labels == -200] = -100
    sizes = [isize] * B
    inner = model.get_model()
    tower = model.get_vision_tower()
    out = {}
    feats = [tower(px[b]) for b in range(B)]
    out["vit"] = torch.stack(feats)[:, :, ::16, :]
    out["glob"] = torch.stack([inner.mm_projector(f[0]) for f in feats])[:, ::8, :]
    if n > 1:
        lc = [inner.sampler.post_qformer(f[1:]) for f in feats]
        out["local_c"] = torch.stack(lc)[:, :, ::8, :]
    # full pipeline through the reference's own entry points
    _, pos_ids, attn, _, embeds, new_labels = model.prepare_inputs_labels_for_multimodal(
        ids, None, mask, None, labels, px, image_sizes=sizes)
    res = model(input_ids=ids, attention_mask=mask, images=px, image_sizes=sizes, labels=labels, use_cache=False)
    logits = res.logits
    out["embeds_rows"] = embeds[:, ::8, :]
    out["embeds_norm"] = embeds.norm(dim=-1)
    out["attention_mask"] = attn
    out["labels"] = new_labels if new_labels is not None else torch.zeros(0)
    lens = attn.sum(1)
    out["lengths"] = lens
    if cfg.tokenizer_padding_side == "left":
        last = torch.stack([logits[b, -1] for b in range(B)])
    else:
        last = torch.stack([logits[b, int(lens[b]) - 1] for b in range(B)])
    out["logits_last"] = last
    out["logits_rows"] = logits[:, ::32, :]
    if with_labels:
        out["loss"] = res.loss.reshape(1)
    # router internals: re-run the reference selector on the reference's merged local features
    if n > 1:
        text_e, text_m = model.get_pure_text_embedding(ids, mask, labels)
        probs, sels = [], []
        for b in range(B):
            lp = inner.mm_projector(lc[b])
            if cfg.mm_patch_merge_type == "spatial":
                from llava.mm_utils import get_anyres_image_grid_shape  # type: ignore

                w, h = get_anyres_image_grid_shape(sizes[b], model.config.image_grid_pinpoints, cfg.vit_image)
                g = inner.sampler.grid_size
                lm = lp.view(h, w, g, g, -1).permute(0, 2, 1, 3, 4).contiguous().flatten(0, 3)
                out["grid_wh"] = torch.tensor([w, h])
            else:
                lm = lp.flatten(0, 1)
            sim = inner.sampler.selector(lm, text_e[b], text_m[b])
            pr = torch.softmax(sim / cfg.mm_resampler_temp, dim=-1)
            sp, si = torch.sort(pr, descending=True)
            cnt = int((torch.cumsum(sp, 0) <= cfg.mm_resampler_topp).sum())
            keep = si[: cnt + 1] if cnt < si.numel() else si
            probs.append(pr)
            sels.append(torch.sort(keep).values)
            if b == 0:
                out["local_m"] = lm[::8, :]
        out["probs"] = torch.stack(probs)
        out["sel_count"] = torch.tensor([s.numel() for s in sels])
        width = max(s.numel() for s in sels)
        out["sel_idx"] = torch.stack([torch.cat([s, s.new_full((width - s.numel(),), -1)]) for s in sels])
    # the reference's own bf16 execution vs its fp32 execution (tolerance calibration)
    mb = build_reference(cfg, dtype=torch.bfloat16)
    resb = mb(input_ids=ids, attention_mask=mask, images=px.to(torch.bfloat16), image_sizes=sizes, use_cache=False)
    if resb.logits.shape == logits.shape:
        lb = resb.logits.float()
        idx = [(-1 if cfg.tokenizer_padding_side == "left" else int(lens[b]) - 1) for b in range(B)]
        out["ref_bf16_rel_err_last"] = torch.tensor(
            [rel(torch.stack([lb[b, idx[b]] for b in range(B)]), last)])
    else:  # the bf16 router kept a different number of tokens
        out["ref_bf16_rel_err_last"] = torch.tensor([float("nan")])
    out["ref_bf16_len"] = torch.tensor(resb.logits.shape[1]).reshape(1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                        **{k: v.detach().cpu().numpy() for k, v in out.items()})
    print(f"{name}: lengths {lens.tolist()} ref bf16-vs-fp32 rel err {out['ref_bf16_rel_err_last'].item():.3e} "
          f"bf16 L {int(out['ref_bf16_len'])}")


if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for case in (sys.argv[1:] or CASES):
        run_case(case)
