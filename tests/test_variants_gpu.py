"""Configuration switches and edge cases of the path, CUDA vs the (golden-pinned) CPU oracle on the tiny model:
use_global_only / use_local_only (reference llava_arch.py:215-216,249-255), mm_learnable_gated in {0, 1}
(multimodal_projector/builder.py:198-201), a batch that mixes crop counts (list input, llava_arch.py:282-286),
a prompt without an image placeholder (llava_arch.py:369-376), and top-p extremes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("over", [dict(use_global_only=True), dict(use_local_only=True), dict(mm_learnable_gated=0),
                                  dict(mm_learnable_gated=1), dict(mm_resampler_topp=0.05),
                                  dict(mm_resampler_topp=1.5)])
def test_config_switches(over):
    from oracle import slime_oracle as O
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_inputs, synth_state_dict

    cfg = preset("tiny", **over)
    sd = synth_state_dict(cfg)
    px, ids, mask = synth_inputs(cfg, 2, 5, 20, image_pos=6, ragged=True)
    grids = [(2, 2)] * 2
    with torch.no_grad():
        ora = O.prefill(sd, cfg, px, ids, mask, grids)
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(sd)
    forced = None if cfg.use_global_only else ora["sel"]
    res = eng.prefill(px, ids, mask, grids=grids, forced_selection=forced)
    assert res.lengths == ora["lengths"], (over, res.lengths, ora["lengths"])
    last = torch.stack([lg[-1] for lg in ora["logits"]])
    e = rel(res.logits_last, last)
    print(over, "logits rel-L2", e)
    assert e < 8e-3  # measured 4.8e-3..6.0e-3 on B200 (x 1.3)
    if not cfg.use_global_only:
        # un-forced: the device selection obeys the rule on the device probabilities
        r2 = eng.prefill(px, ids, mask, grids=grids, want_probs=True, want_last=False, run_decoder=False)
        for b in range(2):
            expect = O.top_p_select(r2.probs[b].cpu(), cfg.mm_resampler_topp)
            k = int(r2.sel_count[b])
            assert r2.sel_idx[b, :k].cpu().tolist() == expect.tolist()
        if over.get("mm_resampler_topp") == 1.5:
            assert int(r2.sel_count[0]) == 4 * cfg.mm_resampler_dim  # threshold above any cumulative sum: keep all
        if over.get("mm_resampler_topp") == 0.05:
            assert int(r2.sel_count[0]) < 64


def test_mixed_crop_counts_and_missing_placeholder():
    """Samples with 5, 1 and 3 crops in one batch ('flat' merge), the last prompt without an image placeholder."""
    from oracle import slime_oracle as O
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_inputs, synth_state_dict

    cfg = preset("tiny", mm_patch_merge_type="flat")
    sd = synth_state_dict(cfg)
    px, ids, mask = synth_inputs(cfg, 3, 5, 18, image_pos=4, ragged=False)
    per = [px[0], px[1][:1], px[2][:3]]
    ids[2, 4] = 77  # sample 2 has no placeholder: its image features are computed but not spliced
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(sd)
    with torch.no_grad():
        enc = [O.encode_images(sd, cfg, per[b][None], ids[b:b + 1], mask[b:b + 1], [None])for b in range(3)]
        feats = [e["feats"][0] for e in enc]
        emb, am, pid, lab, lens = O.splice(sd["model.embed_tokens.weight"], ids, mask, None, feats)
        logits = O.llama_prefill(sd, cfg, emb, lens)
    forced = [enc[b]["sel"][0] for b in range(3)]
    res = eng.prefill(per, ids, mask, forced_selection=forced)
    assert res.lengths == lens
    assert lens[2] == 18 and lens[1] == 17 + 577
    last = torch.stack([lg[-1] for lg in logits])
    e = rel(res.logits_last, last)
    print("mixed batch logits rel-L2", e)
    assert e < 8e-3  # measured 4.8e-3..6.0e-3 on B200 (x 1.3)


@pytest.mark.parametrize("images,per_image", [(3, 5), (4, 1), (2, 3), (27, 5)])
def test_vision_tower_split_output_is_a_regrouping(images, per_image):
    """slime_vision_tower_fwd_split: the same features, global crops of all images first and the local crops behind them
    (the per-sample slices of reference llava_arch.py:212-225 as two contiguous views); 27 x 5 = 135 crops crosses the
    128-crop pass boundary of the tower.  The prefill that uses it equals the one that gathers (keep_stages path)."""
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_inputs, synth_state_dict

    cfg = preset("tiny")
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(synth_state_dict(cfg))
    torch.manual_seed(images * 7 + per_image)
    px = torch.randn(images * per_image, 3, cfg.vit_image, cfg.vit_image, device="cuda").to(torch.bfloat16)
    feats = eng.vision_tower(px).view(images, per_image, cfg.vit_patches, cfg.vit_hidden)
    g, l = eng.vision_tower_split(px, per_image)
    assert torch.equal(g, feats[:, 0])
    assert torch.equal(l, feats[:, 1:].reshape(-1, cfg.vit_patches, cfg.vit_hidden))
    if per_image == 5 and images == 3:
        pxs, ids, mask = synth_inputs(cfg, 3, 5, 20, image_pos=6, ragged=True)
        a = eng.prefill(pxs, ids, mask, grids=[(2, 2)] * 3)
        b = eng.prefill(pxs, ids, mask, grids=[(2, 2)] * 3, keep_stages=True)
        assert a.lengths == b.lengths and torch.equal(a.logits_last, b.logits_last)
