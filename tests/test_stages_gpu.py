"""Stage-by-stage and end-to-end parity of the CUDA path (through the C-ABI) against the CPU oracle
and the committed golden vectors of the unmodified reference, on the small configurations the
oracle finishes in seconds.

Tolerances (north_star: "within 1e-3 relative bf16 tolerance, bit-exact for the token-index splice"):
  * integer results - selected indices given the probabilities, lengths, cu_seqlens, attention_mask,
    position_ids, labels, and the spliced rows as a gather of the stage outputs - are compared EXACTLY;
  * floating-point stages are bf16 pipelines compared with the fp32 oracle: one bf16 rounding is
    already 1.7e-3 rel-L2, so the bounds are the values MEASURED on B200 x 1.3 (profiles/r02_stage_errors.txt):
    stages <= STAGE_TOL = 8e-3 (measured 2.3e-3..6.1e-3), last-token logits (teacher-forced selection, SURVEY.md
    8a row R) <= LAST_TOL = 8e-3 (measured 4.9e-3..6.1e-3), all-position logits <= E2E_TOL = 1.1e-2 (measured
    7.2e-3..8.2e-3) - the unmodified reference's own bf16 run sits at 0.8e-2 on the last token of the same case (golden key
    ref_bf16_rel_err_last); the full-size floor test (test_fullsize_gpu.py) shows these errors are the bf16 floor.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
STAGE_TOL = 8e-3
LAST_TOL = 8e-3
E2E_TOL = 1.1e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


_engines = {}


def setup_case(name):
    """engine + oracle outputs for one golden case (cached per process)."""
    if name in _engines:
        return _engines[name]
    from oracle import slime_oracle as O
    from oracle.gen_golden import CASES
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_inputs, synth_state_dict

    pname, over, B, n, T, ipos, ragged, isize, with_labels = CASES[name]
    cfg = preset(pname, **over)
    sd = synth_state_dict(cfg)
    px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=ipos, ragged=ragged)
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[:, : ipos + 2] = -100
        labels[labels == -200] = -100
    grids = [O.grid_shape(isize, cfg.vit_image)] * B
    with torch.no_grad():
        ora = O.prefill(sd, cfg, px, ids, mask, grids, labels=labels)
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(sd)
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}
    _engines[name] = (cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold)
    return _engines[name]


CASE_NAMES = ["tiny_spatial_b2", "tiny_global_only_crop", "tiny_flat_left_trunc", "small_wide_topp50"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_vision_tower_stage(name):
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case(name)
    feats = eng.vision_tower(px.flatten(0, 1))
    ref = torch.cat(ora["vit"], 0)
    e = rel(feats, ref)
    print(f"[{name}] vision tower rel-L2 vs fp32 oracle: {e:.3e}")
    assert e < STAGE_TOL
    eg = rel(feats.view(px.shape[0], px.shape[1], 576, -1)[:, :, ::16, :], gold["vit"])
    print(f"[{name}] vision tower rel-L2 vs reference golden: {eg:.3e}")
    assert eg < STAGE_TOL


@pytest.mark.parametrize("name", CASE_NAMES)
def test_resampler_projector_stages(name):
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case(name)
    from oracle import slime_oracle as O

    # feed the ORACLE's features so each stage is judged on its own arithmetic
    vit = torch.stack(ora["vit"]).to(torch.bfloat16)  # [B, n, 576, D]
    vit32 = vit.float()
    xg = vit[:, 0]
    glob = eng.gated_projector(xg)
    with torch.no_grad():
        ref_g = torch.stack([O.gated_projector(sd, vit32[b, 0], cfg.mm_learnable_gated) for b in range(vit.shape[0])])
    e = rel(glob, ref_g)
    print(f"[{name}] gated global projector rel-L2: {e:.3e}")
    assert e < STAGE_TOL
    r576 = eng.resampler(1, xg)
    with torch.no_grad():
        ref_r = O.resampler(sd, "model.mm_projector.attn.", vit32[:, 0])
    e = rel(r576, ref_r)
    print(f"[{name}] 576-query resampler rel-L2: {e:.3e}")
    assert e < STAGE_TOL
    if vit.shape[1] > 1:
        xl = vit[:, 1:].reshape(-1, 576, vit.shape[-1])
        lc = eng.resampler(0, xl)
        with torch.no_grad():
            ref_lc = O.resampler(sd, "model.sampler.post_qformer.", xl.float())
        e = rel(lc, ref_lc)
        print(f"[{name}] local compression rel-L2: {e:.3e}")
        assert e < STAGE_TOL
        lp = eng.projector(lc)
        with torch.no_grad():
            ref_lp = O.projection(sd, lc.float().cpu()).flatten(0, 1)
        e = rel(lp, ref_lp)
        print(f"[{name}] local projection rel-L2: {e:.3e}")
        assert e < STAGE_TOL


@pytest.mark.parametrize("name", [n for n in CASE_NAMES if n != "tiny_global_only_crop"])
def test_router_selection_exact(name):
    """The selection is integer work: given probabilities it must reproduce the reference rule exactly."""
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case(name)
    from oracle import slime_oracle as O

    # (1) the reference's own fp32 probabilities -> the reference's own selected indices (golden)
    probs = gold["probs"]
    sel_idx, sel_count = eng.router_select(probs)
    assert sel_count.cpu().tolist() == gold["sel_count"].tolist()
    for b in range(probs.shape[0]):
        k = int(gold["sel_count"][b])
        assert sel_idx[b, :k].cpu().tolist() == gold["sel_idx"][b, :k].tolist()
    # (2) full router on the CUDA path's own features: probabilities close to the oracle's, and the
    #     device selection identical to the reference rule applied to those same device probabilities
    res = eng.prefill(px, ids, mask, grids=grids, want_probs=True, want_last=False)
    p_dev = res.probs.cpu()
    for b in range(p_dev.shape[0]):
        n = ora["probs"][b].numel()
        e = rel(p_dev[b, :n], ora["probs"][b])
        print(f"[{name}] router probabilities rel-L2 (sample {b}): {e:.3e}")
        assert e < STAGE_TOL
        expect = O.top_p_select(p_dev[b, :n], cfg.mm_resampler_topp)
        k = int(res.sel_count[b])
        assert k == expect.numel()
        assert res.sel_idx[b, :k].cpu().tolist() == expect.tolist()


def test_router_select_ties_and_edges():
    cfg, eng, *_ = setup_case("tiny_spatial_b2")
    from oracle import slime_oracle as O

    g = torch.Generator().manual_seed(5)
    rows = []
    rows.append(torch.full((576,), 1.0 / 576))                     # all equal: stable order decides
    x = torch.rand(576, generator=g).to(torch.bfloat16).float()     # heavy ties (bf16-valued probabilities)
    rows.append(x / x.sum())
    rows.append(torch.softmax(torch.randn(576, generator=g) * 4, 0))
    one = torch.zeros(576)
    one[123] = 1.0
    rows.append(one)                                                # count == N-1 ... keep exactly up to the hit
    probs = torch.stack(rows)
    sel_idx, sel_count = eng.router_select(probs)
    for b in range(probs.shape[0]):
        expect = O.top_p_select(probs[b], cfg.mm_resampler_topp)
        k = int(sel_count[b])
        assert k == expect.numel(), f"row {b}: {k} vs {expect.numel()}"
        assert sel_idx[b, :k].cpu().tolist() == expect.tolist(), f"row {b}"


@pytest.mark.parametrize("name", CASE_NAMES)
def test_splice_bit_exact(name):
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case(name)
    B = px.shape[0]
    forced = ora["sel"] if px.shape[1] > 1 else None
    res = eng.prefill(px, ids, mask, grids=grids, labels=labels, forced_selection=forced, keep_stages=True)
    # integer outputs vs the unmodified reference (golden) - exact
    assert res.lengths == gold["lengths"].tolist()
    sp = res.stages["splice"]
    assert torch.equal(sp["attention_mask"].cpu(), gold["attention_mask"].bool())
    if labels is not None:
        assert torch.equal(sp["labels"].cpu(), gold["labels"])
    assert torch.equal(sp["position_ids"].cpu(), ora["position_ids"])
    cu = res.cu_seqlens.cpu().tolist()
    assert cu == [0] + list(np.cumsum(res.lengths))
    # the spliced rows are a pure gather of the stage outputs: rebuild them with torch indexing and compare bits
    emb_table = eng.weights["llm.embed"]
    left = cfg.tokenizer_padding_side == "left"
    for b in range(B):
        m = mask[b].bool()
        cid = ids[b][m]
        p = int((cid == -200).nonzero()[0]) if (cid == -200).any() else None
        parts = []
        if p is None:
            parts.append(emb_table[cid.cuda()])
        else:
            parts.append(emb_table[cid[:p].cuda()])
            if not cfg.use_local_only:
                parts.append(res.stages["glob"][b])
                if not cfg.use_global_only:
                    parts.append(emb_table[cfg.seperator][None])
            if not cfg.use_global_only and px.shape[1] > 1:
                k = int(res.sel_count[b])
                parts.append(res.stages["local_m"][b][res.sel_idx[b, :k].long()])
            parts.append(emb_table[cid[p + 1:].cuda()])
        expect = torch.cat(parts)[: res.lengths[b]]
        got = res.embeds[cu[b]:cu[b + 1]]
        assert torch.equal(got.view(torch.int16), expect.view(torch.int16)), f"sample {b}: packed rows differ"
        pad = sp["inputs_embeds"][b]
        L = res.lengths[b]
        real = pad[pad.shape[0] - L:] if left else pad[:L]
        assert torch.equal(real.view(torch.int16), expect.view(torch.int16))
        rest = pad[: pad.shape[0] - L] if left else pad[L:]
        assert (rest == 0).all()
    # and the spliced embeddings agree with the reference's inputs_embeds numerically
    e = rel(sp["inputs_embeds"][:, ::8, :], gold["embeds_rows"])
    print(f"[{name}] inputs_embeds rel-L2 vs reference golden: {e:.3e}")
    assert e < STAGE_TOL


@pytest.mark.parametrize("name", CASE_NAMES)
def test_end_to_end_logits(name):
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case(name)
    forced = ora["sel"] if px.shape[1] > 1 else None
    res = eng.prefill(px, ids, mask, grids=grids, forced_selection=forced, want_all_logits=True)
    assert res.lengths == gold["lengths"].tolist()
    e_last = rel(res.logits_last, gold["logits_last"])
    ref_bf16 = float(gold["ref_bf16_rel_err_last"][0])
    print(f"[{name}] last-token logits rel-L2 vs reference fp32 golden: {e_last:.3e} "
          f"(reference's own bf16 run: {ref_bf16:.3e})")
    assert e_last < LAST_TOL
    cu = res.cu_seqlens.cpu().tolist()
    for b in range(px.shape[0]):
        got = res.logits_all[cu[b]:cu[b + 1]]
        e = rel(got, ora["logits"][b])
        print(f"[{name}] all-position logits rel-L2 sample {b}: {e:.3e}")
        assert e < E2E_TOL
        # last-token fp32 logits and the bf16 all-position logits describe the same row
        assert rel(res.logits_last[b], got[-1].float()) < 1e-2
        assert int(res.logits_last[b].argmax()) == int(ora["logits"][b][-1].argmax()) or \
            rel(res.logits_last[b], ora["logits"][b][-1]) < E2E_TOL


def test_decoder_stage_alone():
    """Decoder on the oracle's spliced embeddings (isolates K14-K19 from upstream differences)."""
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case("tiny_spatial_b2")
    lens = ora["lengths"]
    rows = torch.cat([ora["inputs_embeds"][b, :L] for b, L in enumerate(lens)]).to(torch.bfloat16).cuda()
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(L) for L in lens]).to(torch.int32).cuda()
    last, allv, hid = eng.decoder_prefill(rows, cu, pos, lens, want_last=True, want_all=True, want_hidden=True)
    ref_last = torch.stack([lg[-1] for lg in ora["logits"]])
    e = rel(last, ref_last)
    print(f"decoder-only last-token logits rel-L2: {e:.3e}")
    assert e < STAGE_TOL


def test_engine_rejects_bad_input():
    cfg, eng, sd, (px, ids, mask, labels, grids, isize), ora, gold = setup_case("tiny_spatial_b2")
    bad = ids.clone()
    bad[0, 1] = -200  # two placeholders in one prompt
    with pytest.raises(RuntimeError, match="more than one image placeholder"):
        eng.prefill(px, bad, mask, grids=grids)
    with pytest.raises(ValueError):
        eng.prefill(px, ids, mask, grids=[(3, 1)] * px.shape[0])
