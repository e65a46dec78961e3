"""GPU tests of the drop-in module API: each reference-named module/method, called the way the reference's own
callers call it (llava_arch.py:222-255, llava_llama.py:57-144), agrees with the CPU oracle / golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.fixture(scope="module")
def setup():
    from oracle import slime_oracle as O
    from slime_b200.synth import synth_inputs, synth_state_dict
    from tests.test_shims_cpu import make_model

    cfg, model = make_model("tiny")
    sd = synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device="cuda", dtype=torch.bfloat16).eval()
    px, ids, mask = synth_inputs(cfg, 2, 5, 24, image_pos=5, ragged=True)
    with torch.no_grad():
        ora = O.prefill(sd, cfg, px, ids, mask, [(2, 2)] * 2)
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "tiny_spatial_b2.npz")).items()}
    return cfg, model, sd, px, ids, mask, ora, gold


def test_module_level_calls(setup):
    cfg, model, sd, px, ids, mask, ora, gold = setup
    from oracle import slime_oracle as O

    inner = model.get_model()
    tower = model.get_vision_tower()
    imgs = px[0].cuda().to(torch.bfloat16)
    feats = tower(imgs)                                   # llava_arch.py:222
    assert feats.shape == (5, 576, cfg.vit_hidden) and feats.dtype == imgs.dtype
    assert rel(feats, ora["vit"][0]) < TOL
    g = inner.mm_projector(feats[0])                      # :224 global -> gated branch
    assert g.shape == (576, cfg.hidden_size)
    with torch.no_grad():
        ref_g = O.gated_projector(sd, feats[0].float().cpu())
    assert rel(g, ref_g) < TOL
    lc = inner.sampler.post_qformer(feats[1:])            # :226 local compression
    assert lc.shape == (4, 144, cfg.vit_hidden)
    with torch.no_grad():
        ref_lc = O.resampler(sd, "model.sampler.post_qformer.", feats[1:].float().cpu())
    assert rel(lc, ref_lc) < TOL
    lp = inner.mm_projector(lc)                           # :227 -> plain projection branch
    assert lp.shape == (4, 144, cfg.hidden_size)
    with torch.no_grad():
        assert rel(lp, O.projection(sd, lc.float().cpu())) < TOL
    # router through the module signature (local_f, text_embedding, attn_mask)  :248
    te, tm = model.get_pure_text_embedding(ids.cuda(), mask.cuda())
    with torch.no_grad():
        ref_te, ref_tm = O.pure_text_embedding(sd["model.embed_tokens.weight"], ids[0], mask[0])
    assert torch.equal(tm[0].cpu(), ref_tm)
    assert rel(te[0], ref_te) < 1e-6 + 4e-3
    lm = O.spatial_merge(lp.float().cpu(), (2, 2), 12).to(torch.bfloat16).cuda()
    kept = inner.sampler(lm, te[0], tm[0])
    with torch.no_grad():
        pr = O.router_probs(lm.float().cpu(), te[0].float().cpu(), tm[0].cpu())
    expect = O.top_p_select(pr, cfg.mm_resampler_topp)
    assert abs(kept.shape[0] - expect.numel()) <= 3      # bf16 vs fp32 probabilities may move the cut by a token
    assert kept.shape[1] == cfg.hidden_size


def test_prepare_inputs_and_forward(setup):
    cfg, model, sd, px, ids, mask, ora, gold = setup
    labels = ids.clone()
    out = model.prepare_inputs_labels_for_multimodal(ids.cuda(), None, mask.cuda(), None, labels.cuda(),
                                                     px.cuda().to(torch.bfloat16), image_sizes=[(672, 672)] * 2)
    none_ids, pos, am, pkv, embeds, new_labels = out
    assert none_ids is None and pos is None and pkv is None
    assert am.dtype == mask.dtype and embeds.dim() == 3 and new_labels.shape == am.shape
    lens = am.sum(1).tolist()
    # lengths may differ from the fp32 golden by the 1-3 tokens the bf16 router moves
    for L, Lg in zip(lens, gold["lengths"].tolist()):
        assert abs(L - Lg) <= 3
    assert (new_labels[:, :5].cpu() == labels[:, :5]).all()
    assert (new_labels[0, 5:5 + 577] == -100).all()
    res = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().to(torch.bfloat16),
                image_sizes=[(672, 672)] * 2, labels=labels.cuda())
    assert res.logits.shape[0] == 2 and res.logits.shape[2] == cfg.vocab_size
    assert torch.isfinite(res.loss)
    # same numbers as the engine fast path
    eng = model._engine()
    fast = eng.prefill(px, ids, mask, grids=[(2, 2)] * 2)
    for b in range(2):
        assert rel(res.logits[b, lens[b] - 1], fast.logits_last[b]) < 1e-2
    # and close to the reference's fp32 logits when the kept sets coincide
    if lens == gold["lengths"].tolist():
        last = torch.stack([res.logits[b, lens[b] - 1] for b in range(2)])
        assert rel(last, gold["logits_last"]) < 2e-2


def test_generate_and_encode_images(setup):
    cfg, model, sd, px, ids, mask, ora, gold = setup
    feats, split = model.encode_images(px.flatten(0, 1).cuda().to(torch.bfloat16), ids.cuda(), [5, 5], mask.cuda(),
                                       image_sizes=[(672, 672)] * 2)
    assert split == [5, 5] and len(feats) == 2
    for b in range(2):
        assert feats[b].dim() == 3 and feats[b].shape[0] == 1 and feats[b].shape[2] == cfg.hidden_size
        assert abs(feats[b].shape[1] - ora["feats"][b].shape[0]) <= 3
        assert rel(feats[b][0, :577], ora["feats"][b][:577]) < TOL
    toks = model.generate(ids[:1].cuda(), images=px[:1].cuda().to(torch.bfloat16), image_sizes=[(672, 672)],
                          attention_mask=mask[:1].cuda(), max_new_tokens=3, do_sample=False)
    assert toks.shape == (1, 3)
    # greedy first token == argmax of the prefill logits
    fast = model._engine().prefill(px[:1], ids[:1], mask[:1], grids=[(2, 2)])
    assert int(toks[0, 0]) == int(fast.logits_last[0].argmax())


def test_images_mask_padding_is_ignored(setup):
    """The training collator pads the crop stack and passes images_mask (reference train.py:913-926,
    llava_arch.py:228-231,299-302): padded crops must not change the result."""
    cfg, model, sd, px, ids, mask, ora, gold = setup
    B = px.shape[0]
    pad = torch.zeros(B, 2, *px.shape[2:])
    px_pad = torch.cat([px, pad], 1).cuda().to(torch.bfloat16)
    images_mask = torch.cat([torch.ones(B, px.shape[1]), torch.zeros(B, 2)], 1).bool().cuda()
    a = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().to(torch.bfloat16),
              image_sizes=[(672, 672)] * B)
    b = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px_pad, images_mask=images_mask,
              image_sizes=[(672, 672)] * B)
    assert a.logits.shape == b.logits.shape
    assert torch.equal(a.logits, b.logits)


def test_half_model_runs_the_float16_build():
    """model.half() (the reference's inference dtype, llava/model/builder.py:43) selects libslime_b200_fp16.so;
    the logits agree with the bf16 run of the same model to bf16 accuracy and are fp16-finite."""
    from slime_b200.synth import synth_inputs, synth_state_dict
    from tests.test_shims_cpu import make_model

    cfg, model = make_model("tiny")
    sd = synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device="cuda", dtype=torch.float16).eval()
    px, ids, mask = synth_inputs(cfg, 2, 5, 24, image_pos=5, ragged=True)
    res = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().half(), image_sizes=[(672, 672)] * 2)
    eng = model._engine()
    assert eng.dtype == torch.float16 and eng.lib.slime_elem_dtype() == 2
    assert torch.isfinite(res.logits.float()).all()
    fast = eng.prefill(px, ids, mask, grids=[(2, 2)] * 2)
    lens = fast.lengths
    for b in range(2):
        assert rel(res.logits[b, lens[b] - 1], fast.logits_last[b]) < 2e-3
    model = model.to(dtype=torch.bfloat16)          # back to bf16: the binding re-creates a bf16 engine
    res2 = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().bfloat16(),
                 image_sizes=[(672, 672)] * 2)
    assert model._engine().dtype == torch.bfloat16
    assert res2.logits.shape[2] == cfg.vocab_size


def test_forward_use_cache_then_single_token_steps(setup):
    """The HF generation loop's contract (reference llava_llama.py:57-104,146-157; llava_arch.py:279 is the early-out of
    the later calls): forward(..., use_cache=True) returns past_key_values; handing it back with ONE new token per
    sequence runs one decode step.  Checked against a cache-less forward of the grown sequence."""
    from slime_b200.model.language_model.llava_llama import SlimeKVCache

    cfg, model, sd, px, ids, mask, ora, gold = setup
    B = px.shape[0]
    images = px.cuda().to(torch.bfloat16)
    plain = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=images, image_sizes=[(672, 672)] * B)
    assert plain.past_key_values is None
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=images, image_sizes=[(672, 672)] * B,
                use_cache=True)
    pkv = out.past_key_values
    assert isinstance(pkv, SlimeKVCache) and pkv.batch == B
    assert torch.equal(out.logits, plain.logits)                      # attaching the cache does not change the prefill
    lens = pkv.lens.cpu().tolist()
    assert pkv.get_seq_length() == max(lens)
    # the spliced embeddings of the prompt (to build the grown sequences for the cache-less check)
    prep = model.prepare_inputs_labels_for_multimodal(ids.cuda(), None, mask.cuda(), None, None, images,
                                                      image_sizes=[(672, 672)] * B)
    emb = prep[4]
    seqs = [emb[b, :lens[b]] for b in range(B)]
    table = model.get_input_embeddings().weight
    logits = torch.stack([out.logits[b, lens[b] - 1] for b in range(B)])
    for step in range(3):
        nxt = logits.float().argmax(-1)
        res = model(input_ids=nxt[:, None], past_key_values=pkv, use_cache=True)
        assert res.past_key_values is pkv and res.logits.shape == (B, 1, cfg.vocab_size)
        assert pkv.lens.cpu().tolist() == [n + step + 1 for n in lens]
        seqs = [torch.cat([seqs[b], table[nxt[b]][None]]) for b in range(B)]
        Lm = max(s.shape[0] for s in seqs)
        padded = torch.stack([torch.cat([s, s.new_zeros(Lm - s.shape[0], s.shape[1])]) for s in seqs])
        am = torch.stack([torch.arange(Lm, device="cuda") < s.shape[0] for s in seqs])
        ref = model(inputs_embeds=padded, attention_mask=am)
        ref_last = torch.stack([ref.logits[b, seqs[b].shape[0] - 1] for b in range(B)])
        logits = res.logits[:, 0]
        assert rel(logits, ref_last) < 1e-2
    with pytest.raises(NotImplementedError, match="one new token"):
        model(input_ids=torch.zeros(B, 2, dtype=torch.long, device="cuda"), past_key_values=pkv)
    with pytest.raises(TypeError):
        model(input_ids=nxt[:, None], past_key_values=((None, None),))
    # growth by re-allocation keeps the cached prefix
    before = pkv.cache[:, :, :, :pkv.get_seq_length()].clone()
    have = pkv.cache.shape[3]
    pkv.reserve(have)  # force a bigger tensor
    assert pkv.cache.shape[3] > have and torch.equal(pkv.cache[:, :, :, :before.shape[3]], before)
    res = model(input_ids=nxt[:, None], past_key_values=pkv, use_cache=True)
    assert torch.isfinite(res.logits.float()).all()


def test_sampling_follows_the_global_rng(setup):
    """generate(do_sample=True) must not return the same sample on every call; torch.manual_seed pins it."""
    cfg, model, sd, px, ids, mask, ora, gold = setup
    kw = dict(images=px[:1].cuda().to(torch.bfloat16), image_sizes=[(672, 672)], attention_mask=mask[:1].cuda(),
              max_new_tokens=8, do_sample=True, temperature=1.5)
    torch.manual_seed(1)
    a = model.generate(ids[:1].cuda(), **kw)
    b = model.generate(ids[:1].cuda(), **kw)
    torch.manual_seed(1)
    c = model.generate(ids[:1].cuda(), **kw)
    assert torch.equal(a, c)
    assert not torch.equal(a, b)
    d = model.generate(ids[:1].cuda(), seed=5, **kw)
    e = model.generate(ids[:1].cuda(), seed=5, **kw)
    assert torch.equal(d, e)


def test_checkpoint_directory_to_prefill_matches_oracle():
    """SURVEY 8f.3 on the GPU: a synthetic HF checkpoint directory (sharded safetensors + index + config.json + a CLIP
    directory) is loaded with load_pretrained_model (the reference's entry point, llava/model/builder.py:26; fp16 like
    the reference, :43) and the loaded model's prefill matches the fp32 oracle on the same weights."""
    import json
    import tempfile

    from safetensors.torch import save_file

    from oracle import slime_oracle as O
    from slime_b200.checkpoint import load_pretrained_model
    from slime_b200.config import preset
    from slime_b200.synth import synth_inputs, synth_state_dict
    from tests.test_checkpoint_cpu import _write_clip, _write_config

    cfg = preset("tiny")
    sd = synth_state_dict(cfg)
    sd16 = {k: v.to(torch.float16) for k, v in sd.items()}
    clip_dir = _write_clip(cfg, sd16)
    llm = {k: v for k, v in sd16.items() if not k.startswith("model.vision_tower.")}
    d = tempfile.mkdtemp(prefix="ckpt_gpu_")
    _write_config(d, cfg, clip_dir)
    keys = sorted(llm)
    shards = {"model-00001-of-00002.safetensors": keys[::2], "model-00002-of-00002.safetensors": keys[1::2]}
    wm = {}
    for fn, ks in shards.items():
        save_file({k: llm[k].contiguous() for k in ks}, os.path.join(d, fn))
        wm.update({k: fn for k in ks})
    with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
        json.dump({"weight_map": wm}, f)
    tok, model, proc, ctx_len = load_pretrained_model(d, None, "slime-tiny", device="cuda")
    assert next(model.parameters()).dtype == torch.float16 and model.get_vision_tower().weights_source == "directory"
    px, ids, mask = synth_inputs(cfg, 2, 5, 24, image_pos=5, ragged=True)
    sdf = {k: v.float() for k, v in sd16.items()}  # the oracle on exactly the stored (fp16-rounded) weights
    with torch.no_grad():
        ora = O.prefill(sdf, cfg, px, ids, mask, [(2, 2)] * 2)
    eng = model._engine()
    assert eng.dtype == torch.float16
    res = eng.prefill(px, ids, mask, grids=[(2, 2)] * 2, forced_selection=ora["sel"])
    assert res.lengths == ora["lengths"]
    ref = torch.stack([lg[-1] for lg in ora["logits"]])
    e = rel(res.logits_last, ref)
    print(f"checkpoint dir -> load_pretrained_model -> prefill (fp16): last-token logits rel-L2 vs fp32 oracle {e:.3e}")
    assert e < 3e-3
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().half(), image_sizes=[(672, 672)] * 2)
    assert torch.isfinite(out.logits.float()).all()


def test_full_model_with_qformer_router_runs():
    """the root binding must register the 'router' weight group when mm_resampler_type == 'qformer'."""
    from slime_b200.synth import synth_inputs, synth_state_dict
    from tests.test_shims_cpu import make_model

    cfg, model = make_model("tiny", mm_resampler_type="qformer")
    model.load_state_dict(synth_state_dict(cfg), strict=True)
    model = model.to(device="cuda", dtype=torch.bfloat16).eval()
    px, ids, mask = synth_inputs(cfg, 2, 5, 24, image_pos=5, ragged=True)
    out = model(input_ids=ids.cuda(), attention_mask=mask.cuda(), images=px.cuda().to(torch.bfloat16),
                image_sizes=[(672, 672)] * 2)
    assert torch.isfinite(out.logits.float()).all() and out.logits.shape[0] == 2
