"""Parity at BASELINE.json's FULL model sizes on a B200.

(1) Full-size oracle check: the CPU oracle's code is plain PyTorch, so here it runs in fp32 ON THE GPU (TF32 off)
    as the checker for the real SliME-Vicuna-7B / SliME-Llama3-8B dimensions (CLIP-L/14-336, 32 decoder layers,
    5 crops of 336 px, T = 128 / 256) - stage outputs and last-token logits of the bf16 CUDA path against it,
    with the oracle's selected indices teacher-forced.
(2) Size-independent properties that need no oracle (also run on the 13B dimensions of config 5):
      * batch invariance   - a sample computed alone and inside a batch gives BIT-IDENTICAL logits
                             (every output element accumulates over K in a fixed order, packed rows never mix)
      * permutation        - permuting the samples permutes the outputs bit-exactly
      * padding invariance - appending masked prompt slots changes nothing (mask strip of the splice)
      * causality          - changing prompt tokens AFTER a position leaves earlier positions' logits bit-identical
      * length accounting  - L_i = (T_i - 1) + 576 + 1 + K_i and cu_seqlens is its prefix sum
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


_cache = {}


def build(name, layers=None, dtype=torch.bfloat16):
    """Engine with on-GPU synthetic weights of the named architecture (+ an accessor for the same tensors)."""
    key = (name, layers, dtype)
    if key in _cache:
        return _cache[key]
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_tensor, weight_specs

    _cache.clear()  # one big model resident at a time
    torch.cuda.empty_cache()
    cfg = preset(name) if layers is None else preset(name, num_hidden_layers=layers)
    specs = {n: (s, k) for n, s, k in weight_specs(cfg)}
    dev = torch.device("cuda", 0)

    def get(n):
        return synth_tensor(n, specs[n][0], specs[n][1], 3407, device=dev, dtype=torch.bfloat16)

    eng = SlimeEngine(cfg, 0, max_pos=4096, dtype=dtype)
    eng.load_weights(get)
    _cache[key] = (cfg, eng, get, specs)
    return _cache[key]


class LazyFp32Dict(dict):
    """Reference-keyed fp32 view of the synthetic weights, generated on demand (keeps 32 GB out of memory)."""

    def __init__(self, get):
        super().__init__()
        self._get = get

    def __missing__(self, k):
        return self._get(k).float()


class LazyCastDict(dict):
    """The same weights cast to a 16-bit dtype on demand: the oracle run "in bf16" is the reference's own low-precision
    execution (plain torch ops on 16-bit tensors: cuBLAS GEMMs with fp32 accumulation, every intermediate rounded)."""

    def __init__(self, get, dtype):
        super().__init__()
        self._get, self._dtype = get, dtype

    def __missing__(self, k):
        return self._get(k).to(self._dtype)


# Bounds (rel-L2 vs the fp32 oracle on the GPU).  north_star asks for "1e-3 relative bf16/fp16 tolerance"; one bf16
# rounding alone is 1.7e-3, so the bf16 path is judged against the FLOOR instead: the oracle itself executed in bf16
# (test_fullsize_floor_*: ours <= 1.2 x that, per stage and on the logits), and the absolute bounds below are the
# values measured on B200 (profiles/r02_fullsize_parity.txt) x 1.3, so a regression of a third already fails.  The fp16
# build (the reference's inference dtype, llava/model/builder.py:43) is the one that reaches the 1e-3 order.
MEASURED = {  # dtype -> stage -> measured rel-L2 (32 layers, real dimensions)
    torch.bfloat16: dict(vit=9.41e-3, glob=1.42e-2, local=4.21e-3, probs=1.41e-3, logits=1.375e-2),
    torch.float16: dict(vit=1.17e-3, glob=1.50e-3, local=4.9e-4, probs=2.0e-4, logits=1.69e-3),
}
TOL = {dt: {k: 1.3 * v for k, v in m.items()} for dt, m in MEASURED.items()}


def _run_case(name, T, dtype, B, n, layers=None, flat=False, floor=False, seed=11):
    """One batch at real dimensions through the CUDA path and through the fp32 oracle on the GPU (teacher-forced
    selection); returns the per-stage errors (and, with floor=True, the errors of the oracle executed in `dtype`)."""
    from oracle import slime_oracle as O
    from slime_b200.synth import grid_for_crops, synth_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, eng, get, specs = build(name, layers, dtype=dtype)
    if flat:
        cfg = cfg.replace(mm_patch_merge_type="flat")
        eng.cfg.mm_patch_merge_type = "flat"
    try:
        px, ids, mask = synth_inputs(cfg, B, n, T, seed=seed, ragged=True)
        px, ids, mask = px.cuda(), ids.cuda(), mask.cuda()
        sd = LazyFp32Dict(get)
        grids = None if flat else [grid_for_crops(n - 1)] * B
        out = {}
        with torch.no_grad():
            enc = O.encode_images(sd, cfg, px, ids, mask, grids)           # fp32 oracle, on the GPU
            res = eng.prefill(px, ids, mask, grids=grids, forced_selection=enc["sel"], keep_stages=True)
            out["vit"] = rel(res.stages["vit"], torch.cat(enc["vit"]))
            out["glob"] = rel(res.stages["glob"], torch.stack(enc["glob"]))
            out["local"] = rel(res.stages["local_m"], torch.stack(enc["local_m"]))
            # router probabilities on the CUDA path's own features, then the exact selection rule on them
            r2 = eng.prefill(px, ids, mask, grids=grids, want_probs=True, want_last=False, run_decoder=False)
            out["probs"] = max(rel(r2.probs[b], enc["probs"][b]) for b in range(B))
            for b in range(B):
                expect = O.top_p_select(r2.probs[b].cpu(), cfg.mm_resampler_topp)
                k = int(r2.sel_count[b])
                assert r2.sel_idx[b, :k].cpu().tolist() == expect.tolist()
            # splice (exact lengths) + decoder against the oracle
            emb, am, pid, lab, lens = O.splice(sd["model.embed_tokens.weight"].cpu(), ids.cpu(), mask.cpu(), None,
                                               [f.cpu() for f in enc["feats"]])
            assert res.lengths == lens
            last = torch.stack([O.llama_last_logits(sd, cfg, emb[b, :L].cuda()) for b, L in enumerate(lens)])
            out["logits"] = rel(res.logits_last, last)
            if floor:
                # the oracle executed in the 16-bit dtype (same inputs, same teacher-forced selection)
                sd16 = LazyCastDict(get, dtype)
                enc16 = O.encode_images(sd16, cfg, px.to(dtype), ids, mask, grids, forced_selection=enc["sel"])
                fl = dict(vit=rel(torch.cat(enc16["vit"]), torch.cat(enc["vit"])),
                          glob=rel(torch.stack(enc16["glob"]), torch.stack(enc["glob"])),
                          local=rel(torch.stack(enc16["local_m"]), torch.stack(enc["local_m"])),
                          probs=max(rel(enc16["probs"][b], enc["probs"][b]) for b in range(B)))
                emb16 = O.splice(sd16["model.embed_tokens.weight"], ids, mask, None, enc16["feats"])[0]
                last16 = torch.stack([O.llama_last_logits(sd16, cfg, emb16[b, :L]) for b, L in enumerate(lens)])
                fl["logits"] = rel(last16, last)
                out["floor"] = fl
        return out
    finally:
        eng.cfg.mm_patch_merge_type = "spatial"


def _report(tag, out):
    line = "  ".join(f"{k} {v:.3e}" for k, v in out.items() if k != "floor")
    print(f"[{tag}] rel-L2 vs fp32 oracle: {line}")
    if "floor" in out:
        print(f"[{tag}] oracle executed in 16 bit : " + "  ".join(f"{k} {v:.3e}" for k, v in out["floor"].items()))


@pytest.mark.parametrize("name,T,dtype", [("vicuna-7b", 128, torch.bfloat16), ("llama3-8b", 256, torch.bfloat16),
                                          ("llama3-8b", 256, torch.float16)],
                         ids=["vicuna-7b-bf16", "llama3-8b-bf16", "llama3-8b-fp16"])
def test_fullsize_against_fp32_oracle_and_floor(name, T, dtype):
    """Config 2 (Vicuna-7B, T 128) and the headline (Llama3-8B, T 256): 5 crops, 2x2 spatial grid, all 32 layers.
    (1) absolute bounds = measured x 1.3;  (2) FLOOR: the CUDA path may not be further from fp32 than 1.2 x the
    oracle itself executed in the same 16-bit dtype (= the reference's own bf16 / fp16 run on this input)."""
    out = _run_case(name, T, dtype, B=2, n=5, floor=True)
    _report(f"{name} {dtype} 5 crops", out)
    tol = TOL[dtype]
    for k in ("vit", "glob", "local", "probs", "logits"):
        assert out[k] < tol[k], f"{k}: {out[k]:.3e} above measured x 1.3 = {tol[k]:.3e}"
        # 1.2 x the floor, with a small absolute allowance where the floor itself is at the 1e-4 level
        assert out[k] <= 1.2 * out["floor"][k] + 2e-4, f"{k}: {out[k]:.3e} vs 16-bit oracle floor {out['floor'][k]:.3e}"


@pytest.mark.parametrize("name,n,T,B,layers", [
    ("llama3-8b", 10, 256, 8, 4),    # config 3: 1008 px = 10 crops, batch 8 ('flat' merge: 3x3 is not a reference grid)
    ("llama3-8b", 8, 256, 4, 4),     # config 4: 8 frames = 8 crops (crop 0 global + 7 local, flat), 4 samples per GPU
    ("vicuna-13b", 17, 512, 2, 4),   # config 5: Vicuna-13B dimensions (H 5120, 40 heads), 17 crops, T = 512
], ids=["config3-10crops", "config4-8crops", "config5-13b-17crops"])
def test_baseline_config_shapes_against_fp32_oracle(name, n, T, B, layers):
    """BASELINE.json configs 3 / 4 / 5 at their crop counts, prompt lengths and model dimensions (decoder sliced to
    4 layers so the fp32 oracle stays in seconds): every stage and the last-token logits against the fp32 oracle
    (reference llava_arch.py:233-234 flat merge; eval/video/llava_arch.py:240 for the 8-crop shape)."""
    out = _run_case(name, T, torch.bfloat16, B=B, n=n, layers=layers, flat=True, floor=True, seed=7)
    _report(f"{name} {n} crops T={T} B={B} ({layers} layers)", out)
    tol = TOL[torch.bfloat16]
    for k in ("vit", "glob", "local", "probs"):
        assert out[k] < tol[k], f"{k}: {out[k]:.3e} above {tol[k]:.3e}"
        assert out[k] <= 1.2 * out["floor"][k] + 2e-4, f"{k}: {out[k]:.3e} vs floor {out['floor'][k]:.3e}"
    # 4 decoder layers accumulate less rounding than 32: bound by the 32-layer figure and by the floor
    assert out["logits"] < tol["logits"]
    assert out["logits"] <= 1.2 * out["floor"]["logits"] + 2e-4


@pytest.mark.parametrize("B", [1, 16])
def test_fullsize_decode_parity(B):
    """Decode step at SliME-Llama3-8B dimensions (32 layers, GQA 32/8, V = 128256, ~1380-token context): every step's
    logits against (a) a fresh packed prefill of the grown sequence on the CUDA path and (b) the fp32 oracle's
    last-token logits for the grown sequence (reference llava_llama.py:139 -> HF generation loop)."""
    import numpy as np

    from oracle import slime_oracle as O
    from slime_b200.synth import synth_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, eng, get, specs = build("llama3-8b")
    sd = LazyFp32Dict(get)
    px, ids, mask = synth_inputs(cfg, B, 5, 256, seed=21, ragged=True)
    grids = [(2, 2)] * B
    steps = 3
    base = eng.prefill(px, ids, mask, grids=grids, keep_stages=True)
    cu = base.cu_seqlens.cpu().tolist()
    seqs = [base.embeds[cu[b]:cu[b + 1]].clone() for b in range(B)]
    table = eng.weights["llm.embed"]
    forced = [base.sel_idx[b, :int(base.sel_count[b])] for b in range(B)]
    eng.attach_kv_cache(B, max(base.lengths) + steps + 2)
    try:
        res = eng.prefill(px, ids, mask, grids=grids, forced_selection=forced)
        assert torch.equal(res.logits_last, base.logits_last)
        lens_d = torch.tensor(res.lengths, dtype=torch.int32, device="cuda")
        logits = res.logits_last
        worst_re, worst_or = 0.0, 0.0
        for s in range(steps):
            nxt = logits.argmax(-1)
            seqs = [torch.cat([seqs[b], table[nxt[b]][None]]) for b in range(B)]
            logits = eng.decode_step(table[nxt], lens_d)
            lens_d = lens_d + 1
            # (a) re-prefill of the grown sequences (cache detached for this call: it must not be overwritten)
            lens = [x.shape[0] for x in seqs]
            rows = torch.cat(seqs).contiguous()
            cu_t = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device="cuda")
            pos = torch.cat([torch.arange(L) for L in lens]).to(device="cuda", dtype=torch.int32)
            cache = eng._kv_cache
            eng.detach_kv_cache()
            ref, _, _ = eng.decoder_prefill(rows, cu_t, pos, lens)
            eng._check(eng.lib.slime_decoder_set_kv_cache(eng._ctx, cache.data_ptr(), B, cache.shape[3]), "set_kv_cache")
            eng._kv_cache = cache
            worst_re = max(worst_re, rel(logits, ref))
            # (b) fp32 oracle on the grown sequence of the first and the last sample
            with torch.no_grad():
                for b in sorted({0, B - 1}):
                    lg = O.llama_last_logits(sd, cfg, seqs[b].float())
                    worst_or = max(worst_or, rel(logits[b], lg))
        print(f"[llama3-8b decode B={B}] {steps} steps: rel-L2 vs re-prefill {worst_re:.3e}, vs fp32 oracle {worst_or:.3e}")
        # two bf16 executions with different rounding points (the decode kernels keep the residual + RMSNorm in fp32, the
        # prefill rounds the stream to bf16 in between) sit ~1.4e-2 from fp32 each and ~1.5e-2 from each other after 32
        # layers (measured 1.54e-2; bound = x 1.3); against fp32 the decode step obeys the same bound as the prefill
        assert worst_re < 2.0e-2, worst_re
        assert worst_or < TOL[torch.bfloat16]["logits"], worst_or
    finally:
        eng.detach_kv_cache()


@pytest.mark.parametrize("name,n_crops,T,layers", [
    ("llama3-8b", 5, 256, None),     # headline shape
    ("llama3-8b", 10, 256, 8),       # config 3 shape: 1008 px = 10 crops ('flat' merge: 3x3 is not a reference grid)
    ("vicuna-13b", 17, 512, 4),      # config 5 dimensions (H = 5120, 40 heads), 17 crops, T = 512, decoder sliced
])
def test_fullsize_invariants(name, n_crops, T, layers):
    from slime_b200.synth import grid_for_crops, synth_inputs

    cfg, eng, get, specs = build(name, layers)
    flat = n_crops in (10, 17)
    if flat:
        eng.cfg.mm_patch_merge_type = "flat"
    try:
        B = 3
        px, ids, mask = synth_inputs(cfg, B, n_crops, T, seed=5, ragged=True)
        grids = None if flat else [grid_for_crops(n_crops - 1)] * B
        full = eng.prefill(px, ids, mask, grids=grids, want_all_logits=True)
        torch.cuda.synchronize()
        q = cfg.mm_resampler_dim
        # length accounting
        for b in range(B):
            t_valid = int(mask[b].sum())
            assert full.lengths[b] == (t_valid - 1) + 576 + 1 + int(full.sel_count[b])
            assert 1 <= int(full.sel_count[b]) <= (n_crops - 1) * q
        cu = full.cu_seqlens.cpu().tolist()
        assert cu == [0] + torch.tensor(full.lengths).cumsum(0).tolist()
        # batch invariance + permutation: bit-identical logits
        for b in range(B):
            one = eng.prefill(px[b:b + 1], ids[b:b + 1], mask[b:b + 1], grids=None if flat else grids[:1])
            assert one.lengths[0] == full.lengths[b]
            assert torch.equal(one.logits_last[0], full.logits_last[b]), f"sample {b}: batch of 1 differs from batch of {B}"
        perm = [2, 0, 1]
        pr = eng.prefill(px[perm], ids[perm], mask[perm], grids=grids)
        assert torch.equal(pr.logits_last, full.logits_last[perm])
        # padding invariance: extra masked prompt slots are stripped by the splice
        ids_p = torch.cat([ids, torch.full((B, 7), cfg.pad_token_id, dtype=ids.dtype)], 1)
        mask_p = torch.cat([mask, torch.zeros(B, 7, dtype=mask.dtype)], 1)
        pad = eng.prefill(px, ids_p, mask_p, grids=grids)
        assert pad.lengths == full.lengths
        assert torch.equal(pad.logits_last, full.logits_last)
        # causality: change the last 5 valid prompt tokens of sample 0 -> logits of earlier positions unchanged.
        # (tokens before the image also feed the router, so only post-image tokens far from it are changed and the
        #  router's selection is teacher-forced to the original one)
        ids_c = ids.clone()
        t0 = int(mask[0].sum())
        ids_c[0, t0 - 5:t0] = (ids_c[0, t0 - 5:t0] + 1) % (cfg.vocab_size - 1000) + 3
        forced = [full.sel_idx[b, :int(full.sel_count[b])] for b in range(B)]
        base = eng.prefill(px, ids, mask, grids=grids, forced_selection=forced, want_all_logits=True)
        chg = eng.prefill(px, ids_c, mask, grids=grids, forced_selection=forced, want_all_logits=True)
        L0 = full.lengths[0]
        assert torch.equal(base.logits_all[cu[0]:cu[0] + L0 - 5], chg.logits_all[cu[0]:cu[0] + L0 - 5])
        assert not torch.equal(base.logits_all[cu[0] + L0 - 1], chg.logits_all[cu[0] + L0 - 1])
        assert torch.equal(base.logits_all[cu[1]:], chg.logits_all[cu[1]:])
    finally:
        eng.cfg.mm_patch_merge_type = "spatial"
