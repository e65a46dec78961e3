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


# tolerances (rel-L2 vs the fp32 oracle): bf16 keeps 8 significand bits per rounding, fp16 11 -> 8x tighter.
# north_star asks for 1e-3 "bf16/fp16 tolerance": the fp16 build is what gets the 32-layer logits to that order;
# in bf16 the reference's own bf16 execution is ~1e-2 away from fp32 as well (tolerances).
TOL = {torch.bfloat16: dict(stage=2e-2, logits=5e-2), torch.float16: dict(stage=2.5e-3, logits=3e-3)}


@pytest.mark.parametrize("name,T,dtype", [("vicuna-7b", 128, torch.bfloat16), ("llama3-8b", 256, torch.bfloat16),
                                          ("llama3-8b", 256, torch.float16)],
                         ids=["vicuna-7b-bf16", "llama3-8b-bf16", "llama3-8b-fp16"])
def test_fullsize_against_fp32_oracle_on_gpu(name, T, dtype):
    from oracle import slime_oracle as O
    from slime_b200.synth import synth_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, eng, get, specs = build(name, dtype=dtype)
    tol = TOL[dtype]
    B, n = 2, 5
    px, ids, mask = synth_inputs(cfg, B, n, T, seed=11, ragged=True)
    px, ids, mask = px.cuda(), ids.cuda(), mask.cuda()
    sd = LazyFp32Dict(get)
    grids = [(2, 2)] * B
    with torch.no_grad():
        enc = O.encode_images(sd, cfg, px, ids, mask, grids)           # fp32 oracle, on the GPU
        res = eng.prefill(px, ids, mask, grids=grids, forced_selection=enc["sel"], keep_stages=True)
        e_vit = rel(res.stages["vit"], torch.cat(enc["vit"]))
        e_glob = rel(res.stages["glob"], torch.stack(enc["glob"]))
        e_loc = rel(res.stages["local_m"], torch.stack(enc["local_m"]))
        print(f"[{name} {dtype}] full-size rel-L2: vit {e_vit:.3e}  gated-global {e_glob:.3e}  local {e_loc:.3e}")
        assert e_vit < tol["stage"] and e_glob < tol["stage"] and e_loc < tol["stage"]
        # router probabilities on the CUDA path's own features, then the exact selection rule on them
        r2 = eng.prefill(px, ids, mask, grids=grids, want_probs=True, want_last=False, run_decoder=False)
        for b in range(B):
            e_p = rel(r2.probs[b], enc["probs"][b])
            assert e_p < tol["stage"], e_p
            expect = O.top_p_select(r2.probs[b].cpu(), cfg.mm_resampler_topp)
            k = int(r2.sel_count[b])
            assert r2.sel_idx[b, :k].cpu().tolist() == expect.tolist()
        # splice + decoder (32 layers) against the oracle
        emb, am, pid, lab, lens = O.splice(sd["model.embed_tokens.weight"].cpu(), ids.cpu(), mask.cpu(), None,
                                           [f.cpu() for f in enc["feats"]])
        assert res.lengths == lens
        # run the oracle decoder per sample on the GPU (fp32) and keep only the last row
        last = []
        for b, L in enumerate(lens):
            lg = _oracle_last_logits(O, sd, cfg, emb[b, :L].cuda())
            last.append(lg)
        e_log = rel(res.logits_last, torch.stack(last))
        print(f"[{name} {dtype}] full-size last-token logits rel-L2 vs fp32 oracle: {e_log:.3e}")
        assert e_log < tol["logits"]


def _oracle_last_logits(O, sd, cfg, x):
    """oracle.llama_prefill for one sequence, last-token logits only (the all-position lm_head would be
    L x 128256 fp32); identical arithmetic."""
    import math

    import torch.nn.functional as F

    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    L = x.shape[0]
    dev = x.device
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device=dev) / hd))
    ang = torch.arange(L, dtype=torch.float32, device=dev)[:, None] * inv[None]
    cos, sin = torch.cat([ang, ang], -1).cos()[None], torch.cat([ang, ang], -1).sin()[None]
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)  # noqa: E731
    causal = torch.ones(L, L, dtype=torch.bool, device=dev).tril()
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        h = O.rms_norm(x, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = (h @ sd[p + "self_attn.q_proj.weight"].t()).view(L, nh, hd).transpose(0, 1)
        k = (h @ sd[p + "self_attn.k_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        v = (h @ sd[p + "self_attn.v_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        k, v = k.repeat_interleave(nh // nkv, 0), v.repeat_interleave(nh // nkv, 0)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        a = torch.softmax(s.masked_fill(~causal, float("-inf")), -1) @ v
        x = x + a.transpose(0, 1).reshape(L, nh * hd) @ sd[p + "self_attn.o_proj.weight"].t()
        h = O.rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = F.silu(h @ sd[p + "mlp.gate_proj.weight"].t()) * (h @ sd[p + "mlp.up_proj.weight"].t())
        x = x + g @ sd[p + "mlp.down_proj.weight"].t()
    x = O.rms_norm(x[-1:], sd["model.norm.weight"], cfg.rms_norm_eps)
    return (x @ sd["lm_head.weight"].t())[0]


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
