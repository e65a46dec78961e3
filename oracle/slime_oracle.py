"""TEST INFRASTRUCTURE - CPU restatement (plain PyTorch, fp32) of the reference algorithm for the
SliME prefill path.  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg as the CHECKER; the product path (slime_b200/) never touches it.

Pinned: tests/test_oracle_cpu.py checks every function here against tests/golden/*.npz, which were
produced by running the UNMODIFIED reference (/root/reference + transformers 5.5.0) on the same
synthetic weights (oracle/gen_golden.py).  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), so those generated fixtures are the pin; tests/test_oracle_live_reference_cpu.py additionally runs the
reference LIVE next to this file on randomised inputs wherever /root/reference exists (integer outputs exact).

Each function cites the reference lines it restates.  "HF:" = transformers 5.5.0
(site-packages/transformers/), the third-party dependency that holds the CLIP / Llama arithmetic
(reference pins transformers==4.37.2, pyproject.toml:16; not vendored under /root/reference).
Inputs: `sd` is a state dict keyed like the reference model (slime_b200/synth.py:weight_specs).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
CLIP = "model.vision_tower.vision_tower.vision_model."


# ------------------------------------------------------------------------------------------
# CLIP vision tower
# ------------------------------------------------------------------------------------------
def clip_tower(sd, cfg, pixels: torch.Tensor) -> torch.Tensor:
    """CLIPVisionTower.forward + feature_select (reference multimodal_encoder/clip_encoder.py:36-58):
    hidden_states[mm_vision_select_layer] of HF CLIPVisionModel without the CLS token.
    HF:models/clip/modeling_clip.py:202-218 (embeddings), :677 (pre_layrnorm), :371-382 (layer),
    :310-334 (attention), :348-350 (quick_gelu MLP).  pixels [N,3,S,S] -> [N, P, D]."""
    D, nh = cfg.vit_hidden, cfg.vit_heads
    hd = D // nh
    x = F.conv2d(pixels, sd[CLIP + "embeddings.patch_embedding.weight"], stride=cfg.vit_patch)  # no bias
    x = x.flatten(2).transpose(1, 2)  # [N, P, D], raster order
    cls = sd[CLIP + "embeddings.class_embedding"].expand(x.shape[0], 1, D)
    x = torch.cat([cls, x], dim=1) + sd[CLIP + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (D,), sd[CLIP + "pre_layrnorm.weight"], sd[CLIP + "pre_layrnorm.bias"], cfg.vit_ln_eps)
    for l in range(cfg.vit_layers_used):
        p = f"{CLIP}encoder.layers.{l}."
        h = F.layer_norm(x, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], cfg.vit_ln_eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        N, S, _ = q.shape
        q, k, v = (t.view(N, S, nh, hd).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1) @ v
        a = a.transpose(1, 2).reshape(N, S, D)
        x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], cfg.vit_ln_eps)
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x[:, 1:]


# ------------------------------------------------------------------------------------------
# Resampler, projector
# ------------------------------------------------------------------------------------------
def resized_pos(pos: torch.Tensor, tgt: int) -> torch.Tensor:
    """get_abs_pos (reference multimodal_resampler/sampler.py:27-36): bicubic resize of the [g*g, D]
    table to [tgt*tgt, D]."""
    src = int(math.sqrt(pos.shape[0]))
    out = F.interpolate(pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2), size=(tgt, tgt), mode="bicubic",
                        align_corners=False)
    return out.permute(0, 2, 3, 1).flatten(0, 2).to(pos.dtype)


def resampler(sd, prefix: str, x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Resampler.forward (reference multimodal_resampler/sampler.py:140-170): one cross-attention layer,
    learned queries + fixed 2-D sincos positions; nn.MultiheadAttention math (packed in_proj, heads of
    128).  x [n, 576, D] -> [n, nq, D]."""
    D = x.shape[-1]
    nh = D // 128
    n, NK, _ = x.shape
    query, pos_q = sd[prefix + "query"], sd[prefix + "pos_embed"]
    nq = query.shape[0]
    pos_k = resized_pos(pos_q, int(math.sqrt(NK)))
    kv = F.layer_norm(x, (D,), sd[prefix + "ln_kv.weight"], sd[prefix + "ln_kv.bias"], eps)
    qin = F.layer_norm(query, (D,), sd[prefix + "ln_q.weight"], sd[prefix + "ln_q.bias"], eps) + pos_q
    W, B = sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"]
    q = F.linear(qin, W[:D], B[:D])                       # [nq, D], shared by every crop
    k = F.linear(kv + pos_k[None], W[D:2 * D], B[D:2 * D])  # [n, NK, D]
    v = F.linear(kv, W[2 * D:], B[2 * D:])
    q = q.view(1, nq, nh, 128).transpose(1, 2)
    k = k.view(n, NK, nh, 128).transpose(1, 2)
    v = v.view(n, NK, nh, 128).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(128.0), dim=-1) @ v
    a = a.transpose(1, 2).reshape(n, nq, D)
    o = F.linear(a, sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"])
    return F.layer_norm(o, (D,), sd[prefix + "ln_post.weight"], sd[prefix + "ln_post.bias"], eps)


def projection(sd, x: torch.Tensor) -> torch.Tensor:
    """GatedBlock.projection = Linear, GELU(erf), Linear (reference multimodal_projector/builder.py:53-57)."""
    p = "model.mm_projector.projection."
    return F.linear(F.gelu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])


def gated_projector(sd, x: torch.Tensor, learnable_gated: int = -1) -> torch.Tensor:
    """GatedBlock.forward, eval mode, on one global crop x [576, D] (reference
    multimodal_projector/builder.py:179-209 and noisy_top_k_gating :137-171 with k = 2 of 2 experts)."""
    e0 = projection(sd, x)
    e1 = projection(sd, resampler(sd, "model.mm_projector.attn.", x[None])[0])
    if learnable_gated >= 0:
        return (e0, e1)[learnable_gated]
    probs = torch.softmax(x @ sd["model.mm_projector.w_gate"].to(x.dtype), dim=1)
    top, idx = probs.topk(2, dim=1)
    gates = torch.zeros_like(probs).scatter(1, idx, top / (top.sum(1, keepdim=True) + 1e-6))
    return gates[:, 0:1] * e0 + gates[:, 1:2] * e1


# ------------------------------------------------------------------------------------------
# host integer math: crop grid
# ------------------------------------------------------------------------------------------
def best_resolution_uhd(size: Tuple[int, int], base: int = 336) -> Tuple[int, int]:
    """select_best_resolution_uhd (reference llava/mm_utils.py:41-97)."""
    ow, oh = size
    scale = math.ceil(ow * oh / (base * base))
    scale = 6 if scale > 6 else (2 if scale == 1 else scale)
    facts = lambda n: [(i, n // i) for i in range(1, n + 1) if n % i == 0]  # noqa: E731
    cands = facts(scale) + facts(scale + 1) if scale <= 2 else facts(scale - 1) + facts(scale) + facts(scale + 1)
    best, best_eff, best_waste = None, 0, float("inf")
    for ws, hs in cands:
        w, h = ws * base, hs * base
        s = min(w / ow, h / oh)
        eff = min(int(ow * s) * int(oh * s), ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (w, h), eff, waste
    return best


def grid_shape(image_size: Tuple[int, int], patch: int = 336) -> Tuple[int, int]:
    """get_anyres_image_grid_shape (reference llava/mm_utils.py:156-174): (num_patch_width, num_patch_height)."""
    w, h = best_resolution_uhd(image_size, patch)
    return w // patch, h // patch


def spatial_merge(local: torch.Tensor, grid_wh: Tuple[int, int], g: int) -> torch.Tensor:
    """Raster re-ordering of compressed local tokens over the whole image (reference llava_arch.py:240-244).
    local [n_l, g*g, H] -> [n_l*g*g, H]."""
    w, h = grid_wh
    t = local.view(h, w, g, g, -1).permute(0, 2, 1, 3, 4).contiguous()
    return t.flatten(0, 3)


# ------------------------------------------------------------------------------------------
# text-guided router
# ------------------------------------------------------------------------------------------
def pure_text_embedding(embed: torch.Tensor, ids: torch.Tensor, mask: torch.Tensor):
    """get_pure_text_embedding for one sample, right padding (reference llava_arch.py:162-210): placeholder
    slots removed, zero rows appended so the length stays T; mask zero on those rows."""
    keep = ids != IMAGE_TOKEN_INDEX
    n_img = int((~keep).sum())
    e = embed[ids[keep]]
    m = mask[keep]
    if n_img:
        e = torch.cat([e, e.new_zeros(n_img, e.shape[1])])
        m = torch.cat([m, m.new_zeros(n_img)])
    return e, m


def router_probs(local: torch.Tensor, text: torch.Tensor, mask: torch.Tensor, temp: float = 1.0) -> torch.Tensor:
    """TextGuidedRouterCosine + softmax (reference multimodal_resampler/builder.py:189-201, :248-258)."""
    sim = F.cosine_similarity(local.unsqueeze(1), text.unsqueeze(0), dim=-1)
    sim = sim.masked_fill((mask == 0).unsqueeze(0), 0.0).sum(-1)
    return torch.softmax(sim / temp, dim=-1)


def router_qformer(sd, local: torch.Tensor, text: torch.Tensor, mask: torch.Tensor, temp: float = 1.0,
                   prefix: str = "model.sampler.selector.") -> torch.Tensor:
    """TextGuidedRouterAttention.forward (reference multimodal_resampler/builder.py:148-160): the local tokens [N, H]
    cross-attend the prompt [T, H] (nn.MultiheadAttention, heads of 128, key_padding_mask = ~mask), ln_post, the
    Linear-ReLU-Linear `prob_proj`, softmax(logits / temp) over the N tokens.  `query` / `self_attn` are unused there."""
    H = local.shape[1]
    heads, hd = H // 128, 128
    ln = lambda x, n: F.layer_norm(x, (H,), sd[prefix + n + ".weight"], sd[prefix + n + ".bias"], 1e-5)  # noqa: E731
    x, t = ln(local, "ln_q"), ln(text, "ln_kv")
    w, b = sd[prefix + "cross_attn.in_proj_weight"], sd[prefix + "cross_attn.in_proj_bias"]
    q = (x @ w[:H].t() + b[:H]).view(-1, heads, hd).transpose(0, 1)
    k = (t @ w[H:2 * H].t() + b[H:2 * H]).view(-1, heads, hd).transpose(0, 1)
    v = (t @ w[2 * H:].t() + b[2 * H:]).view(-1, heads, hd).transpose(0, 1)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    sc = sc.masked_fill((mask == 0)[None, None, :], float("-inf"))
    o = (torch.softmax(sc, -1) @ v).transpose(0, 1).reshape(-1, H)
    o = o @ sd[prefix + "cross_attn.out_proj.weight"].t() + sd[prefix + "cross_attn.out_proj.bias"]
    o = ln(o, "ln_post")
    h = F.relu(o @ sd[prefix + "prob_proj.0.weight"].t() + sd[prefix + "prob_proj.0.bias"])
    logit = (h @ sd[prefix + "prob_proj.2.weight"].t() + sd[prefix + "prob_proj.2.bias"]).squeeze(-1)
    return torch.softmax(logit / temp, dim=-1)


def top_p_select(probs: torch.Tensor, top_p: float) -> torch.Tensor:
    """The selection rule of TextGuidedSampler.forward (reference multimodal_resampler/builder.py:259-273),
    with the tie-break the framework specifies (stable: lower index first among equal probabilities;
    torch.sort(descending) is unstable).  Returns ascending kept indices."""
    probs = probs.detach().cpu()
    order = sorted(range(probs.numel()), key=lambda i: (-float(probs[i]), i))
    sp = probs[order]
    cum = torch.cumsum(sp, dim=0)
    count = int((cum <= top_p).sum())
    keep = count + 1 if count < probs.numel() else probs.numel()
    return torch.tensor(sorted(order[:keep]), dtype=torch.long)


# ------------------------------------------------------------------------------------------
# splice
# ------------------------------------------------------------------------------------------
def splice(embed: torch.Tensor, ids: torch.Tensor, mask: torch.Tensor, labels: Optional[torch.Tensor],
           image_feats: Sequence[torch.Tensor], max_len: Optional[int] = None, left_pad: bool = False):
    """prepare_inputs_labels_for_multimodal after encode_images (reference llava_arch.py:361-459).
    ids/mask [B,T]; image_feats[b] [n_b, H].  Returns (inputs_embeds [B,Lmax,H], attention_mask bool [B,Lmax],
    position_ids [B,Lmax], labels [B,Lmax], lengths)."""
    B = ids.shape[0]
    if labels is None:
        labels = torch.full_like(ids, IGNORE_INDEX)
    seqs, labs = [], []
    for b in range(B):
        m = mask[b].bool()
        cid, clab = ids[b][m], labels[b][m]
        pos = (cid == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        assert len(pos) <= 1, "the SliME path pairs one image with each sample"
        if not pos:
            seqs.append(embed[cid])
            labs.append(clab)
            continue
        p = pos[0]
        seqs.append(torch.cat([embed[cid[:p]], image_feats[b], embed[cid[p + 1:]]]))
        labs.append(torch.cat([clab[:p], clab.new_full((image_feats[b].shape[0],), IGNORE_INDEX), clab[p + 1:]]))
    if max_len is not None:
        seqs = [s[:max_len] for s in seqs]
        labs = [l[:max_len] for l in labs]
    lens = [s.shape[0] for s in seqs]
    Lmax, H = max(lens), embed.shape[1]
    out = embed.new_zeros(B, Lmax, H)
    dev = embed.device
    am = torch.zeros(B, Lmax, dtype=torch.bool, device=dev)
    pid = torch.zeros(B, Lmax, dtype=torch.long, device=dev)
    lab = torch.full((B, Lmax), IGNORE_INDEX, dtype=torch.long, device=dev)
    for b, L in enumerate(lens):
        sl = slice(Lmax - L, Lmax) if left_pad else slice(0, L)
        out[b, sl] = seqs[b]
        am[b, sl] = True
        pid[b, sl] = torch.arange(L, device=dev)
        lab[b, sl] = labs[b]
    return out, am, pid, lab, lens


# ------------------------------------------------------------------------------------------
# Llama decoder
# ------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    """HF:models/llama/modeling_llama.py:62-67 - variance in fp32, cast back to the input dtype BEFORE the weight
    multiply (identical to the plain formula for fp32 inputs; matters when the oracle is run in bf16 as the
    "reference's own bf16 execution" of the full-size floor test)."""
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def llama_hidden(sd, cfg, x: torch.Tensor, layers: Optional[Sequence[int]] = None) -> torch.Tensor:
    """The decoder layer stack of LlamaModel.forward for ONE sequence x [L, H] (HF:models/llama/modeling_llama.py:
    :73-168 rotary, :187-288 attention, :171-184 MLP, :291-332 layer), before the final norm.  Runs on x's device
    in x's dtype (cos / sin are computed in fp32 and cast like HF does); `layers` selects which layers' weights to
    run (default: all cfg.num_hidden_layers) - the CPU baseline uses it to time every layer position while holding
    only a few layers' weights in RAM."""
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    L, dev = x.shape[0], x.device
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device=dev) / hd))
    ang = torch.arange(L, dtype=torch.float32, device=dev)[:, None] * inv[None]
    cos, sin = torch.cat([ang, ang], -1).cos()[None].to(x.dtype), torch.cat([ang, ang], -1).sin()[None].to(x.dtype)
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)  # noqa: E731
    causal = torch.ones(L, L, dtype=torch.bool, device=dev).tril()
    for l in (range(cfg.num_hidden_layers) if layers is None else layers):
        p = f"model.layers.{l}."
        h = rms_norm(x, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = (h @ sd[p + "self_attn.q_proj.weight"].t()).view(L, nh, hd).transpose(0, 1)
        k = (h @ sd[p + "self_attn.k_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        v = (h @ sd[p + "self_attn.v_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        k, v = k.repeat_interleave(nh // nkv, 0), v.repeat_interleave(nh // nkv, 0)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        a = torch.softmax(s.masked_fill(~causal, float("-inf")), -1) @ v
        x = x + a.transpose(0, 1).reshape(L, nh * hd) @ sd[p + "self_attn.o_proj.weight"].t()
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = F.silu(h @ sd[p + "mlp.gate_proj.weight"].t()) * (h @ sd[p + "mlp.up_proj.weight"].t())
        x = x + g @ sd[p + "mlp.down_proj.weight"].t()
    return x


def llama_last_logits(sd, cfg, x: torch.Tensor, layers: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Last-token logits [V] of one sequence x [L, H] (final norm + lm_head on the last row only: the all-position
    lm_head of a 128256-entry vocabulary would be L x V fp32); same arithmetic as llama_prefill."""
    h = llama_hidden(sd, cfg, x, layers)
    h = rms_norm(h[-1:], sd["model.norm.weight"], cfg.rms_norm_eps)
    return (h @ sd["lm_head.weight"].t())[0]


def llama_prefill(sd, cfg, embeds: torch.Tensor, lengths: Sequence[int]) -> List[torch.Tensor]:
    """LlamaForCausalLM.forward(inputs_embeds=...) (HF:models/llama/modeling_llama.py:355-507), run per
    sequence (causal attention never crosses samples, so this equals the padded+masked batch).
    embeds [B, Lmax, H] right-padded; returns per-sample logits [L_b, V]."""
    out = []
    for b, L in enumerate(lengths):
        x = llama_hidden(sd, cfg, embeds[b, :L])
        x = rms_norm(x, sd["model.norm.weight"], cfg.rms_norm_eps)
        out.append(x @ sd["lm_head.weight"].t())
    return out


# ------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------
def encode_images(sd, cfg, pixels: torch.Tensor, ids: torch.Tensor, mask: torch.Tensor,
                  grids: Sequence[Tuple[int, int]], forced_selection: Optional[Sequence[torch.Tensor]] = None):
    """encode_images, sampler branch (reference llava_arch.py:212-255).  pixels [B, n, 3, S, S].
    Returns a dict of per-stage tensors (lists over the batch)."""
    embed = sd["model.embed_tokens.weight"]
    g = cfg.resampler_grid
    res = dict(vit=[], glob=[], local_c=[], local_m=[], probs=[], sel=[], feats=[])
    for b in range(pixels.shape[0]):
        f = clip_tower(sd, cfg, pixels[b])
        res["vit"].append(f)
        parts = []
        if not cfg.use_local_only:
            gl = gated_projector(sd, f[0], cfg.mm_learnable_gated)
            res["glob"].append(gl)
            parts += [gl, embed[cfg.seperator][None]] if not cfg.use_global_only else [gl]
        if not cfg.use_global_only:
            if f.shape[0] > 1:
                lc = resampler(sd, "model.sampler.post_qformer.", f[1:])
                lp = projection(sd, lc)
                lm = spatial_merge(lp, grids[b], g) if cfg.mm_patch_merge_type == "spatial" else lp.flatten(0, 1)
                te, tm = pure_text_embedding(embed, ids[b], mask[b])
                if cfg.mm_resampler_type == "qformer":
                    # the sampler soft-maxes the router's soft-max once more (builder.py:160 and :258)
                    pr = torch.softmax(router_qformer(sd, lm, te, tm, cfg.mm_resampler_temp) / cfg.mm_resampler_temp, -1)
                else:
                    pr = router_probs(lm, te, tm, cfg.mm_resampler_temp)
                sel = forced_selection[b] if forced_selection is not None else top_p_select(pr, cfg.mm_resampler_topp)
                res["local_c"].append(lc)
                res["local_m"].append(lm)
                res["probs"].append(pr)
                res["sel"].append(sel)
                parts.append(lm[sel])
            else:
                H = embed.shape[1]
                res["local_c"].append(f.new_zeros(0, g * g, f.shape[-1]))
                res["local_m"].append(f.new_zeros(0, H))
                res["probs"].append(f.new_zeros(0))
                res["sel"].append(torch.zeros(0, dtype=torch.long))
                parts.append(f.new_zeros(0, H))
        res["feats"].append(torch.cat(parts, 0))
    return res


def prefill(sd, cfg, pixels, ids, mask, grids, labels=None, forced_selection=None):
    """LlavaLlamaForCausalLM.forward(images=...) end to end (reference llava_llama.py:57-104)."""
    enc = encode_images(sd, cfg, pixels, ids, mask, grids, forced_selection)
    emb, am, pid, lab, lens = splice(sd["model.embed_tokens.weight"], ids, mask, labels, enc["feats"],
                                     cfg.tokenizer_model_max_length, cfg.tokenizer_padding_side == "left")
    if cfg.tokenizer_padding_side == "left":
        packed = torch.stack([torch.cat([emb[b, emb.shape[1] - L:], emb.new_zeros(emb.shape[1] - L, emb.shape[2])])
                              for b, L in enumerate(lens)])
    else:
        packed = emb
    logits = llama_prefill(sd, cfg, packed, lens)
    enc.update(inputs_embeds=emb, attention_mask=am, position_ids=pid, labels=lab, lengths=lens, logits=logits)
    return enc
