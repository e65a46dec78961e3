"""SlimeEngine: the host side of the B200-native SliME prefill path.

Owns a `slime_ctx` (include/slime_b200.h), the packed bf16 weights (borrowed by the library), one
growable workspace tensor, and a small cache of spatial-merge row maps.  Every method is a thin
ctypes call into libslime_b200.so - PyTorch is only the allocator / stream provider here.

Stage methods mirror the reference functions they replace (file:line in each docstring); `prefill`
chains them for a whole batch:  pixels + prompt ids -> last-token logits, with the samples PACKED
(no padding rows ever reach a GEMM) and a single host synchronisation (the spliced lengths).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .config import IMAGE_TOKEN_INDEX, SlimeConfig
from .mm_utils import get_anyres_image_grid_shape
from .weights import ALL_GROUPS, pack_weights


FUSED_ROPE_DEFAULT = "1"  # RoPE in the QKV GEMM epilogue (SLIME_FUSED_ROPE=0: separate in-place pass)
NORM_FOLD_DEFAULT = "1"   # Llama RMSNorms folded into the projections that follow them (SLIME_NORM_FOLD=0: separate passes)


def _locked(fn):
    """Serialises every entry into the (non re-entrant) slime_ctx and its shared workspace."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self._lock, torch.cuda.device(self.device):
            return fn(self, *a, **kw)

    return wrapper


@dataclass
class PrefillResult:
    logits_last: Optional[torch.Tensor]      # [B, V] fp32: logits of the last real token of each sample
    logits_all: Optional[torch.Tensor]       # [total, V] bf16 packed rows (only if requested)
    cu_seqlens: torch.Tensor                 # [B+1] int32 (device)
    lengths: Optional[List[int]]             # spliced length of every sample (host); None after a no-host-sync prefill
                                             # until resolve_lengths() reads cu_seqlens back
    sel_idx: Optional[torch.Tensor]          # [B, n_per] int32 kept local-token indices (ascending)
    sel_count: Optional[torch.Tensor]        # [B] int32
    probs: Optional[torch.Tensor]            # [B, n_per] fp32 router probabilities (if requested)
    embeds: Optional[torch.Tensor]           # [total, H] bf16 packed spliced embeddings (if requested)
    stages: Optional[dict] = None            # per-stage tensors (if requested)

    def resolve_lengths(self) -> List[int]:
        """Host copy of the spliced lengths (a no-host-sync prefill leaves them on the device: this synchronises)."""
        if self.lengths is None:
            cu = self.cu_seqlens.cpu().tolist()
            self.lengths = [cu[i + 1] - cu[i] for i in range(len(cu) - 1)]
        return self.lengths

    @property
    def total_tokens(self) -> int:
        return int(sum(self.resolve_lengths()))


class SlimeEngine:
    def __init__(self, cfg: SlimeConfig, device: int | str | torch.device = 0, max_pos: Optional[int] = None,
                 dtype=torch.bfloat16, fused_rope: Optional[bool] = None, norm_fold: Optional[bool] = None):
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError("SlimeEngine needs a CUDA (B200) device: there is no CPU path")
        # 16-bit element type of every activation / weight: bfloat16 (default) or float16, the reference's
        # inference dtype (llava/model/builder.py:43) - each has its own build of the library
        self.dtype = L.torch_dtype(L.variant_of(dtype))
        self.lib = L.load(self.dtype)
        self._check = lambda rc, what="": L.check(rc, what, self.lib)
        self._lock = threading.RLock()  # a slime_ctx is not re-entrant (serve/model_worker.py runs generate on threads)
        self._ws: Optional[torch.Tensor] = None
        self._row_maps: Dict[tuple, torch.Tensor] = {}
        self.weights: Dict[str, torch.Tensor] = {}
        d = L.ModelDesc()
        d.vit_hidden, d.vit_layers_used, d.vit_heads, d.vit_mlp = cfg.vit_hidden, cfg.vit_layers_used, cfg.vit_heads, cfg.vit_mlp
        d.vit_image, d.vit_patch, d.vit_ln_eps = cfg.vit_image, cfg.vit_patch, cfg.vit_ln_eps
        d.rs_local_queries, d.rs_global_queries, d.rs_ln_eps = cfg.mm_resampler_dim, 576, 1e-6
        d.mm_learnable_gated = cfg.mm_learnable_gated
        d.hidden, d.layers, d.heads, d.kv_heads = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.num_key_value_heads
        d.head_dim, d.mlp, d.vocab = cfg.head_dim, cfg.intermediate_size, cfg.vocab_size
        d.rope_theta, d.rms_eps = cfg.rope_theta, cfg.rms_norm_eps
        d.max_pos = int(max_pos or cfg.max_position_embeddings)
        d.top_p, d.temp = cfg.mm_resampler_topp, cfg.mm_resampler_temp
        d.image_token, d.sep_token = IMAGE_TOKEN_INDEX, cfg.seperator
        d.max_len = int(cfg.tokenizer_model_max_length or 0)
        # RoPE inside the QKV GEMM's epilogue (q / k weight rows interleaved at load) instead of a separate pass;
        # SLIME_FUSED_ROPE=0/1 overrides the default
        if fused_rope is None:
            fused_rope = os.environ.get("SLIME_FUSED_ROPE", FUSED_ROPE_DEFAULT) not in ("0", "")
        self.fused_rope = bool(fused_rope)
        # RMSNorm folding (weights.py / csrc/gemm.h): gamma into the qkv / gate-up weight columns, 1/rms in their epilogues
        if norm_fold is None:
            norm_fold = os.environ.get("SLIME_NORM_FOLD", NORM_FOLD_DEFAULT) not in ("0", "")
        self.norm_fold = bool(norm_fold) and cfg.hidden_size % 64 == 0
        d.flags = ((L.SLIME_FLAG_LEFT_PAD if cfg.tokenizer_padding_side == "left" else 0)
                   | (L.SLIME_FLAG_USE_GLOBAL_ONLY if cfg.use_global_only else 0)
                   | (L.SLIME_FLAG_USE_LOCAL_ONLY if cfg.use_local_only else 0)
                   | (L.SLIME_FLAG_ROPE_INTERLEAVED if self.fused_rope else 0)
                   | (L.SLIME_FLAG_NORM_FOLDED if self.norm_fold else 0)
                   | (L.SLIME_FLAG_ROUTER_QFORMER if cfg.mm_resampler_type == "qformer" else 0))
        self._desc = d
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self.lib.slime_ctx_create(C.byref(self._ctx), self.device.index or 0, C.byref(d)), "ctx_create")

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.slime_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], groups=ALL_GROUPS) -> None:
        """Accepts the reference model's state-dict keys."""
        self.load_weights(lambda name: state_dict[name], groups)

    def load_weights(self, get: Callable[[str], torch.Tensor], groups=ALL_GROUPS) -> None:
        """groups: which weight groups to register ("vit", "rs_local", "rs_global", "proj", "llm"); a stage
        whose group is absent fails loudly (used by the stand-alone module shims of slime_b200/model)."""
        with self._lock, torch.cuda.device(self.device):
            self._register(pack_weights(self.cfg, get, self.device, groups, self.dtype, rope_interleaved=self.fused_rope,
                                        norm_folded=self.norm_fold))

    def _register(self, weights: Dict[str, torch.Tensor]) -> None:
        """Hand the packed tensors to the library (borrowed pointers: self.weights keeps them alive) and let it
        derive the input-independent Resampler tensors."""
        with self._lock, torch.cuda.device(self.device):
            self.weights = weights
            for name, t in self.weights.items():
                self._check(self.lib.slime_ctx_set_weight(self._ctx, name.encode(), L.ptr(t), t.shape[0], t.shape[1]),
                            f"set_weight({name})")
            ws = self._workspace(self.lib.slime_finalize_workspace_bytes(self._ctx))
            self._check(self.lib.slime_ctx_finalize_weights(self._ctx, L.ptr(ws), ws.numel(), L.stream_ptr()), "finalize")
            torch.cuda.current_stream().synchronize()

    def clone(self, **cfg_overrides) -> "SlimeEngine":
        """A second context over the SAME packed weights (no copy) with some configuration attributes changed -
        top-p, merge type, padding side, ... (everything that lives in slime_model_desc rather than in the weights)."""
        other = SlimeEngine(self.cfg.replace(**cfg_overrides), self.device, max_pos=self._desc.max_pos, dtype=self.dtype,
                            fused_rope=self.fused_rope, norm_fold=self.norm_fold)
        other._register(self.weights)
        return other

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=self.device)
        return self._ws

    def _bf16(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ stages
    @_locked
    def vision_tower(self, pixels: torch.Tensor) -> torch.Tensor:
        """CLIPVisionTower.forward (reference multimodal_encoder/clip_encoder.py:46-58): [N,3,S,S] -> [N,576,D]."""
        px = pixels.to(device=self.device, dtype=self.dtype).contiguous()
        n = px.shape[0]
        feats = self._bf16(n, self.cfg.vit_patches, self.cfg.vit_hidden)
        if n == 0:
            return feats
        ws = self._workspace(self.lib.slime_vision_tower_workspace_bytes(self._ctx, n))
        self._check(self.lib.slime_vision_tower_fwd(self._ctx, L.ptr(px), n, L.ptr(feats), L.ptr(ws), ws.numel(),
                                                L.stream_ptr()), "vision_tower_fwd")
        return feats

    @_locked
    def vision_tower_split(self, pixels: torch.Tensor, crops_per_image: int):
        """The tower over B images of `crops_per_image` crops each ([B * C, 3, S, S], crop 0 of an image = its global view)
        with the output grouped for the adapter: returns (global [B,576,D], local [B * (C - 1), 576, D]) as two views of
        one buffer - the per-sample slices of reference llava_arch.py:212-225 without a gather."""
        px = pixels.to(device=self.device, dtype=self.dtype).contiguous()
        n, C = px.shape[0], int(crops_per_image)
        assert C >= 1 and n % C == 0, (n, C)
        B = n // C
        feats = self._bf16(n, self.cfg.vit_patches, self.cfg.vit_hidden)
        if n == 0:
            return feats, feats
        ws = self._workspace(self.lib.slime_vision_tower_workspace_bytes(self._ctx, n))
        self._check(self.lib.slime_vision_tower_fwd_split(self._ctx, L.ptr(px), B, C, L.ptr(feats), L.ptr(ws), ws.numel(),
                                                          L.stream_ptr()), "vision_tower_fwd_split")
        return feats[:B], feats[B:]

    @_locked
    def resampler(self, which: int, x: torch.Tensor) -> torch.Tensor:
        """Resampler.forward (reference multimodal_resampler/sampler.py:140-170); which 0 = local 144-query
        compression (sampler.post_qformer), 1 = the projector's 576-query resampler.  [n,576,D] -> [n,nq,D]."""
        x = x.to(device=self.device, dtype=self.dtype).contiguous()
        n = x.shape[0]
        nq = self.cfg.mm_resampler_dim if which == 0 else 576
        out = self._bf16(n, nq, self.cfg.vit_hidden)
        if n == 0:
            return out
        ws = self._workspace(self.lib.slime_resampler_workspace_bytes(self._ctx, which, n))
        self._check(self.lib.slime_resampler_fwd(self._ctx, which, L.ptr(x), n, L.ptr(out), L.ptr(ws), ws.numel(),
                                             L.stream_ptr()), "resampler_fwd")
        return out

    @_locked
    def projector(self, x: torch.Tensor, row_map: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None
                  ) -> torch.Tensor:
        """GatedBlock.projection (reference multimodal_projector/builder.py:53-57,180-181): [rows,D] -> [rows,H];
        row_map scatters output rows (the spatial merge of llava_arch.py:240-244 folded into the store)."""
        x2 = x.to(device=self.device, dtype=self.dtype).reshape(-1, self.cfg.vit_hidden).contiguous()
        rows = x2.shape[0]
        if out is None:
            out = self._bf16(rows, self.cfg.hidden_size)
        if rows == 0:
            return out
        ws = self._workspace(self.lib.slime_projector_workspace_bytes(self._ctx, rows))
        self._check(self.lib.slime_projector_fwd(self._ctx, L.ptr(x2), rows, L.ptr(row_map), L.ptr(out), L.ptr(ws),
                                             ws.numel(), L.stream_ptr()), "projector_fwd")
        return out

    @_locked
    def gated_projector(self, x: torch.Tensor) -> torch.Tensor:
        """GatedBlock.forward on global crops (reference multimodal_projector/builder.py:179-209): [n,576,D] -> [n,576,H]."""
        x = x.to(device=self.device, dtype=self.dtype).reshape(-1, 576, self.cfg.vit_hidden).contiguous()
        n = x.shape[0]
        out = self._bf16(n, 576, self.cfg.hidden_size)
        if n == 0:
            return out
        ws = self._workspace(self.lib.slime_gated_projector_workspace_bytes(self._ctx, n))
        self._check(self.lib.slime_gated_projector_fwd(self._ctx, L.ptr(x), n, L.ptr(out), L.ptr(ws), ws.numel(),
                                                   L.stream_ptr()), "gated_projector_fwd")
        return out

    @_locked
    def router(self, local: torch.Tensor, ids: torch.Tensor, mask: Optional[torch.Tensor],
               n_valid: Optional[torch.Tensor] = None, want_probs: bool = False):
        """TextGuidedSampler.forward + cosine selector + get_pure_text_embedding (reference
        multimodal_resampler/builder.py:189-201,248-281; llava_arch.py:162-210).
        local [B, n_per, H] -> (sel_idx [B,n_per] int32, sel_count [B] int32, probs or None)."""
        local = local.to(device=self.device, dtype=self.dtype).contiguous()
        B, n_per = local.shape[0], local.shape[1]
        T = ids.shape[1]
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        m8 = None if mask is None else (mask != 0).to(device=self.device, dtype=torch.uint8).contiguous()
        sel_idx = torch.zeros(B, max(n_per, 1), dtype=torch.int32, device=self.device)
        sel_count = torch.zeros(B, dtype=torch.int32, device=self.device)
        probs = torch.zeros(B, max(n_per, 1), dtype=torch.float32, device=self.device) if want_probs else None
        ws = self._workspace(self.lib.slime_router_workspace_bytes(self._ctx, B, n_per, T))
        self._check(self.lib.slime_router_fwd(self._ctx, L.ptr(local), n_per, L.ptr(n_valid), L.ptr(ids),
                                          L.ptr(m8), B, T, L.ptr(probs), L.ptr(sel_idx), L.ptr(sel_count), L.ptr(ws),
                                          ws.numel(), L.stream_ptr()), "router_fwd")
        return sel_idx, sel_count, probs

    @_locked
    def router_embeds(self, local: torch.Tensor, text_embeds: torch.Tensor, mask: Optional[torch.Tensor],
                      want_probs: bool = False):
        """The reference's module-level signature TextGuidedSampler.forward(local_f, text_embedding, attn_mask)
        (multimodal_resampler/builder.py:248): the prompt arrives as embeddings [B,T,H], not ids."""
        local = local.to(device=self.device, dtype=self.dtype).contiguous()
        text = text_embeds.to(device=self.device, dtype=self.dtype).contiguous()
        B, n_per = local.shape[0], local.shape[1]
        T = text.shape[1]
        m8 = None if mask is None else (mask != 0).to(device=self.device, dtype=torch.uint8).contiguous()
        sel_idx = torch.zeros(B, max(n_per, 1), dtype=torch.int32, device=self.device)
        sel_count = torch.zeros(B, dtype=torch.int32, device=self.device)
        probs = torch.zeros(B, max(n_per, 1), dtype=torch.float32, device=self.device) if want_probs else None
        ws = self._workspace(self.lib.slime_router_workspace_bytes(self._ctx, B, n_per, T))
        self._check(self.lib.slime_router_fwd_embeds(self._ctx, L.ptr(local), n_per, None, L.ptr(text), L.ptr(m8), B, T,
                                                 L.ptr(probs), L.ptr(sel_idx), L.ptr(sel_count), L.ptr(ws), ws.numel(),
                                                 L.stream_ptr()), "router_fwd_embeds")
        return sel_idx, sel_count, probs

    @_locked
    def router_select(self, probs: torch.Tensor, n_valid: Optional[torch.Tensor] = None):
        """The top-p selection rule alone, from given probabilities (bit-exact index parity)."""
        probs = probs.to(device=self.device, dtype=torch.float32).contiguous()
        B, n_per = probs.shape
        sel_idx = torch.zeros(B, n_per, dtype=torch.int32, device=self.device)
        sel_count = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.slime_router_select(self._ctx, L.ptr(probs), B, n_per, L.ptr(n_valid), L.ptr(sel_idx),
                                             L.ptr(sel_count), L.stream_ptr()), "router_select")
        return sel_idx, sel_count

    @_locked
    def splice(self, ids, mask, glob, local, sel_idx, sel_count, n_global: int, has_sep: bool,
               labels: Optional[torch.Tensor] = None, padded: bool = False, bound_rows: Optional[int] = None):
        """prepare_inputs_labels_for_multimodal's splice (reference llava_arch.py:249-255,361-459).
        Returns packed embeds [total,H], pos_ids [total] int32, cu_seqlens [B+1] int32 (device), lengths (host)
        and, when `padded`, the reference's padded (inputs_embeds, attention_mask, position_ids, labels)."""
        B, T = ids.shape
        H = self.cfg.hidden_size
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        m8 = None if mask is None else (mask != 0).to(device=self.device, dtype=torch.uint8).contiguous()
        plan = torch.empty(int(self.lib.slime_splice_plan_ints(B, T)), dtype=torch.int32, device=self.device)
        if bound_rows is not None:
            # no host round trip (slime_splice_plan_async): the lengths stay on the device, every buffer is sized by the
            # caller's upper bound and the rows past the real total are zero
            assert not padded, "the padded views need the lengths on the host"
            self._check(self.lib.slime_splice_plan_async(self._ctx, L.ptr(ids), L.ptr(m8), B, T, n_global, int(has_sep),
                                                     L.ptr(sel_count), L.ptr(plan), L.stream_ptr()), "splice_plan_async")
            lengths, total = None, int(bound_rows)
        else:
            host_cu = (C.c_int32 * (B + 1))()
            self._check(self.lib.slime_splice_plan(self._ctx, L.ptr(ids), L.ptr(m8), B, T, n_global, int(has_sep),
                                               L.ptr(sel_count), L.ptr(plan), host_cu, L.stream_ptr()), "splice_plan")
            cu_host = list(host_cu)
            lengths = [cu_host[i + 1] - cu_host[i] for i in range(B)]
            total = cu_host[B]
        embeds = self._bf16(total, H)
        pos_ids = torch.empty(total, dtype=torch.int32, device=self.device)
        g_rows = glob.shape[1] if glob is not None and glob.dim() == 3 else 0
        l_rows = local.shape[1] if local is not None and local.dim() == 3 else 0
        sel_stride = sel_idx.shape[1] if sel_idx is not None else 0
        self._check(self.lib.slime_splice_gather(self._ctx, L.ptr(ids), B, T, L.ptr(plan), L.ptr(glob), n_global, g_rows,
                                             L.ptr(local), l_rows, L.ptr(sel_idx), sel_stride, int(has_sep),
                                             L.ptr(embeds), L.ptr(pos_ids), total, L.stream_ptr()), "splice_gather")
        cu_dev = plan[B * T + B * 8: B * T + B * 8 + B + 1]
        out = dict(embeds=embeds, pos_ids=pos_ids, cu_seqlens=cu_dev, lengths=lengths, plan=plan)
        if padded:
            lmax = max(lengths) if lengths else 0
            pe = torch.empty(B, lmax, H, dtype=self.dtype, device=self.device)
            pm = torch.empty(B, lmax, dtype=torch.uint8, device=self.device)
            pp = torch.empty(B, lmax, dtype=torch.int64, device=self.device)
            pl = torch.empty(B, lmax, dtype=torch.int64, device=self.device)
            lab = None if labels is None else labels.to(device=self.device, dtype=torch.int64).contiguous()
            self._check(self.lib.slime_splice_pad(self._ctx, L.ptr(plan), L.ptr(embeds), L.ptr(lab), B, T, lmax, L.ptr(pe),
                                              L.ptr(pm), L.ptr(pp), L.ptr(pl), L.stream_ptr()), "splice_pad")
            out.update(inputs_embeds=pe, attention_mask=pm.bool(), position_ids=pp, labels=pl)
        return out

    @_locked
    def decoder_prefill(self, embeds: torch.Tensor, cu_seqlens: torch.Tensor, pos_ids: torch.Tensor,
                        lengths: Optional[Sequence[int]], want_last: bool = True, want_all: bool = False,
                        want_hidden: bool = False, batch: Optional[int] = None, max_len: Optional[int] = None):
        """LlamaForCausalLM.forward(inputs_embeds=...) (HF llama/modeling_llama.py:355-507) on packed rows.
        `lengths` (host) gives the batch size and the longest sequence; a no-host-sync caller passes `batch` and an
        upper bound `max_len` instead (embeds may then hold zero rows past cu_seqlens[-1])."""
        total = embeds.shape[0]
        B = len(lengths) if lengths is not None else int(batch)
        if max_len is None:
            max_len = max(lengths) if lengths else 0
        V = self.cfg.vocab_size
        last = torch.empty(B, V, dtype=torch.float32, device=self.device) if want_last else None
        allv = self._bf16(total, V) if want_all else None
        hid = self._bf16(total, self.cfg.hidden_size) if want_hidden else None
        ws = self._workspace(self.lib.slime_decoder_workspace_bytes(self._ctx, total, B))
        self._check(self.lib.slime_decoder_prefill_fwd(self._ctx, L.ptr(embeds), L.ptr(cu_seqlens), L.ptr(pos_ids), B, total,
                                                   int(max_len), L.ptr(last), L.ptr(allv), L.ptr(hid),
                                                   L.ptr(ws), ws.numel(), L.stream_ptr()), "decoder_prefill_fwd")
        return last, allv, hid

    # ------------------------------------------------------------------ spatial merge addressing
    def merge_row_map(self, n_local: Sequence[int], grids: Optional[Sequence[Tuple[int, int]]], n_per: int
                      ) -> torch.Tensor:
        """dst row (inside the [B, n_per, H] local buffer) of every compressed local token, in crop-major
        source order.  'spatial': raster order over the whole image (reference llava_arch.py:240-244);
        'flat': crop after crop (llava_arch.py:233-234)."""
        g = self.cfg.resampler_grid
        spatial = self.cfg.mm_patch_merge_type == "spatial"
        key = (tuple(n_local), tuple(grids) if (grids is not None and spatial) else None, n_per, spatial)
        hit = self._row_maps.get(key)
        if hit is not None:
            return hit
        rows: List[torch.Tensor] = []
        q = g * g
        for b, nl in enumerate(n_local):
            if nl == 0:
                continue
            if spatial:
                w, h = grids[b]
                if w * h != nl:
                    raise ValueError(f"sample {b}: grid {w}x{h} does not match its {nl} local crops")
                c = torch.arange(nl).view(nl, 1, 1)
                gy = torch.arange(g).view(1, g, 1)
                gx = torch.arange(g).view(1, 1, g)
                dst = (((c // w) * g + gy) * w + (c % w)) * g + gx
                rows.append(dst.reshape(-1) + b * n_per)
            else:
                rows.append(torch.arange(nl * q) + b * n_per)
        out = (torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)).to(dtype=torch.int32, device=self.device)
        if len(self._row_maps) > 256:
            self._row_maps.clear()
        self._row_maps[key] = out
        return out

    # ------------------------------------------------------------------ KV cache + decode step
    def new_kv_cache(self, batch: int, cache_len: int) -> torch.Tensor:
        """An (unattached) KV cache tensor [layers, 2, batch, cache_len, kv_heads*head_dim]."""
        cfg = self.cfg
        return torch.zeros(cfg.num_hidden_layers, 2, batch, cache_len, cfg.num_key_value_heads * cfg.head_dim,
                           dtype=self.dtype, device=self.device)

    @_locked
    def use_kv_cache(self, cache: torch.Tensor) -> torch.Tensor:
        """Attach an existing cache tensor (e.g. the one a `past_key_values` object carries between forward calls)."""
        assert cache.dim() == 5 and cache.is_contiguous() and cache.dtype == self.dtype and cache.device == self.device
        self._check(self.lib.slime_decoder_set_kv_cache(self._ctx, L.ptr(cache), cache.shape[2], cache.shape[3]),
                    "set_kv_cache")
        self._kv_cache = cache
        return cache

    def attach_kv_cache(self, batch: int, cache_len: int) -> torch.Tensor:
        """Allocate and attach a KV cache [layers, 2, batch, cache_len, kv_heads*head_dim]; while attached, every
        decoder_prefill stores K (post-RoPE) / V of its sequences into it (sequence b -> cache slot b)."""
        return self.use_kv_cache(self.new_kv_cache(batch, cache_len))

    @_locked
    def detach_kv_cache(self) -> None:
        self._check(self.lib.slime_decoder_set_kv_cache(self._ctx, None, 0, 0), "set_kv_cache")
        self._kv_cache = None

    @_locked
    def decode_step(self, x: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """One decode step: x [B, H] embeddings of the tokens to append, lens [B] int32 (device) tokens already cached
        -> logits [B, V] fp32 (HF generation loop step after the prefill; reference llava_llama.py:139)."""
        x = x.to(device=self.device, dtype=self.dtype).contiguous()
        B = x.shape[0]
        logits = torch.empty(B, self.cfg.vocab_size, dtype=torch.float32, device=self.device)
        ws = self._workspace(self.lib.slime_decoder_decode_workspace_bytes(self._ctx, B))
        self._check(self.lib.slime_decoder_decode_fwd(self._ctx, L.ptr(x), L.ptr(lens), B, L.ptr(logits), L.ptr(ws),
                                                  ws.numel(), L.stream_ptr()), "decode_fwd")
        return logits

    @torch.no_grad()
    def generate(self, pixels, input_ids, attention_mask=None, grids=None, image_sizes=None, max_new_tokens: int = 20,
                 eos_token_ids=(), sample_fn=None) -> torch.Tensor:
        """Prefill (with KV cache) + greedy / sampled decode.  sample_fn(logits [B,V]) -> next ids [B]; default
        argmax.  Returns the generated ids [B, <= max_new_tokens] (like the reference with inputs_embeds)."""
        if grids is None and image_sizes is not None and self.cfg.mm_patch_merge_type == "spatial":
            grids = [get_anyres_image_grid_shape(sz, None, self.cfg.vit_image) for sz in image_sizes]
        s = self.prefill_splice(pixels, input_ids, attention_mask, grids=grids, padded=False)["splice"]
        return self.generate_packed(s["embeds"], s["cu_seqlens"], s["pos_ids"], s["lengths"], max_new_tokens,
                                    eos_token_ids, sample_fn)

    @torch.no_grad()
    def generate_packed(self, rows, cu_seqlens, pos_ids, lengths, max_new_tokens: int = 20, eos_token_ids=(),
                        sample_fn=None, on_step=None) -> torch.Tensor:
        """Decoder prefill of packed embedding rows with the KV cache attached, then decode steps.
        on_step(next_ids [B], generated_so_far list of [B]) is called after every sampled token (streamers, stopping
        criteria of the HF generate() API); a true return value ends the generation."""
        B = len(lengths)
        cache_len = min(self._desc.max_pos, max(lengths) + max_new_tokens + 1)
        self.attach_kv_cache(B, cache_len)
        try:
            logits, _, _ = self.decoder_prefill(rows, cu_seqlens, pos_ids, lengths, want_last=True)
            lens = torch.tensor(list(lengths), dtype=torch.int32, device=self.device)
            table = self.weights["llm.embed"]
            out, done = [], torch.zeros(B, dtype=torch.bool, device=self.device)
            eos = torch.tensor(list(eos_token_ids), dtype=torch.long, device=self.device)
            for step in range(max_new_tokens):
                nxt = logits.argmax(-1) if sample_fn is None else sample_fn(logits)
                out.append(nxt)
                if on_step is not None and on_step(nxt, out):
                    break
                if eos.numel():
                    done |= torch.isin(nxt, eos)
                    if bool(done.all()):
                        break
                if step + 1 == max_new_tokens or int(lens.max()) + 1 >= cache_len:
                    break
                logits = self.decode_step(table[nxt], lens)
                lens = lens + 1
            return torch.stack(out, 1)
        finally:
            self.detach_kv_cache()

    # ------------------------------------------------------------------ module-API helpers (slime_b200/model)
    @torch.no_grad()
    def prefill_splice(self, pixels, input_ids, attention_mask=None, grids=None, labels=None, padded=True,
                       forced_selection=None) -> dict:
        """encode_images + splice without the decoder: what prepare_inputs_labels_for_multimodal returns
        (reference llava_arch.py:274-459).  Result: {'splice': {...padded + packed tensors...}, 'sel_idx', ...}."""
        res = self.prefill(pixels, input_ids, attention_mask, grids=grids, labels=labels,
                           forced_selection=forced_selection, want_last=False, keep_stages=True, run_decoder=False)
        return dict(splice=res.stages["splice"], sel_idx=res.sel_idx, sel_count=res.sel_count, lengths=res.lengths,
                    glob=res.stages.get("glob"), local=res.stages.get("local_m"))

    @torch.no_grad()
    def prefill_features(self, pixels, input_ids, attention_mask=None, grids=None) -> List[torch.Tensor]:
        """Per-sample image feature blocks [1, 576 + 1 + K_b, H] exactly as the sampler branch of the reference's
        encode_images returns them (llava_arch.py:249-255): global tokens, separator, kept local tokens."""
        cfg = self.cfg
        res = self.prefill(pixels, input_ids, attention_mask, grids=grids, want_last=False, keep_stages=True,
                           run_decoder=False, run_splice=False)
        st = res.stages
        counts = res.sel_count.cpu().tolist() if res.sel_count is not None else None
        sep = self.weights["llm.embed"][cfg.seperator][None]
        out = []
        B = input_ids.shape[0]
        for b in range(B):
            parts = []
            if not cfg.use_local_only:
                parts.append(st["glob"][b])
                if not cfg.use_global_only:
                    parts.append(sep)
            if not cfg.use_global_only:
                parts.append(st["local_m"][b][res.sel_idx[b, :counts[b]].long()])
            out.append(torch.cat(parts, 0).unsqueeze(0))
        return out

    # ------------------------------------------------------------------ whole path
    @torch.no_grad()
    def prefill(self, pixels, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                image_sizes: Optional[Sequence[Tuple[int, int]]] = None, grids: Optional[Sequence[Tuple[int, int]]] = None,
                labels: Optional[torch.Tensor] = None, forced_selection: Optional[Sequence[torch.Tensor]] = None,
                want_last: bool = True, want_all_logits: bool = False, want_probs: bool = False,
                keep_stages: bool = False, run_decoder: bool = True, run_splice: bool = True,
                sync_free: bool = False) -> PrefillResult:
        """LlavaLlamaForCausalLM.forward(images=...) (reference llava_llama.py:57-104 -> llava_arch.py:274-459).

        pixels: [B, n, 3, S, S] tensor or list of per-sample [n_b, 3, S, S] tensors (crop 0 = global view).
        forced_selection: optional per-sample index tensors that replace the router's choice (teacher forcing
        for logits parity).
        sync_free: no host synchronisation anywhere (the reference syncs per sample, llava_arch.py:170,378; the default
        path here syncs once for the spliced lengths): buffers are sized by the upper bound "every local token kept",
        the result's `lengths` stay on the device (PrefillResult.resolve_lengths()).  This is what GraphedPrefill
        captures into a CUDA graph."""
        cfg = self.cfg
        with self._lock, torch.cuda.device(self.device):
            if isinstance(pixels, (list, tuple)):
                per = [p.unsqueeze(0) if p.dim() == 3 else p for p in pixels]
                counts = [p.shape[0] for p in per]
                px = torch.cat(per, 0)
            else:
                counts = [pixels.shape[1]] * pixels.shape[0]
                px = pixels.reshape(-1, *pixels.shape[2:])
            B = len(counts)
            assert input_ids.shape[0] == B, "one image (stack of crops) per sample"
            px = px.to(device=self.device, dtype=self.dtype, non_blocking=True)
            ids = input_ids.to(device=self.device, dtype=torch.int64, non_blocking=True)
            mask = None if attention_mask is None else attention_mask.to(device=self.device, non_blocking=True)
            stages = {} if keep_stages else None

            starts = [0]
            for c in counts:
                starts.append(starts[-1] + c)
            uniform = len(set(counts)) == 1
            n_local = [c - 1 for c in counts]
            if uniform and not keep_stages:
                # every image has the same number of crops: the tower writes the global crops first and the local crops
                # behind them, so both adapter inputs are contiguous views
                xg, xl = self.vision_tower_split(px, counts[0])
                feats = None
            elif uniform:
                feats = self.vision_tower(px)                               # [Nc, 576, D]
                fv = feats.view(B, counts[0], cfg.vit_patches, cfg.vit_hidden)
                xg = fv[:, 0]
                xl = fv[:, 1:].reshape(-1, cfg.vit_patches, cfg.vit_hidden)
            else:
                feats = self.vision_tower(px)
                gi = torch.tensor(starts[:-1], device=self.device)
                li = torch.tensor([i for b in range(B) for i in range(starts[b] + 1, starts[b + 1])], device=self.device,
                                  dtype=torch.long)
                xg, xl = feats.index_select(0, gi), feats.index_select(0, li)

            glob = local = sel_idx = sel_count = probs = None
            n_global, has_sep = 0, False
            if not cfg.use_local_only:
                glob = self.gated_projector(xg)                              # [B, 576, H]
                n_global = 576
                has_sep = not cfg.use_global_only
            q = cfg.mm_resampler_dim
            n_per = max(n_local) * q if n_local else 0
            if not cfg.use_global_only:
                local = torch.zeros(B, max(n_per, 1), cfg.hidden_size, dtype=self.dtype, device=self.device) \
                    if not uniform else self._bf16(B, max(n_per, 1), cfg.hidden_size)
                if n_per > 0:
                    if grids is None and cfg.mm_patch_merge_type == "spatial":
                        if image_sizes is None:
                            raise ValueError("spatial merge needs image_sizes or grids")
                        grids = [get_anyres_image_grid_shape(s, None, cfg.vit_image) for s in image_sizes]
                    lc = self.resampler(0, xl)                               # [Nl, 144, D]
                    rmap = self.merge_row_map(n_local, grids, n_per)
                    self.projector(lc, row_map=rmap, out=local.view(-1, cfg.hidden_size))
                    n_valid = None if uniform else torch.tensor([n * q for n in n_local], dtype=torch.int32,
                                                                device=self.device)
                    sel_idx, sel_count, probs = self.router(local, ids, mask, n_valid, want_probs)
                    if forced_selection is not None:
                        sel_idx = torch.zeros_like(sel_idx)
                        for b, s in enumerate(forced_selection):
                            sel_idx[b, : s.numel()] = s.to(device=self.device, dtype=torch.int32)
                        sel_count = torch.tensor([s.numel() for s in forced_selection], dtype=torch.int32,
                                                 device=self.device)
                    if stages is not None:
                        stages.update(local_c=lc)
                else:
                    sel_idx = torch.zeros(B, 1, dtype=torch.int32, device=self.device)
                    sel_count = torch.zeros(B, dtype=torch.int32, device=self.device)
            if not run_splice:
                if stages is not None:
                    stages.update(vit=feats, glob=glob, local_m=local)
                return PrefillResult(logits_last=None, logits_all=None, cu_seqlens=None, lengths=[], sel_idx=sel_idx,
                                     sel_count=sel_count, probs=probs, embeds=None, stages=stages)
            bound_rows = bound_len = None
            if sync_free:
                assert not keep_stages and labels is None, "sync_free: no padded views / labels (they need host lengths)"
                bound_len = ids.shape[1] + n_global + int(has_sep) + (n_per if not cfg.use_global_only else 0)
                if cfg.tokenizer_model_max_length:
                    bound_len = min(bound_len, int(cfg.tokenizer_model_max_length))
                bound_rows = B * bound_len
            sp = self.splice(ids, mask, glob, local, sel_idx, sel_count, n_global, has_sep, labels=labels,
                             padded=keep_stages, bound_rows=bound_rows)
            last = allv = None
            if run_decoder:
                last, allv, _ = self.decoder_prefill(sp["embeds"], sp["cu_seqlens"], sp["pos_ids"], sp["lengths"],
                                                     want_last=want_last, want_all=want_all_logits, batch=B,
                                                     max_len=bound_len)
            if stages is not None:
                stages.update(vit=feats, glob=glob, local_m=local, splice=sp)
            return PrefillResult(logits_last=last, logits_all=allv, cu_seqlens=sp["cu_seqlens"], lengths=sp["lengths"],
                                 sel_idx=sel_idx, sel_count=sel_count, probs=probs,
                                 embeds=sp["embeds"] if keep_stages else None, stages=stages)


class GraphedPrefill:
    """One prefill for FIXED shapes (batch, crops per image, prompt length) captured in a CUDA graph: a step is three small
    copies into static input buffers + one cudaGraphLaunch instead of ~400 kernel launches and a host synchronisation.
    This is the batch-1 latency path (the reference's eval loop is batch 1, llava/eval/model_vqa_loader.py:103-119):
    at that size the kernels are 10-100 us long and launch gaps / the mid-pipeline sync are a visible share of the step.

        g = GraphedPrefill(engine, batch=1, n_crops=5, prompt_len=256, grids=[(2, 2)])
        res = g(pixels, input_ids, attention_mask)       # res.logits_last [B, V] fp32 (static buffer, overwritten by
                                                         # the next call), res.cu_seqlens on the device
    Built on SlimeEngine.prefill(sync_free=True): buffers are sized by the upper bound "every local token kept", the
    real lengths stay on the device, rows past them are zero and belong to no sequence."""

    def __init__(self, engine: SlimeEngine, batch: int, n_crops: int, prompt_len: int,
                 grids: Optional[Sequence[Tuple[int, int]]] = None, warmup: int = 2):
        cfg = engine.cfg
        self.engine = engine
        dev = engine.device
        S = cfg.vit_image
        self.pixels = torch.zeros(batch, n_crops, 3, S, S, dtype=engine.dtype, device=dev)
        self.ids = torch.zeros(batch, prompt_len, dtype=torch.int64, device=dev)
        self.ids[:, min(1, prompt_len - 1)] = IMAGE_TOKEN_INDEX
        self.mask = torch.ones(batch, prompt_len, dtype=torch.int64, device=dev)
        self.grids = list(grids) if grids is not None else None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up on a side stream: sizes the workspace, fills the tensor-map / row-map
            for _ in range(max(1, warmup)):  # caches and sets the kernels' attributes before the capture
                engine.prefill(self.pixels, self.ids, self.mask, grids=self.grids, sync_free=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = engine.prefill(self.pixels, self.ids, self.mask, grids=self.grids, sync_free=True)

    @torch.no_grad()
    def __call__(self, pixels: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None
                 ) -> PrefillResult:
        with self.engine._lock:
            self.pixels.copy_(pixels.reshape(self.pixels.shape), non_blocking=True)
            self.ids.copy_(input_ids, non_blocking=True)
            if attention_mask is not None:
                self.mask.copy_(attention_mask, non_blocking=True)
            else:
                self.mask.fill_(1)
            self.graph.replay()
            r = self.result
            return PrefillResult(logits_last=r.logits_last, logits_all=None, cu_seqlens=r.cu_seqlens, lengths=None,
                                 sel_idx=r.sel_idx, sel_count=r.sel_count, probs=None, embeds=None)
