"""Drop-in for the reference's llava/model/language_model/llava_llama.py: LlavaConfig, LlavaLlamaModel and
LlavaLlamaForCausalLM with the reference's forward / generate / prepare_inputs_for_generation signatures and
state-dict keys (model.embed_tokens, model.layers.N.*, model.norm, lm_head, model.vision_tower.*,
model.mm_projector.*, model.sampler.*).  The nn.Modules only HOLD the parameters; every FLOP of forward()
runs in libslime_b200 (vision tower, adapter, router, splice, Llama prefill on packed rows).

generate() = packed prefill with the KV cache attached + native decode steps (slime_decoder_decode_fwd); greedy or
temperature / top-p sampling (the token choice itself is torch plumbing).  forward(use_cache=True) returns a
`SlimeKVCache` as `past_key_values`; handing it back with one new token per sequence runs one native decode step (the
HF generation loop's contract, reference :57-104,146-157).  Not built: beam search (INTEGRATION.md names the callers
that pass num_beams), training (loss is provided for parity of the forward signature, computed from the returned
logits with torch).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ...config import IGNORE_INDEX, SlimeConfig
from .._runtime import EngineBinding, bind, binding_of
from ..llava_arch import LlavaMetaForCausalLM, LlavaMetaModel


class LlavaConfig:
    """Attribute bag with LlamaConfig's defaults; `model_type` as registered by the reference (:30-31)."""

    model_type = "llava_llama"
    _defaults = dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                     num_attention_heads=32, num_key_value_heads=None, rms_norm_eps=1e-5, rope_theta=10000.0,
                     max_position_embeddings=4096, pad_token_id=0, bos_token_id=1, eos_token_id=2, pretraining_tp=1,
                     tie_word_embeddings=False, torch_dtype="bfloat16")

    def __init__(self, **kwargs):
        for k, v in self._defaults.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            setattr(self, k, v)
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads

    @classmethod
    def from_pretrained(cls, path: str, **kwargs):
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        raw.update(kwargs)
        return cls(**raw)

    def to_dict(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}


@dataclass
class CausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.Tensor]] = None
    attentions: Optional[Tuple[torch.Tensor]] = None

    def __getitem__(self, i):
        return tuple(v for v in (self.loss, self.logits) if v is not None)[i]


class SlimeKVCache:
    """`past_key_values` of the drop-in model: the engine's KV cache tensor [layers, 2, B, cache_len, kv_heads*head_dim]
    (K post-RoPE) plus the number of cached tokens per sequence.  Callers treat it as opaque and hand it back to
    forward(); get_seq_length() is what HF's generation utilities ask a cache for.  Grows by re-allocation."""

    def __init__(self, engine, cache: torch.Tensor, lens: torch.Tensor):
        self.engine, self.cache, self.lens = engine, cache, lens
        self._max_len = int(lens.max()) if lens.numel() else 0

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self._max_len

    @property
    def batch(self) -> int:
        return self.cache.shape[2]

    def reserve(self, extra: int) -> None:
        """Room for `extra` more tokens per sequence (doubling re-allocation, bounded by the engine's max positions)."""
        need = self._max_len + extra
        have = self.cache.shape[3]
        if need <= have:
            return
        limit = self.engine._desc.max_pos
        if need > limit:
            raise RuntimeError(f"KV cache would need {need} positions, the engine was built for {limit}")
        new_len = min(limit, max(need, 2 * have))
        new = self.engine.new_kv_cache(self.batch, new_len)
        new[:, :, :, :have] = self.cache
        self.cache = new

    def advance(self, n: int = 1) -> None:
        self.lens = self.lens + n
        self._max_len += n


class _Norm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _Bag(nn.Module):
    pass


class _LlamaParams(nn.Module):
    """Parameter tree of HF LlamaModel (embed_tokens, layers.N.{self_attn,mlp,*layernorm}, norm)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H, I = config.hidden_size, config.intermediate_size
        hd = getattr(config, "head_dim", None) or H // config.num_attention_heads
        qd, kd = config.num_attention_heads * hd, config.num_key_value_heads * hd
        self.embed_tokens = nn.Embedding(config.vocab_size, H)
        layers = []
        for _ in range(config.num_hidden_layers):
            l = _Bag()
            sa = _Bag()
            sa.q_proj = nn.Linear(H, qd, bias=False)
            sa.k_proj = nn.Linear(H, kd, bias=False)
            sa.v_proj = nn.Linear(H, kd, bias=False)
            sa.o_proj = nn.Linear(qd, H, bias=False)
            l.self_attn = sa
            mlp = _Bag()
            mlp.gate_proj = nn.Linear(H, I, bias=False)
            mlp.up_proj = nn.Linear(H, I, bias=False)
            mlp.down_proj = nn.Linear(I, H, bias=False)
            l.mlp = mlp
            l.input_layernorm = _Norm(H)
            l.post_attention_layernorm = _Norm(H)
            layers.append(l)
        self.layers = nn.ModuleList(layers)
        self.norm = _Norm(H)


class LlavaLlamaModel(LlavaMetaModel, _LlamaParams):
    config_class = LlavaConfig

    def __init__(self, config):
        super(LlavaLlamaModel, self).__init__(config)


class LlavaLlamaForCausalLM(nn.Module, LlavaMetaForCausalLM):
    config_class = LlavaConfig

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = LlavaLlamaModel(config)
        self.pretraining_tp = getattr(config, "pretraining_tp", 1)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        # one engine for the whole tree: sub-modules called on their own (tower, projector, sampler) share it
        # ("router" is only packed when mm_resampler_type == 'qformer': slime_b200/weights.py)
        bind(self, EngineBinding(self, self._slime_config, "", ("vit", "rs_local", "rs_global", "proj", "llm", "router")))

    # ------------------------------------------------------------------ plumbing
    def get_model(self):
        return self.model

    @property
    def device(self):
        return self.lm_head.weight.device

    @property
    def dtype(self):
        return self.lm_head.weight.dtype

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    def _slime_config(self) -> SlimeConfig:
        vt = self.get_vision_tower()
        return SlimeConfig.from_hf_config(self.config, vt.config if vt is not None else None)

    def _engine(self, device=None):
        vt = self.get_vision_tower()
        if vt is None:
            raise RuntimeError("config has no mm_vision_tower: nothing to run on the SliME path")
        if not vt.is_loaded:
            vt.load_model()
            vt.to(device=self.device, dtype=self.dtype)
        b = binding_of(self)
        return b.engine(device if device is not None and device.type == "cuda" else None)

    # ------------------------------------------------------------------ forward (reference :57-104)
    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                images=None, images_mask=None, image_sizes=None, return_dict=None):
        if output_attentions or output_hidden_states:
            raise NotImplementedError("output_attentions / output_hidden_states are not produced by the fused path")
        eng = self._engine(input_ids.device if input_ids is not None else inputs_embeds.device)
        left = getattr(self.config, "tokenizer_padding_side", "right") == "left"
        if past_key_values is not None:
            # one decode step: the HF generation loop hands back the cache with ONE new token per sequence
            # (reference :57-104 passes past_key_values through; llava_arch.py:279 is the early-out for this call)
            if not isinstance(past_key_values, SlimeKVCache):
                raise TypeError("past_key_values must be the SlimeKVCache a previous forward(use_cache=True) returned")
            x = inputs_embeds if inputs_embeds is not None else eng.weights["llm.embed"][input_ids.to(eng.device)]
            if x.dim() != 3 or x.shape[1] != 1 or x.shape[0] != past_key_values.batch:
                raise NotImplementedError("with past_key_values, forward() takes exactly one new token per cached sequence "
                                          f"(got {tuple(x.shape[:2])} for a cache of {past_key_values.batch} sequences)")
            pkv = past_key_values
            pkv.reserve(1)
            eng.use_kv_cache(pkv.cache)
            try:
                logits = eng.decode_step(x[:, 0], pkv.lens)
            finally:
                eng.detach_kv_cache()
            pkv.advance(1)
            return CausalLMOutputWithPast(logits=logits[:, None, :].to(eng.dtype), past_key_values=pkv)
        if inputs_embeds is None:
            if images is None or input_ids.shape[1] == 1:
                inputs_embeds = eng.weights["llm.embed"][input_ids.to(eng.device)]
            else:
                sp = self._spliced(input_ids, attention_mask, labels, images, image_sizes, images_mask,
                                   padded=True)["splice"]
                return self._decode_packed(eng, sp["embeds"], sp["cu_seqlens"], sp["pos_ids"], sp["lengths"], left,
                                           sp["labels"] if labels is not None else None, use_cache)
        # caller-provided embeddings [B, L, H] (+ attention_mask): pack the real rows and run the decoder
        B, Lm, _ = inputs_embeds.shape
        am = torch.ones(B, Lm, dtype=torch.bool, device=inputs_embeds.device) if attention_mask is None \
            else attention_mask.bool()
        lengths = am.sum(1).tolist()
        rows = inputs_embeds[am].to(device=eng.device, dtype=eng.dtype).contiguous()
        cu = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0)), dtype=torch.int32, device=eng.device)
        pos = torch.cat([torch.arange(L) for L in lengths]).to(device=eng.device, dtype=torch.int32)
        if position_ids is not None:
            pos = position_ids.expand(B, Lm)[am].to(device=eng.device, dtype=torch.int32)
        return self._decode_packed(eng, rows, cu, pos, lengths, left, labels, use_cache)

    def _decode_packed(self, eng, rows, cu, pos, lengths, left, labels, use_cache=None):
        pkv = None
        if use_cache:
            # the prefill stores K (post-RoPE) / V of every real token into the attached cache (sequence b -> slot b)
            cache = eng.attach_kv_cache(len(lengths), min(eng._desc.max_pos, max(lengths) + 256))
            try:
                _, allv, _ = eng.decoder_prefill(rows, cu, pos, lengths, want_last=False, want_all=True)
            finally:
                eng.detach_kv_cache()
            pkv = SlimeKVCache(eng, cache, torch.tensor(list(lengths), dtype=torch.int32, device=eng.device))
        else:
            _, allv, _ = eng.decoder_prefill(rows, cu, pos, lengths, want_last=False, want_all=True)
        B, Lmax, V = len(lengths), max(lengths), self.vocab_size
        logits = torch.zeros(B, Lmax, V, dtype=allv.dtype, device=allv.device)
        off = 0
        for b, L in enumerate(lengths):
            if left:
                logits[b, Lmax - L:] = allv[off:off + L]
            else:
                logits[b, :L] = allv[off:off + L]
            off += L
        loss = None
        if labels is not None:
            shift_logits = logits[:, :-1].float().reshape(-1, V)
            shift_labels = labels[:, 1:].reshape(-1).to(logits.device)
            loss = torch.nn.functional.cross_entropy(shift_logits, shift_labels, ignore_index=IGNORE_INDEX)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=pkv)

    # ------------------------------------------------------------------ generate (reference :106-144)
    @torch.no_grad()
    def generate(self, inputs=None, images=None, image_sizes=None, **kwargs):
        position_ids = kwargs.pop("position_ids", None)
        attention_mask = kwargs.pop("attention_mask", None)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")
        max_new = int(kwargs.pop("max_new_tokens", 20))
        do_sample = bool(kwargs.pop("do_sample", False))
        temperature = float(kwargs.pop("temperature", 1.0) or 1.0)
        top_p = kwargs.pop("top_p", None)
        if int(kwargs.pop("num_beams", 1) or 1) != 1:
            raise NotImplementedError("beam search is not built: the reference's eval loops default to num_beams=1 "
                                      "(llava/eval/model_vqa_loader.py:117,142; run_llava.py:122,146) - see INTEGRATION.md")
        # HF generate() hooks the reference's callers use: serve/cli.py:95-105 and serve/model_worker.py:168-189 read the
        # output through a `streamer`, eval scripts pass `stopping_criteria=[KeywordsStoppingCriteria(...)]` (mm_utils.py:292)
        streamer = kwargs.pop("streamer", None)
        stopping = kwargs.pop("stopping_criteria", None) or []
        eos = kwargs.pop("eos_token_id", getattr(self.config, "eos_token_id", None))
        eos = set(eos) if isinstance(eos, (list, tuple)) else ({eos} if eos is not None else set())
        eng = self._engine(inputs.device)
        table = eng.weights["llm.embed"]
        if images is not None:
            sp = self._spliced(inputs, attention_mask, None, images, image_sizes, None, padded=False)["splice"]
            rows, cu_t, pos, lengths = sp["embeds"], sp["cu_seqlens"], sp["pos_ids"], list(sp["lengths"])
        else:
            am = torch.ones_like(inputs, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
            seqs = [table[inputs[b][am[b]].to(table.device)] for b in range(inputs.shape[0])]
            lengths = [s.shape[0] for s in seqs]
            rows = torch.cat(seqs).contiguous()
            cu_t = torch.tensor([0] + list(torch.tensor(lengths).cumsum(0)), dtype=torch.int32, device=eng.device)
            pos = torch.cat([torch.arange(L) for L in lengths]).to(device=eng.device, dtype=torch.int32)
        # sampling follows torch's global RNG (torch.manual_seed), like HF generate(); `seed=` pins one call
        gen = torch.Generator(device=eng.device)
        seed = kwargs.pop("seed", None)
        gen.manual_seed(int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item()))

        def sample(last):
            if not do_sample:
                return last.argmax(-1)
            probs = torch.softmax(last / max(temperature, 1e-5), dim=-1)
            if top_p is not None and top_p < 1.0:
                sp_, si = torch.sort(probs, descending=True)
                keep = (torch.cumsum(sp_, -1) - sp_) < top_p
                sp_ = sp_ * keep
                probs = torch.zeros_like(probs).scatter(1, si, sp_ / sp_.sum(-1, keepdim=True))
            return torch.multinomial(probs, 1, generator=gen).squeeze(1)

        if streamer is not None:
            # like GenerationMixin with inputs_embeds: the first put() carries the (empty) prompt ids
            streamer.put(torch.zeros(inputs.shape[0], 0, dtype=torch.long))

        def on_step(nxt, so_far):
            if streamer is not None:
                streamer.put(nxt.cpu())
            if stopping:
                ids_so_far = torch.stack(so_far, 1)  # generated tokens only, as HF passes them for inputs_embeds prompts
                for crit in stopping:
                    r = crit(ids_so_far, None)
                    if bool(r.all()) if torch.is_tensor(r) else bool(r):
                        return True
            return False

        # prefill with the KV cache attached, then native decode steps (slime_decoder_decode_fwd)
        toks = eng.generate_packed(rows, cu_t, pos, lengths, max_new, tuple(eos), sample,
                                   on_step if (streamer is not None or stopping) else None)
        if streamer is not None:
            streamer.end()
        # like HF: positions after a sequence's EOS are padded
        pad = getattr(self.config, "pad_token_id", 0) or 0
        if eos:
            e = torch.tensor(sorted(eos), device=toks.device)
            hit = torch.isin(toks, e)
            after = (hit.cumsum(1) - hit.long()) > 0
            toks = toks.masked_fill(after, pad)
        return toks.to(inputs.device)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        images = kwargs.pop("images", None)
        image_sizes = kwargs.pop("image_sizes", None)
        inputs = dict(input_ids=input_ids, past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)
        if images is not None:
            inputs["images"] = images
        if image_sizes is not None:
            inputs["image_sizes"] = image_sizes
        return inputs


def _register_with_transformers():
    """AutoConfig / AutoModelForCausalLM registration as in the reference (:159-160), when transformers is
    importable and the name is still free (the reference itself may already be registered)."""
    try:
        from transformers import AutoConfig

        AutoConfig.register("llava_llama", LlavaConfig)  # type: ignore[arg-type]
    except Exception:
        pass
