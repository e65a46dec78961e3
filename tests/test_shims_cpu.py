"""CPU checks of the drop-in module API (slime_b200/model): same class names / attributes / state-dict keys as
the reference's llava/model package, and no silent CPU fallback."""
import json
import os
import tempfile

import pytest
import torch

from slime_b200.config import SlimeConfig, preset
from slime_b200.synth import synth_state_dict


def make_model(pname="tiny", **over):
    from slime_b200.model import LlavaConfig, LlavaLlamaForCausalLM

    cfg = preset(pname, **over)
    tmp = tempfile.mkdtemp(prefix="slime_clip_cfg_")
    with open(os.path.join(tmp, "config.json"), "w") as f:
        json.dump(dict(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                       num_attention_heads=cfg.vit_heads, image_size=cfg.vit_image, patch_size=cfg.vit_patch,
                       layer_norm_eps=cfg.vit_ln_eps), f)
    hf = LlavaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, vocab_size=cfg.vocab_size,
                     rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                     max_position_embeddings=cfg.max_position_embeddings, pad_token_id=cfg.pad_token_id,
                     mm_vision_tower=tmp, mm_vision_select_layer=cfg.mm_vision_select_layer,
                     mm_vision_select_feature="patch", mm_projector_type="gated", mm_hidden_size=cfg.vit_hidden,
                     mm_resampler_type=cfg.mm_resampler_type, mm_resampler_dim=cfg.mm_resampler_dim,
                     mm_resampler_topp=cfg.mm_resampler_topp, mm_resampler_temp=cfg.mm_resampler_temp,
                     mm_learnable_gated=-1, mm_patch_merge_type=cfg.mm_patch_merge_type, image_aspect_ratio="anyres",
                     seperator=cfg.seperator, tokenizer_padding_side=cfg.tokenizer_padding_side,
                     tokenizer_model_max_length=cfg.tokenizer_model_max_length,
                     image_grid_pinpoints=[[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]])
    model = LlavaLlamaForCausalLM(hf)
    model.get_vision_tower().load_model()
    return cfg, model


def test_state_dict_keys_match_reference_exactly():
    """synth_state_dict is strict-loaded by the UNMODIFIED reference in oracle/gen_golden.py; loading the same
    dict strictly here proves the shim tree has exactly the reference's parameter names and shapes."""
    cfg, model = make_model()
    sd = synth_state_dict(cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(model.state_dict().keys()) == set(sd.keys())


def test_qformer_router_state_dict_keys_match_reference():
    """mm_resampler_type='qformer' (TextGuidedRouterAttention, reference multimodal_resampler/builder.py:94-162): the same
    synthetic dict the unmodified reference strict-loads for the golden case tiny_qformer_router_b2."""
    cfg, model = make_model(mm_resampler_type="qformer")
    sd = synth_state_dict(cfg)
    assert "model.sampler.selector.cross_attn.in_proj_weight" in sd and "model.sampler.selector.query" in sd
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(model.state_dict().keys()) == set(sd.keys())
    assert type(model.get_model().sampler.selector).__name__ == "TextGuidedRouterAttention"


def test_reference_attributes_present():
    cfg, model = make_model()
    inner = model.get_model()
    assert inner.has_sampler is True
    vt = model.get_vision_tower()
    assert vt.is_loaded and vt.hidden_size == cfg.vit_hidden and vt.num_patches == 576
    assert vt.num_patches_per_side == 24 and vt.config.image_size == 336
    assert inner.sampler.grid_size == 12 and hasattr(inner.sampler, "post_qformer")
    assert inner.mm_projector.expert_ffn[0] is inner.mm_projector.projection
    assert inner.mm_projector.expert_ffn[1] is inner.mm_projector.attn
    for name in ("forward", "generate", "prepare_inputs_for_generation", "encode_images", "get_pure_text_embedding",
                 "prepare_inputs_labels_for_multimodal", "initialize_vision_tokenizer", "get_model",
                 "get_vision_tower"):
        assert callable(getattr(model, name))
    from slime_b200.model import LlavaConfig

    assert LlavaConfig.model_type == "llava_llama"
    got = model._slime_config()
    for k in ("vit_hidden", "vit_layers", "vit_heads", "vit_mlp", "hidden_size", "num_hidden_layers",
              "num_attention_heads", "num_key_value_heads", "head_dim", "intermediate_size", "vocab_size",
              "mm_resampler_dim", "mm_resampler_topp", "seperator", "mm_patch_merge_type"):
        assert getattr(got, k) == getattr(cfg, k), k


def test_no_cpu_fallback():
    cfg, model = make_model()
    model.load_state_dict(synth_state_dict(cfg))
    ids = torch.randint(3, 100, (1, 8))
    ids[0, 2] = -200
    px = torch.randn(1, 5, 3, 336, 336)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_shims_gpu.py")
    with pytest.raises(RuntimeError, match="CUDA"):
        model(input_ids=ids, images=px, image_sizes=[(672, 672)])
    with pytest.raises(RuntimeError, match="CUDA"):
        model.get_vision_tower()(px[0])


def test_unsupported_variants_fail_loudly():
    from types import SimpleNamespace

    from slime_b200.model.multimodal_projector.builder import build_vision_projector
    from slime_b200.model.multimodal_resampler.builder import build_vision_sampler

    with pytest.raises(NotImplementedError):
        build_vision_projector(SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=1024, hidden_size=4096))
    with pytest.raises(NotImplementedError):
        build_vision_sampler(SimpleNamespace(mm_resampler_type="perceiver", mm_resampler_dim=144, mm_resampler_topp=0.9,
                                             mm_resampler_temp=1.0, mm_hidden_size=1024, hidden_size=4096))


class _StubEngine:
    """Stands in for SlimeEngine in generate(): emits a scripted token sequence through the same generate_packed
    contract (sample_fn on [B, V] logits, on_step hook, EOS handling) so the HF-API plumbing is testable without a GPU."""

    def __init__(self, vocab, hidden, script):
        self.device = torch.device("cpu")
        self.weights = {"llm.embed": torch.randn(vocab, hidden)}
        self.script = script  # list of token ids per step (same for every sequence)
        self.vocab = vocab

    def generate_packed(self, rows, cu_seqlens, pos_ids, lengths, max_new_tokens=20, eos_token_ids=(), sample_fn=None,
                        on_step=None):
        B = len(lengths)
        out, done = [], torch.zeros(B, dtype=torch.bool)
        eos = torch.tensor(list(eos_token_ids), dtype=torch.long)
        for step in range(max_new_tokens):
            logits = torch.full((B, self.vocab), -10.0)
            logits[:, self.script[min(step, len(self.script) - 1)]] = 10.0
            nxt = logits.argmax(-1) if sample_fn is None else sample_fn(logits)
            out.append(nxt)
            if on_step is not None and on_step(nxt, out):
                break
            if eos.numel():
                done |= torch.isin(nxt, eos)
                if bool(done.all()):
                    break
        return torch.stack(out, 1)


def test_generate_streamer_and_stopping_criteria_hooks(monkeypatch):
    """generate(streamer=..., stopping_criteria=[...]) as the reference's callers use it (serve/cli.py:95-105,
    serve/model_worker.py:168-189, mm_utils.py:292 KeywordsStoppingCriteria): prompt put first, one put per token,
    end(); a criterion returning True (bool or tensor) stops the loop; tokens after EOS are padded."""
    cfg, model = make_model()
    stub = _StubEngine(cfg.vocab_size, cfg.hidden_size, script=[11, 12, 13, 14, 15, 16])
    monkeypatch.setattr(type(model), "_engine", lambda self, device=None: stub)
    ids = torch.randint(3, 100, (1, 7))

    class Streamer:
        def __init__(self):
            self.puts, self.ended = [], False

        def put(self, v):
            self.puts.append(v.clone())

        def end(self):
            self.ended = True

    st = Streamer()
    out = model.generate(ids, max_new_tokens=4, streamer=st)
    assert out.tolist() == [[11, 12, 13, 14]]
    assert st.ended and len(st.puts) == 5 and st.puts[0].shape == (1, 0)
    assert [int(p[0]) for p in st.puts[1:]] == [11, 12, 13, 14]

    seen = []

    def crit_bool(output_ids, scores):
        seen.append(output_ids.clone())
        return output_ids.shape[1] >= 3

    out = model.generate(ids, max_new_tokens=6, stopping_criteria=[crit_bool])
    assert out.tolist() == [[11, 12, 13]]
    assert [s.shape[1] for s in seen] == [1, 2, 3] and seen[-1].tolist() == [[11, 12, 13]]
    out = model.generate(ids, max_new_tokens=6, stopping_criteria=[lambda o, s: torch.tensor([o[0, -1] == 12])])
    assert out.tolist() == [[11, 12]]
    # EOS: generation ends when every sequence has produced it
    out = model.generate(ids, max_new_tokens=6, eos_token_id=13)
    assert out.tolist() == [[11, 12, 13]]
    with pytest.raises(NotImplementedError):
        model.generate(ids, max_new_tokens=2, num_beams=4)
