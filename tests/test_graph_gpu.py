"""No-host-sync prefill (slime_splice_plan_async; the reference synchronises per sample, llava_arch.py:170,378) and
its CUDA-graph capture: results must be BIT-IDENTICAL to the default path (same kernels, same order; the extra zero
rows past the real total belong to no sequence)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(pname, **over):
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_state_dict

    cfg = preset(pname, **over)
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(synth_state_dict(cfg))
    return cfg, eng


@pytest.mark.parametrize("pname,over", [("tiny", {}), ("small", {}), ("tiny", {"mm_patch_merge_type": "flat"}),
                                        ("tiny", {"tokenizer_model_max_length": 700})])
def test_sync_free_prefill_is_bit_identical(pname, over):
    from slime_b200.synth import synth_inputs

    cfg, eng = _engine(pname, **over)
    B, n, T = 3, 5, 24
    px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=5, ragged=True)
    grids = None if cfg.mm_patch_merge_type == "flat" else [(2, 2)] * B
    ref = eng.prefill(px, ids, mask, grids=grids)
    res = eng.prefill(px, ids, mask, grids=grids, sync_free=True)
    assert res.lengths is None                                   # nothing came back to the host
    assert torch.equal(res.cu_seqlens.cpu(), ref.cu_seqlens.cpu())
    assert res.resolve_lengths() == ref.lengths and res.total_tokens == ref.total_tokens
    assert torch.equal(res.logits_last, ref.logits_last)
    assert torch.equal(res.sel_count, ref.sel_count)
    # with a KV cache attached the zero rows past the real total must not touch any cache slot
    eng.attach_kv_cache(B, max(ref.lengths) + 4)
    try:
        eng.prefill(px, ids, mask, grids=grids)
        a = eng._kv_cache.clone()
        eng._kv_cache.zero_()
        eng.prefill(px, ids, mask, grids=grids, sync_free=True)
        assert torch.equal(eng._kv_cache, a)
    finally:
        eng.detach_kv_cache()


def test_graphed_prefill_matches_eager_and_follows_new_inputs():
    from slime_b200.engine import GraphedPrefill
    from slime_b200.synth import synth_inputs

    cfg, eng = _engine("small")
    B, n, T = 2, 5, 20
    g = GraphedPrefill(eng, B, n, T, grids=[(2, 2)] * B)
    for seed in (1, 2, 3):
        px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=4, ragged=True, seed=seed)
        ref = eng.prefill(px, ids, mask, grids=[(2, 2)] * B)
        res = g(px.cuda(), ids.cuda(), mask.cuda())
        assert torch.equal(res.logits_last, ref.logits_last), f"seed {seed}"
        assert res.resolve_lengths() == ref.lengths


def test_graphed_prefill_fullsize_batch1_latency():
    """SliME-Llama3-8B, batch 1 (the reference's eval loop, llava/eval/model_vqa_loader.py:103-119): the captured graph is
    bit-identical to the eager path; both latencies are printed."""
    from slime_b200.engine import GraphedPrefill
    from slime_b200.synth import synth_inputs
    from tests.test_fullsize_gpu import build

    cfg, eng, get, specs = build("llama3-8b")
    px, ids, mask = synth_inputs(cfg, 1, 5, 256, seed=3)
    px, ids, mask = px.cuda().to(torch.bfloat16), ids.cuda(), mask.cuda()
    ref = eng.prefill(px, ids, mask, grids=[(2, 2)])
    g = GraphedPrefill(eng, 1, 5, 256, grids=[(2, 2)])
    res = g(px, ids, mask)
    assert torch.equal(res.logits_last, ref.logits_last)
    assert res.resolve_lengths() == ref.lengths

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_eager = timed(lambda: eng.prefill(px, ids, mask, grids=[(2, 2)]).total_tokens)
    t_graph = timed(lambda: g(px, ids, mask))
    print(f"[llama3-8b batch 1, L={ref.lengths[0]}] prefill latency: eager {t_eager:.2f} ms, CUDA graph {t_graph:.2f} ms")
    assert t_graph < t_eager * 1.05


@pytest.mark.parametrize("pname,B", [("tiny", 3), ("small", 2)])
def test_prefill_with_programmatic_dependent_launch_is_bit_identical(pname, B):
    """slime_set_prefill_pdl(1): the GEMM / attention / norm kernels are launched with the programmatic attribute, so a
    kernel's prologue overlaps its predecessor's tail; each of them waits for the predecessor before its first global access.
    A missing wait would show up as different bits here - eager, no-host-sync, CUDA-graph and with a KV cache attached,
    over repeated runs."""
    from slime_b200.engine import GraphedPrefill
    from slime_b200.synth import synth_inputs

    cfg, eng = _engine(pname)
    n, T = 5, 24
    px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=5, ragged=True)
    grids = [(2, 2)] * B
    ref = eng.prefill(px, ids, mask, grids=grids, want_all_logits=True)
    eng.lib.slime_set_prefill_pdl(1)
    try:
        for _ in range(4):
            res = eng.prefill(px, ids, mask, grids=grids, want_all_logits=True)
            assert res.lengths == ref.lengths
            assert torch.equal(res.logits_last, ref.logits_last) and torch.equal(res.logits_all, ref.logits_all)
        sf = eng.prefill(px, ids, mask, grids=grids, sync_free=True)
        assert torch.equal(sf.logits_last, ref.logits_last)
        g = GraphedPrefill(eng, B, n, T, grids=grids)
        for _ in range(3):
            assert torch.equal(g(px.cuda(), ids.cuda(), mask.cuda()).logits_last, ref.logits_last)
        eng.attach_kv_cache(B, max(ref.lengths) + 4)
        try:
            eng.lib.slime_set_prefill_pdl(0)
            eng.prefill(px, ids, mask, grids=grids)
            a = eng._kv_cache.clone()
            eng._kv_cache.zero_()
            eng.lib.slime_set_prefill_pdl(1)
            eng.prefill(px, ids, mask, grids=grids)
            assert torch.equal(eng._kv_cache, a)
        finally:
            eng.detach_kv_cache()
    finally:
        eng.lib.slime_set_prefill_pdl(-1)
