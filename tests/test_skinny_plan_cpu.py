"""CPU mirror of the planning logic of the weight-streaming decode GEMM (slime_b200/csrc/gemm_skinny.cu make_plan) and of
the scratch the decode step reserves for it (csrc/api.cu decode_body): for every released model size and every batch the
decode path serves with this kernel (1..32 sequences) each projection of a decoder layer must get a valid plan - k-splits
that are non-empty multiples of 32, staged activations within the shared-memory budget - and the split-K scratch the
stage allocates must cover the plan it will pick (otherwise the kernel would silently fall back to an unsplit or a
tile-kernel launch: a performance cliff, not an error)."""
import pytest

from slime_b200.config import preset

SK_WARPS, SK_KC_MAX, NUM_SMS = 16, 4096, 148


def make_plan(M, N, K, ws_floats, fused_norm=False):
    """mirror of make_plan() for the decode step's calls (no forced splits, aligned operands)"""
    if not (1 <= M <= 32) or K % 32 or N % 8:
        return None
    if fused_norm and N > 8192:
        return None
    mt = 1 if M <= 16 else 2
    kc_max = SK_KC_MAX // mt
    k32, items, warps = K // 32, N // 8, NUM_SMS * SK_WARPS
    s = 1
    while s < 8 and items * s < (warps * 2) // 3 and K // (s * 2) >= 512:
        s *= 2
    while s < 64 and (k32 + s - 1) // s * 32 > kc_max:
        s *= 2
    if s > NUM_SMS:
        return None
    kc32 = (k32 + s - 1) // s
    s = (k32 + kc32 - 1) // kc32
    kc = kc32 * 32
    if kc > kc_max:
        return None
    use_partial = s > 1 or fused_norm
    if use_partial and ws_floats < s * M * N:
        if fused_norm or K > kc_max:
            return None
        s, kc, use_partial = 1, K, False
    return dict(mt=mt, splits=s, kc=kc, use_partial=use_partial, need=s * M * N if use_partial else 0,
                smem=(M if (mt == 1 and M <= 8) else mt * 16) * (kc + (32 if kc % 64 == 0 else 0)) * 2)


def decode_scratch_floats(cfg, B):
    """mirror of decode_body: B * max(8 * max(QKV, H), 4 * max(2 I, V))"""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    qkv = (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * cfg.head_dim
    return B * max(8 * max(qkv, H), 4 * max(2 * I, V))


@pytest.mark.parametrize("pname", ["vicuna-7b", "llama3-8b", "vicuna-13b", "tiny", "small"])
@pytest.mark.parametrize("B", [1, 2, 4, 8, 9, 16, 17, 24, 32])
def test_every_decode_projection_gets_a_plan_that_fits_the_scratch(pname, B):
    cfg = preset(pname)
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    qd = cfg.num_attention_heads * cfg.head_dim
    qkv = qd + 2 * cfg.num_key_value_heads * cfg.head_dim
    ws = decode_scratch_floats(cfg, B)
    calls = [("qkv", qkv, H, False), ("o_proj", H, qd, True), ("gate_up", 2 * I, H, False), ("down", H, I, True),
             ("lm_head", V, H, False)]
    for name, N, K, fused in calls:
        pl = make_plan(B, N, K, ws, fused)
        assert pl is not None, f"{pname} B={B} {name}: no weight-streaming plan"
        assert pl["kc"] % 32 == 0 and pl["kc"] * (pl["splits"] - 1) < K <= pl["kc"] * pl["splits"], (name, pl)
        assert pl["need"] <= ws, f"{pname} B={B} {name}: needs {pl['need']} floats of split-K scratch, stage reserves {ws}"
        # the kernel raises its dynamic shared-memory limit to MT * 16 * (SK_KC_MAX / MT + 32) * 2 bytes
        assert pl["smem"] <= pl["mt"] * 16 * (SK_KC_MAX // pl["mt"] + 32) * 2 <= 227 * 1024, (name, pl)
        # with the scratch the stage reserves the plan is the one the heuristic wants (no silent unsplit fallback)
        assert make_plan(B, N, K, 1 << 60, fused) == pl


def test_headline_model_splits():
    """SliME-Llama3-8B at one sequence: QKV and the two 4096-wide projections are split 4 ways over the GPU, gate/up and
    the LM head have enough rows for one item per warp (profiles/r01_decode_bench.txt)."""
    cfg = preset("llama3-8b")
    ws = decode_scratch_floats(cfg, 1)
    assert make_plan(1, 6144, 4096, ws)["splits"] == 4
    assert make_plan(1, 4096, 4096, ws, True)["splits"] == 4
    assert make_plan(1, 4096, 14336, ws, True)["splits"] == 4
    assert make_plan(1, 28672, 4096, ws)["splits"] == 1
    assert make_plan(1, 128256, 4096, ws)["splits"] == 1
    assert make_plan(1, 4096, 4096, ws)["smem"] == 1 * (1024 + 32) * 2   # one real row, 1024-wide k-split
    assert make_plan(17, 128256, 4096, decode_scratch_floats(cfg, 17))["splits"] == 2  # 32 staged rows: 2048 k each
