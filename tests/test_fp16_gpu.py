"""The float16 build (libslime_b200_fp16.so = the same sources with -DSLIME_FP16): the reference's inference dtype
(llava/model/builder.py:43).  Kernels against plain fp32 PyTorch math on the same fp16 inputs - one fp16 rounding of
the result: rel-L2 <= 6e-4 (8x tighter than the bf16 build's 4e-3) - and the tiny model end to end against the
CPU oracle.  Integer results stay exact."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

FP16_REL_L2 = 6e-4


def L():
    from slime_b200 import _lib

    return _lib


def lib():
    return L().load(torch.float16)


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.float16)


def check(out, ref, what, tol=FP16_REL_L2):
    assert torch.isfinite(out.float()).all(), what
    e = rel_l2(out, ref)
    print(f"{what}: rel-L2 {e:.3e}")
    assert e <= tol, f"{what}: rel-L2 {e:.3e} > {tol:.1e}"


def gemm(a, w, bias=None, residual=None, epi=0):
    M, K = a.shape
    N = w.shape[0]
    out = torch.zeros(M, N // 2 if epi == L().EPI_SWIGLU else N, device="cuda", dtype=torch.float16)
    rc = lib().slime_op_gemm(L().ptr(a), a.stride(0), L().ptr(w), w.stride(0), M, N, K, L().ptr(bias), L().ptr(residual),
                             residual.stride(0) if residual is not None else 0, 0, None, epi, L().ptr(out), None,
                             out.stride(0), L().stream_ptr())
    L().check(rc, "op_gemm fp16", lib())
    torch.cuda.synchronize()
    return out


def test_library_reports_its_element_type():
    assert lib().slime_elem_dtype() == 2 and L().load().slime_elem_dtype() == 0
    assert lib() is not L().load()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (100, 136, 72), (577 * 5, 3072, 1024), (1408 * 8, 6144, 4096),
                                   (8, 128256, 4096)])
def test_gemm_fp16(M, N, K):
    torch.manual_seed(M + N + K)
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N)
    check(gemm(a, w, bias=bias), a.float() @ w.float().t() + bias.float(), f"gemm fp16 {M}x{N}x{K}")


def test_gemm_fp16_epilogues():
    torch.manual_seed(2)
    a, w, bias, h = rnd(700, 1024), rnd(4096, 1024, scale=0.05), rnd(4096), rnd(700, 4096)
    pre = a.float() @ w.float().t() + bias.float()
    # tanh.approx / erff: activation error ~1e-3 of the activation's scale, above one fp16 rounding
    check(gemm(a, w, bias=bias, epi=L().EPI_QUICK_GELU), pre * torch.sigmoid(1.702 * pre), "quick_gelu fp16", 2e-3)
    check(gemm(a, w, bias=bias, epi=L().EPI_GELU_ERF), torch.nn.functional.gelu(pre), "gelu fp16")
    check(gemm(a, w, bias=bias, residual=h), pre + h.float(), "residual fp16")
    gate, up = rnd(1024, 512, scale=0.05), rnd(1024, 512, scale=0.05)
    x = rnd(900, 512)
    inter = torch.stack([gate, up], dim=1).reshape(2048, 512).contiguous()
    ref = torch.nn.functional.silu(x.float() @ gate.float().t()) * (x.float() @ up.float().t())
    check(gemm(x, inter, epi=L().EPI_SWIGLU), ref, "swiglu fp16", 2e-3)


def _ref_attention(q, k, v, scale, causal):
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = q.shape[-2], k.shape[-2]
        s = s.masked_fill(~torch.ones(Sq, Sk, device=q.device, dtype=torch.bool).tril(Sk - Sq), float("-inf"))
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("impl", [0], ids=["two_q_tiles"])
@pytest.mark.parametrize("shape", ["clip", "decoder_gqa"])
def test_attention_fp16(impl, shape):
    torch.manual_seed(6)
    if shape == "clip":
        B, h, kvh, d, S, causal = 3, 16, 16, 64, 577, 0
    else:
        B, h, kvh, d, S, causal = 2, 32, 8, 128, 700, 1
    W = (h + 2 * kvh) * d
    qkv = rnd(B * S, W)
    o = torch.zeros(B * S, h * d, device="cuda", dtype=torch.float16)
    rc = lib().slime_op_attention(L().ptr(qkv), L().ptr(qkv[:, h * d:]), L().ptr(qkv[:, (h + kvh) * d:]), L().ptr(o), W, W,
                                  W, h * d, None, None, S, S, S, S, S, B, h, kvh, d, 1.0 / math.sqrt(d), causal,
                                  B * S, B * S, impl, L().stream_ptr())
    L().check(rc, "op_attention fp16", lib())
    torch.cuda.synchronize()
    x = qkv.float().view(B, S, h + 2 * kvh, d)
    q, k, v = x[:, :, :h].transpose(1, 2), x[:, :, h:h + kvh].transpose(1, 2), x[:, :, h + kvh:].transpose(1, 2)
    k, v = k.repeat_interleave(h // kvh, 1), v.repeat_interleave(h // kvh, 1)
    ref = _ref_attention(q, k, v, 1.0 / math.sqrt(d), causal).transpose(1, 2).reshape(B * S, h * d)
    # P is rounded to fp16 before the PV product (as every flash kernel does): ~2 roundings
    check(o, ref, f"attention fp16 {shape} impl {impl}", 1e-3)


def test_norms_fp16():
    torch.manual_seed(7)
    x, w, b = rnd(1000, 1024), rnd(1024), rnd(1024)
    y = torch.empty_like(x)
    L().check(lib().slime_op_layernorm(L().ptr(x), L().ptr(w), L().ptr(b), L().ptr(y), 1000, 1024, 1e-5, L().stream_ptr()),
              "layernorm fp16", lib())
    check(y, torch.nn.functional.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5), "layernorm fp16")
    x, w = rnd(777, 4096), rnd(4096)
    y = torch.empty_like(x)
    L().check(lib().slime_op_rmsnorm(L().ptr(x), L().ptr(w), L().ptr(y), 777, 4096, 1e-5, L().stream_ptr()),
              "rmsnorm fp16", lib())
    xf = x.float()
    ref = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.float16).float()
    check(y, ref, "rmsnorm fp16")


@pytest.mark.parametrize("case", ["spatial", "flat_ragged"])
def test_tiny_model_end_to_end_fp16(case):
    """All stages + logits of the tiny model in fp16 against the CPU oracle (fp32), teacher-forced selection;
    lengths / selection rule exact; decode step consistent with a re-prefill."""
    from oracle import slime_oracle as O
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_inputs, synth_state_dict

    cfg = preset("tiny") if case == "spatial" else preset("tiny", mm_patch_merge_type="flat")
    sd = synth_state_dict(cfg)
    px, ids, mask = synth_inputs(cfg, 2, 5, 20, image_pos=6, ragged=(case != "spatial"))
    grids = [(2, 2)] * 2 if case == "spatial" else None
    with torch.no_grad():
        ora = O.prefill(sd, cfg, px, ids, mask, grids)
    eng = SlimeEngine(cfg, 0, dtype=torch.float16)
    assert eng.dtype == torch.float16
    eng.load_state_dict(sd)
    res = eng.prefill(px, ids, mask, grids=grids, forced_selection=ora["sel"], keep_stages=True)
    assert res.logits_last.dtype == torch.float32 and res.stages["vit"].dtype == torch.float16
    assert res.lengths == ora["lengths"]
    e_vit = rel_l2(res.stages["vit"].cpu(), torch.cat(ora["vit"]))
    last = torch.stack([lg[-1] for lg in ora["logits"]])
    e_log = rel_l2(res.logits_last.cpu(), last)
    print(f"tiny fp16 [{case}]: vit {e_vit:.3e}  logits {e_log:.3e}")
    assert e_vit < 1e-3 and e_log < 1.2e-3
    r2 = eng.prefill(px, ids, mask, grids=grids, want_probs=True, want_last=False, run_decoder=False)
    for b in range(2):
        expect = O.top_p_select(r2.probs[b].cpu(), cfg.mm_resampler_topp)
        assert r2.sel_idx[b, :int(r2.sel_count[b])].cpu().tolist() == expect.tolist()
