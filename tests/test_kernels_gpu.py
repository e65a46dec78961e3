"""Per-kernel parity tests on a B200, called through the C-ABI single-op entry points.

The reference for each floating-point kernel is plain PyTorch fp32 math on the same bf16 inputs
(the kernels accumulate in fp32 and round once to bf16, so they must agree with the fp32 result
to within one bf16 rounding: rel-L2 <= 4e-3 and every element within 2 bf16 ulps of the largest
magnitude).  Integer results (router selection) are compared exactly in test_stages_gpu.py.
"""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_REL_L2 = 4e-3


def _lib():
    from slime_b200 import _lib as L

    return L


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def assert_close_bf16(out, ref, what, tol=BF16_REL_L2):
    ref = ref.float()
    out = out.float()
    assert out.shape == ref.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = rel_l2(out, ref)
    max_abs = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol, f"{what}: rel-L2 {err:.3e} > {tol:.1e} (max abs err {max_abs:.3e}, ref max {scale:.3e})"
    assert max_abs <= 2 ** -6 * scale + 1e-6, f"{what}: max abs err {max_abs:.3e} vs scale {scale:.3e}"


def gemm(a, w, bias=None, residual=None, res_period=0, row_map=None, epi=0, out_rows=None, f32=False):
    L = _lib()
    lib = L.load()
    M, K = a.shape
    N = w.shape[0]
    out_cols = N // 2 if epi == L.EPI_SWIGLU else N
    rows = out_rows if out_rows is not None else M
    if f32:
        out = torch.zeros(rows, out_cols, device="cuda", dtype=torch.float32)
    else:
        out = torch.zeros(rows, out_cols, device="cuda", dtype=torch.bfloat16)
    rc = lib.slime_op_gemm(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), M, N, K, L.ptr(bias), L.ptr(residual),
                           residual.stride(0) if residual is not None else 0, res_period, L.ptr(row_map), epi,
                           None if f32 else L.ptr(out), L.ptr(out) if f32 else None, out.stride(0), L.stream_ptr())
    L.check(rc, "op_gemm")
    torch.cuda.synchronize()
    return out


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64),       # one tile, one k-block
    (128, 256, 128),
    (256, 512, 256),
    (100, 136, 72),       # ragged M, N % 128 != 0, K % 64 != 0 (TMA zero fill + store masks)
    (577 * 5, 1024, 1024),   # CLIP out-proj at 5 crops
    (577 * 5, 3072, 1024),   # CLIP QKV
    (576 * 5, 1024, 640),    # patch embed (K padded 588 -> 640)
    (1408 * 8, 6144, 4096),  # Llama-3 QKV, BLOCK_N = 256 path, many tiles per CTA
    (1408, 4096, 14336),     # Llama-3 down proj (long K)
    (8, 128256, 4096),       # lm_head on 8 last tokens
])
def test_gemm_plain(M, N, K):
    torch.manual_seed(M + N + K)
    a, w = rnd(M, K), rnd(N, K, scale=0.05)
    bias = rnd(N)
    out = gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias.float()
    assert_close_bf16(out, ref, f"gemm {M}x{N}x{K}")


def test_gemm_fp32_out_and_no_bias():
    torch.manual_seed(1)
    a, w = rnd(300, 512), rnd(1000, 512, scale=0.05)
    out = gemm(a, w, f32=True)
    ref = a.float() @ w.float().t()
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("epi", ["quick_gelu", "gelu_erf"])
def test_gemm_activations(epi):
    L = _lib()
    torch.manual_seed(2)
    a, w, bias = rnd(700, 1024), rnd(4096, 1024, scale=0.05), rnd(4096)
    pre = a.float() @ w.float().t() + bias.float()
    if epi == "quick_gelu":
        ref = pre * torch.sigmoid(1.702 * pre)
        out = gemm(a, w, bias=bias, epi=L.EPI_QUICK_GELU)
    else:
        ref = torch.nn.functional.gelu(pre)
        out = gemm(a, w, bias=bias, epi=L.EPI_GELU_ERF)
    assert_close_bf16(out, ref, f"gemm+{epi}")


def test_gemm_residual_inplace_and_periodic():
    torch.manual_seed(3)
    a, w, bias = rnd(577 * 3, 1024), rnd(1024, 1024, scale=0.05), rnd(1024)
    h = rnd(577 * 3, 1024)
    ref = a.float() @ w.float().t() + bias.float() + h.float()
    L = _lib()
    lib = L.load()
    out = h.clone()
    rc = lib.slime_op_gemm(L.ptr(a), 1024, L.ptr(w), 1024, a.shape[0], 1024, 1024, L.ptr(bias), L.ptr(out), 1024, 0,
                           None, 0, L.ptr(out), None, 1024, L.stream_ptr())
    L.check(rc, "gemm residual in place")
    torch.cuda.synchronize()
    assert_close_bf16(out, ref, "gemm+residual (in place)")
    # periodic residual table (Resampler key position term): row r uses table[r % 576]
    a2, w2 = rnd(576 * 4, 1024), rnd(2048, 1024, scale=0.05)
    table = rnd(576, 2048)
    out2 = gemm(a2, w2, bias=None, residual=table, res_period=576)
    ref2 = a2.float() @ w2.float().t() + table.float().repeat(4, 1)
    assert_close_bf16(out2, ref2, "gemm+periodic residual")


def test_gemm_row_map_scatter():
    torch.manual_seed(4)
    a, w = rnd(576, 1024), rnd(4096, 1024, scale=0.05)
    perm = torch.randperm(576, device="cuda").to(torch.int32)
    perm[5] = -1  # dropped row
    out = gemm(a, w, row_map=perm)
    ref = a.float() @ w.float().t()
    expect = torch.zeros_like(ref)
    keep = perm >= 0
    expect[perm[keep].long()] = ref[keep]
    assert_close_bf16(out, expect, "gemm+row_map")


def test_gemm_swiglu():
    L = _lib()
    torch.manual_seed(5)
    I, H, M = 1024, 512, 900
    a = rnd(M, H)
    gate, up = rnd(I, H, scale=0.05), rnd(I, H, scale=0.05)
    inter = torch.stack([gate, up], dim=1).reshape(2 * I, H).contiguous()  # rows (g0,u0,g1,u1,...)
    out = gemm(a, inter, epi=L.EPI_SWIGLU)
    ref = torch.nn.functional.silu(a.float() @ gate.float().t()) * (a.float() @ up.float().t())
    assert_close_bf16(out, ref, "gemm+swiglu")


def attention(q, k, v, o_rows, o_ld, **kw):
    L = _lib()
    lib = L.load()
    o = torch.zeros(o_rows, o_ld, device="cuda", dtype=torch.bfloat16)
    rc = lib.slime_op_attention(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), kw["q_ld"], kw["k_ld"], kw["v_ld"], o_ld,
                                L.ptr(kw.get("cu_q")), L.ptr(kw.get("cu_k")), kw["seqlen_q"], kw["seqlen_k"],
                                kw.get("q_batch_rows", 0), kw.get("k_batch_rows", 0), kw.get("o_batch_rows", 0),
                                kw["batch"], kw["heads"], kw["kv_heads"], kw["head_dim"], kw["scale"],
                                kw.get("causal", 0), kw.get("total_q_rows", 0), kw.get("total_k_rows", 0),
                                kw.get("impl", 0), L.stream_ptr())
    L.check(rc, "op_attention")
    torch.cuda.synchronize()
    return o


def ref_attention(q, k, v, scale, causal):
    # q [B,h,Sq,d], k/v [B,h,Sk,d] fp32
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = q.shape[-2], k.shape[-2]
        mask = torch.ones(Sq, Sk, device=q.device, dtype=torch.bool).tril(Sk - Sq)
        s = s.masked_fill(~mask, float("-inf"))
    return torch.softmax(s, dim=-1) @ v


IMPLS = [pytest.param(0, id="two_q_tiles")]  # one prefill attention kernel (attention_tc2.cu)


@pytest.mark.parametrize("impl", IMPLS)
def test_attention_clip_shape(impl):
    """16 heads x 64, S = 577 (4*128+65: ragged tail tile), non-causal, packed qkv rows."""
    torch.manual_seed(6)
    B, h, d, S = 3, 16, 64, 577
    D = h * d
    qkv = rnd(B * S, 3 * D)
    o = attention(qkv, qkv[:, D:], qkv[:, 2 * D:], B * S, D, q_ld=3 * D, k_ld=3 * D, v_ld=3 * D, seqlen_q=S,
                  seqlen_k=S, q_batch_rows=S, k_batch_rows=S, o_batch_rows=S, batch=B, heads=h, kv_heads=h,
                  head_dim=d, scale=d ** -0.5, impl=impl)
    x = qkv.float().view(B, S, 3, h, d).permute(2, 0, 3, 1, 4)
    ref = ref_attention(x[0], x[1], x[2], d ** -0.5, False).permute(0, 2, 1, 3).reshape(B * S, D)
    assert_close_bf16(o, ref, "attention clip")


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("nq", [144, 576])
def test_attention_resampler_shape(nq, impl):
    """8 heads x 128, shared learned queries (q_batch_rows = 0) against 576 keys per crop."""
    torch.manual_seed(7)
    n, h, d, NK = 4, 8, 128, 576
    D = h * d
    q = rnd(nq, D)
    kv = rnd(n * NK, 2 * D)
    o = attention(q, kv, kv[:, D:], n * nq, D, q_ld=D, k_ld=2 * D, v_ld=2 * D, seqlen_q=nq, seqlen_k=NK,
                  q_batch_rows=0, k_batch_rows=NK, o_batch_rows=nq, batch=n, heads=h, kv_heads=h, head_dim=d,
                  scale=d ** -0.5, impl=impl)
    qf = q.float().view(1, nq, h, d).permute(0, 2, 1, 3).expand(n, h, nq, d)
    kf = kv[:, :D].float().view(n, NK, h, d).permute(0, 2, 1, 3)
    vf = kv[:, D:].float().view(n, NK, h, d).permute(0, 2, 1, 3)
    ref = ref_attention(qf, kf, vf, d ** -0.5, False).permute(0, 2, 1, 3).reshape(n * nq, D)
    assert_close_bf16(o, ref, f"attention resampler nq={nq}")


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("h,kvh", [(8, 2), (4, 4), (6, 2)], ids=["gqa4", "mha", "gqa3"])
def test_attention_decoder_causal_varlen_gqa(impl, h, kvh):
    """Packed variable-length causal attention, heads x 128: GQA group 4 (Llama-3 layout: the two-tile kernel pairs q
    heads of a group), MHA (Vicuna: it pairs consecutive query tiles, the earlier one sees its last kv tile fully
    masked) and an odd group size."""
    torch.manual_seed(8)
    d = 128
    lens = [1, 63, 64, 65, 127, 128, 129, 200, 333, 1400]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device="cuda", dtype=torch.int32)
    total = sum(lens)
    W = (h + 2 * kvh) * d
    qkv = rnd(total, W)
    o = attention(qkv, qkv[:, h * d:], qkv[:, (h + kvh) * d:], total, h * d, q_ld=W, k_ld=W, v_ld=W, cu_q=cu,
                  cu_k=cu, seqlen_q=max(lens), seqlen_k=max(lens), batch=len(lens), heads=h, kv_heads=kvh,
                  head_dim=d, scale=d ** -0.5, causal=1, total_q_rows=total, total_k_rows=total, impl=impl)
    off = 0
    for L_ in lens:
        blk = qkv[off:off + L_].float()
        q = blk[:, :h * d].view(L_, h, d).permute(1, 0, 2)[None]
        k = blk[:, h * d:(h + kvh) * d].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
        v = blk[:, (h + kvh) * d:].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
        ref = ref_attention(q, k, v, d ** -0.5, True)[0].permute(1, 0, 2).reshape(L_, h * d)
        assert_close_bf16(o[off:off + L_], ref, f"attention causal len={L_}")
        off += L_


@pytest.mark.timeout(180)
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("causal,d", [(1, 128), (0, 128), (0, 64), (1, 64)])
def test_attention_many_ragged_items(impl, causal, d):
    """Scheduling stress for the persistent kernel: 96 sequences of random length 0..400 x 8 heads = hundreds of
    work items per CTA, most of them one or two kv tiles long (the item hand-over, the Q / O double buffers and the
    deferred epilogue are exercised on every boundary), empty sequences in the list, GQA when causal."""
    g = torch.Generator().manual_seed(10 + d + causal)
    h, kvh = (8, 2) if causal else (8, 8)
    lens = torch.randint(0, 401, (96,), generator=g).tolist()
    lens[3] = 0
    lens[50] = 0
    lens[-1] = 1
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device="cuda", dtype=torch.int32)
    total = sum(lens)
    W = (h + 2 * kvh) * d
    torch.manual_seed(11)
    qkv = rnd(total, W)
    for rep in range(3):  # back-to-back launches: barrier state / TMEM are per launch, results must not change
        o = attention(qkv, qkv[:, h * d:], qkv[:, (h + kvh) * d:], total, h * d, q_ld=W, k_ld=W, v_ld=W, cu_q=cu,
                      cu_k=cu, seqlen_q=max(lens), seqlen_k=max(lens), batch=len(lens), heads=h, kv_heads=kvh,
                      head_dim=d, scale=d ** -0.5, causal=causal, total_q_rows=total, total_k_rows=total, impl=impl)
        if rep == 0:
            first = o.clone()
        else:
            assert torch.equal(o, first), "attention is not deterministic across launches"
    off = 0
    worst = 0.0
    for L_ in lens:
        if L_ == 0:
            continue
        blk = qkv[off:off + L_].float()
        q = blk[:, :h * d].view(L_, h, d).permute(1, 0, 2)[None]
        k = blk[:, h * d:(h + kvh) * d].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
        v = blk[:, (h + kvh) * d:].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
        ref = ref_attention(q, k, v, d ** -0.5, bool(causal))[0].permute(1, 0, 2).reshape(L_, h * d)
        worst = max(worst, rel_l2(o[off:off + L_], ref))
        off += L_
    assert worst <= BF16_REL_L2, f"worst per-sequence rel-L2 {worst:.3e}"


@pytest.mark.timeout(180)
@pytest.mark.parametrize("impl", IMPLS)
def test_attention_cross_lengths_causal_offset(impl):
    """Sq != Sk with the causal diagonal anchored at the END of the keys (query i sees keys <= i + Sk - Sq), several
    batches, so a CTA walks items with different tile counts."""
    torch.manual_seed(12)
    B, h, d, Sq, Sk = 5, 4, 128, 200, 520
    q, kv = rnd(B * Sq, h * d), rnd(B * Sk, 2 * h * d)
    o = attention(q, kv, kv[:, h * d:], B * Sq, h * d, q_ld=h * d, k_ld=2 * h * d, v_ld=2 * h * d, seqlen_q=Sq,
                  seqlen_k=Sk, q_batch_rows=Sq, k_batch_rows=Sk, o_batch_rows=Sq, batch=B, heads=h, kv_heads=h,
                  head_dim=d, scale=d ** -0.5, causal=1, impl=impl)
    qf = q.float().view(B, Sq, h, d).permute(0, 2, 1, 3)
    kf = kv[:, :h * d].float().view(B, Sk, h, d).permute(0, 2, 1, 3)
    vf = kv[:, h * d:].float().view(B, Sk, h, d).permute(0, 2, 1, 3)
    ref = ref_attention(qf, kf, vf, d ** -0.5, True).permute(0, 2, 1, 3).reshape(B * Sq, h * d)
    assert_close_bf16(o, ref, "attention Sq != Sk causal")


@pytest.mark.parametrize("D", [128, 1024, 4096, 5120])
def test_layernorm_rmsnorm(D):
    L = _lib()
    lib = L.load()
    torch.manual_seed(9)
    rows = 517
    x, w, b = rnd(rows, D), rnd(D), rnd(D)
    y = torch.empty_like(x)
    L.check(lib.slime_op_layernorm(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), rows, D, 1e-5, L.stream_ptr()), "ln")
    ref = torch.nn.functional.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5)
    torch.cuda.synchronize()
    assert_close_bf16(y, ref, f"layernorm D={D}")
    L.check(lib.slime_op_rmsnorm(L.ptr(x), L.ptr(w), L.ptr(y), rows, D, 1e-5, L.stream_ptr()), "rms")
    torch.cuda.synchronize()
    xf = x.float()
    normed = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16)
    ref = (w * normed).float()  # HF: weight * hidden_states.to(input_dtype), product in bf16
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-6
    assert rel_l2(y, ref) < 3e-3


@pytest.fixture
def force_2cta():
    L = _lib()
    lib = L.load()
    lib.slime_gemm_set_2cta_mode(1)
    yield
    lib.slime_gemm_set_2cta_mode(-1)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 512), (300, 520, 200), (577 * 5, 3072, 1024),
                                   (1408 * 8, 6144, 4096), (1408 * 4, 4096, 14336)])
def test_gemm_2cta_plain(force_2cta, M, N, K):
    """cta_group::2 kernel (256 x 256 cluster tiles): same results as the fp32 reference."""
    torch.manual_seed(M + N + K)
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N)
    out = gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias.float()
    assert_close_bf16(out, ref, f"gemm 2cta {M}x{N}x{K}")


def test_gemm_2cta_epilogues(force_2cta):
    L = _lib()
    torch.manual_seed(12)
    a, h = rnd(1000, 1024), rnd(1000, 1024)
    w, bias = rnd(1024, 1024, scale=0.05), rnd(1024)
    out = gemm(a, w, bias=bias, residual=h)
    assert_close_bf16(out, a.float() @ w.float().t() + bias.float() + h.float(), "gemm 2cta + residual")
    gate, up = rnd(1024, 512, scale=0.05), rnd(1024, 512, scale=0.05)
    x = rnd(900, 512)
    inter = torch.stack([gate, up], dim=1).reshape(2048, 512).contiguous()
    out = gemm(x, inter, epi=L.EPI_SWIGLU)
    ref = torch.nn.functional.silu(x.float() @ gate.float().t()) * (x.float() @ up.float().t())
    assert_close_bf16(out, ref, "gemm 2cta + swiglu")


@pytest.mark.parametrize("epi_mode", [0, 1])
def test_gemm_2cta_tile_width_bit_identical(epi_mode):
    """192-column cluster tiles (small problems: fewer width-weighted waves) against the 256-column throughput shape: every
    output element accumulates over k in the same order, so all epilogues must give the same bits - ragged M, N not a
    multiple of either width, row map, periodic residual, fp32 out, SwiGLU; also against fp32 math."""
    L = _lib()
    lib = L.load()
    torch.manual_seed(9)
    M, N, K = 1379, 4096, 512
    a, w, bias, h = rnd(M, K), rnd(N, K, scale=0.05), rnd(N), rnd(M, N)
    a2, w2, b2 = rnd(1000, 256), rnd(1064, 256, scale=0.05), rnd(1064)
    table = rnd(37, 1064)
    row_map = torch.randperm(1000, device="cuda").to(torch.int32)
    row_map[::7] = -1
    gu = rnd(2048, 256, scale=0.05)
    res = {}
    lib.slime_gemm_set_2cta_mode(1)
    L.check(lib.slime_gemm_set_epi_mode(epi_mode), "set_epi_mode")
    try:
        for bn in (256, 192):
            L.check(lib.slime_gemm_set_tile_n(bn), "set_tile_n")
            res[bn] = [
                gemm(a, w, bias=bias),
                gemm(a, w, bias=bias, residual=h),
                gemm(a2, w2, residual=table, res_period=37),
                gemm(a2, w2, bias=b2, epi=L.EPI_QUICK_GELU),
                gemm(a2, w2, bias=b2, epi=L.EPI_GELU_ERF),
                gemm(a2, w2, bias=b2, row_map=row_map),
                gemm(a2, gu, epi=L.EPI_SWIGLU),
                gemm(a2, w2, bias=b2, f32=True),
            ]
    finally:
        lib.slime_gemm_set_tile_n(-1)
        lib.slime_gemm_set_epi_mode(-1)
        lib.slime_gemm_set_2cta_mode(-1)
    for i, (x, y) in enumerate(zip(res[256], res[192])):
        assert torch.equal(x, y), f"192-wide tiles differ from 256-wide in case {i} (epi_mode={epi_mode})"
    assert_close_bf16(res[192][1], a.float() @ w.float().t() + bias.float() + h.float(), "gemm 192-wide + residual")
    assert lib.slime_gemm_set_tile_n(7) < 0  # only -1, 0, 192, 256


# ------------------------------------------------------------------------------------------------
# opt-in / alternative code paths: staged GEMM epilogue, softmax variants, RoPE in the QKV epilogue
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("two_cta", [0, 1])
def test_gemm_staged_epilogue_bit_identical(two_cta):
    """GemmParams::epi_mode 1 (chunks transposed through shared memory, coalesced stores and residual loads) only
    changes the HBM access pattern: every epilogue must produce the same bits as the direct path, including ragged
    M, N not a multiple of 32, the row-scatter map with dropped rows and the row-periodic residual table."""
    L = _lib()
    lib = L.load()
    torch.manual_seed(5)
    M, N, K = 1000, 1064, 256
    a, w, bias, h = rnd(M, K), rnd(N, K, scale=0.05), rnd(N), rnd(M, N)
    table = rnd(37, N)
    row_map = torch.randperm(M, device="cuda").to(torch.int32)
    row_map[::7] = -1
    gu = rnd(2048, K, scale=0.05)
    res = {}
    lib.slime_gemm_set_2cta_mode(two_cta)
    try:
        for mode in (0, 1):
            L.check(lib.slime_gemm_set_epi_mode(mode), "set_epi_mode")
            res[mode] = [
                gemm(a, w, bias=bias),
                gemm(a, w, bias=bias, residual=h),
                gemm(a, w, residual=table, res_period=37),
                gemm(a, w, bias=bias, epi=L.EPI_QUICK_GELU),
                gemm(a, w, bias=bias, epi=L.EPI_GELU_ERF),
                gemm(a, w, bias=bias, row_map=row_map),
                gemm(a, gu, epi=L.EPI_SWIGLU),
                gemm(a, w, bias=bias, f32=True),
            ]
    finally:
        lib.slime_gemm_set_epi_mode(0)
        lib.slime_gemm_set_2cta_mode(-1)
    for i, (x, y) in enumerate(zip(res[0], res[1])):
        assert torch.equal(x, y), f"staged epilogue differs from direct in case {i} (2cta={two_cta})"
    assert_close_bf16(res[1][1], a.float() @ w.float().t() + bias.float() + h.float(), "staged gemm + residual")


@pytest.mark.parametrize("variant", [0, 2, 3, 4])
@pytest.mark.parametrize("causal,d", [(1, 128), (0, 64)])
def test_attention_softmax_variants(variant, causal, d):
    """Shares of polynomial exp2 (0 / 2 / 3 / 4 of every 8 column pairs on the FMA pipe instead of MUFU.EX2) against fp32
    math, incl. peaked scores (large logits: very negative exponents, lazy rescale) and ragged lengths (masked chunks keep
    the scalar path)."""
    L = _lib()
    lib = L.load()
    torch.manual_seed(variant * 10 + d)
    lens = [700, 130, 1, 577, 129]
    h, kvh = (8, 2) if d == 128 else (4, 4)
    total = sum(lens)
    W = (h + 2 * kvh) * d
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device="cuda", dtype=torch.int32)
    try:
        for scale in (1.0, 3.0):
            qkv = (torch.randn(total, W, device="cuda") * scale).to(torch.bfloat16)
            o = torch.zeros(total, h * d, device="cuda", dtype=torch.bfloat16)
            L.check(lib.slime_attention_set_poly(variant), "set_poly")
            rc = lib.slime_op_attention(L.ptr(qkv), L.ptr(qkv[:, h * d:]), L.ptr(qkv[:, (h + kvh) * d:]), L.ptr(o), W, W, W,
                                        h * d, L.ptr(cu), L.ptr(cu), max(lens), max(lens), 0, 0, 0, len(lens), h, kvh, d,
                                        d ** -0.5, causal, total, total, 0, L.stream_ptr())
            L.check(rc, "op_attention")
            torch.cuda.synchronize()
            ref = torch.zeros(total, h * d, device="cuda")
            for b, n in enumerate(lens):
                r0 = int(cu[b])
                q = qkv[r0:r0 + n, :h * d].float().reshape(n, h, d).transpose(0, 1)
                k = qkv[r0:r0 + n, h * d:(h + kvh) * d].float().reshape(n, kvh, d).transpose(0, 1)
                v = qkv[r0:r0 + n, (h + kvh) * d:].float().reshape(n, kvh, d).transpose(0, 1)
                k, v = k.repeat_interleave(h // kvh, 0), v.repeat_interleave(h // kvh, 0)
                s = (q @ k.transpose(1, 2)) * d ** -0.5
                if causal:
                    s = s.masked_fill(torch.triu(torch.ones(n, n, device="cuda", dtype=torch.bool), 1), float("-inf"))
                ref[r0:r0 + n] = (s.softmax(-1) @ v).transpose(0, 1).reshape(n, h * d)
            assert_close_bf16(o, ref, f"attention variant {variant} causal={causal} d={d} scale={scale}", tol=6e-3)
    finally:
        lib.slime_attention_set_poly(-1)


@pytest.mark.parametrize("rows,two_cta", [(1, -1), (8, -1), (700, -1), (700, 1), (333, 1)])
def test_qkv_rope_fused_epilogue(rows, two_cta):
    """RoPE fused into the QKV GEMM epilogue (SLIME_FLAG_ROPE_INTERLEAVED: q / k weight rows of every head interleaved
    at load) against fp32 math of HF apply_rotary_pos_emb, and against the unfused path (GEMM + in-place pass) after
    undoing the feature permutation.  Covers the 1-CTA kernels (decode-sized M) and the 2-CTA kernel."""
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.weights import rope_interleave_rows

    L = _lib()
    cfg = preset("small")
    hd, nh, nkv, H = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.hidden_size
    torch.manual_seed(rows)
    x = rnd(rows, H)
    wq, wk, wv = rnd(nh * hd, H, scale=0.05), rnd(nkv * hd, H, scale=0.05), rnd(nkv * hd, H, scale=0.05)
    pos = torch.randint(0, 900, (rows,), device="cuda", dtype=torch.int32)
    outs = {}
    L.load().slime_gemm_set_2cta_mode(two_cta)
    for fused in (False, True):
        eng = SlimeEngine(cfg, 0, max_pos=1024, fused_rope=fused)
        perm = (lambda w: rope_interleave_rows(w, hd)) if fused else (lambda w: w)
        w = torch.cat([perm(wq), perm(wk), wv]).contiguous()
        out = torch.zeros(rows, (nh + 2 * nkv) * hd, device="cuda", dtype=torch.bfloat16)
        L.check(eng.lib.slime_op_qkv_rope(eng._ctx, L.ptr(x), L.ptr(w), rows, L.ptr(pos), L.ptr(out), L.stream_ptr()),
                "op_qkv_rope")
        torch.cuda.synchronize()
        if fused:  # undo the interleave of the q / k feature order: column 2i <- feature i, 2i+1 <- feature i + hd/2
            qk = out[:, :(nh + nkv) * hd].reshape(rows, nh + nkv, hd // 2, 2).transpose(2, 3).reshape(rows, -1)
            out = torch.cat([qk, out[:, (nh + nkv) * hd:]], dim=1)
        outs[fused] = out
        eng.close()
    L.load().slime_gemm_set_2cta_mode(-1)
    # fp32 reference (HF llama/modeling_llama.py:152-176)
    y = x.float() @ torch.cat([wq, wk, wv]).float().t()
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, device="cuda", dtype=torch.float32) / hd))
    ang = pos.float()[:, None] * inv[None, :]
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1), torch.cat([ang.sin(), ang.sin()], -1)
    qk = y[:, :(nh + nkv) * hd].reshape(rows, nh + nkv, hd)
    rot = torch.cat([-qk[..., hd // 2:], qk[..., :hd // 2]], -1)
    ref = torch.cat([(qk * cos[:, None, :] + rot * sin[:, None, :]).reshape(rows, -1), y[:, (nh + nkv) * hd:]], dim=1)
    assert_close_bf16(outs[True], ref, f"fused qkv+rope rows={rows}")
    assert_close_bf16(outs[False], ref, f"unfused qkv+rope rows={rows}", tol=6e-3)
    assert torch.equal(outs[True][:, (nh + nkv) * hd:], outs[False][:, (nh + nkv) * hd:]), "v columns must not change"


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("causal,d", [(0, 128), (1, 128), (0, 64)])
def test_attention_scores_growing_along_the_keys(impl, causal, d):
    """Scores that grow by hundreds of nats from the first key tile to the last one: every tile moves the reference max
    by far more than the lazy-rescale threshold (2^8), so O and the row sums are rescaled tile after tile - on the hd-64
    path while the next score tile is already being computed.  Must match the fp32 softmax (which concentrates on the
    last visible keys)."""
    torch.manual_seed(13)
    B, h, S = 2, 4, 700
    q = (torch.rand(B * S, h * d, device="cuda") * 0.5 + 0.75).to(torch.bfloat16)            # positive, O(1)
    ramp = torch.arange(S, device="cuda", dtype=torch.float32).repeat(B)[:, None] / S          # 0 .. 1 along the keys
    k = (ramp * 3.0 * torch.ones(1, h * d, device="cuda")).to(torch.bfloat16)                  # q.k grows to ~ 3 d
    v = rnd(B * S, h * d)
    scale = d ** -0.5 * 8.0
    o = attention(q, k, v, B * S, h * d, q_ld=h * d, k_ld=h * d, v_ld=h * d, seqlen_q=S, seqlen_k=S, q_batch_rows=S,
                  k_batch_rows=S, o_batch_rows=S, batch=B, heads=h, kv_heads=h, head_dim=d, scale=scale,
                  causal=causal, impl=impl)
    qf = q.float().view(B, S, h, d).permute(0, 2, 1, 3)
    kf = k.float().view(B, S, h, d).permute(0, 2, 1, 3)
    vf = v.float().view(B, S, h, d).permute(0, 2, 1, 3)
    ref = ref_attention(qf, kf, vf, scale, bool(causal)).permute(0, 2, 1, 3).reshape(B * S, h * d)
    assert_close_bf16(o, ref, f"attention growing scores causal={causal} d={d}")
