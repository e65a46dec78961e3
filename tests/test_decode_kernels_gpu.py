"""Decode-step kernels through the C-ABI single-op entry points: the weight-streaming GEMM for
M <= 32 rows (csrc/gemm_skinny.cu: k-splits, every epilogue, fused RMSNorm) and the split-KV single-query attention
(csrc/decode_attn.cu), against fp32 PyTorch math on the same bf16 inputs, against the tcgen05 / one-CTA-per-head
kernels they replace, and for run-to-run determinism (fixed-order split reductions, no atomics)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _L():
    from slime_b200 import _lib as L

    return L


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def skinny(a, w, splits=1, bias=None, residual=None, epi=0, f32=False, out=None, norm_w=None, eps=1e-5, rope=None,
           counters=None, a_norm_w=None):
    """rope = (pos int32 [M], table fp32 [max_pos, half, 2], half, cols); counters (int32 [N / 8], zero): split-K sum
    finished inside the kernel; a_norm_w [K]: RMSNorm(eps) of the rows of a while they are staged"""
    L = _L()
    lib = L.load()
    M, K = a.shape
    N = w.shape[0]
    cols = N // 2 if epi == L.EPI_SWIGLU else N
    if out is None:
        out = torch.zeros(M, cols, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    ws = torch.empty(max(1, splits * M * N), device="cuda", dtype=torch.float32)
    norm_out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if norm_w is not None else None
    pos, table, half, rcols = rope if rope is not None else (None, None, 0, 0)
    args = (L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), M, N, K, L.ptr(bias), L.ptr(residual),
            residual.stride(0) if residual is not None else 0, epi,
            None if f32 else L.ptr(out), L.ptr(out) if f32 else None, out.stride(0), splits,
            L.ptr(ws), ws.numel(), L.ptr(norm_w), L.ptr(norm_out), eps, L.ptr(pos), L.ptr(table),
            half, rcols, table.shape[0] if table is not None else 0)
    if counters is None and a_norm_w is None:
        rc = lib.slime_op_gemm_skinny(*args, L.stream_ptr())
    else:
        rc = lib.slime_op_gemm_skinny_fused(*args, L.ptr(counters), L.ptr(a_norm_w), eps, L.stream_ptr())
    L.check(rc, "op_gemm_skinny")
    torch.cuda.synchronize()
    return (out, norm_out) if norm_w is not None else out


@pytest.mark.parametrize("M", [1, 5, 8, 9, 16, 17, 32])
@pytest.mark.parametrize("N,K,splits", [
    (4096, 4096, 1), (4096, 4096, 4), (6144, 4096, 2), (4096, 14336, 4), (4096, 14336, 8),
    (1024, 512, 1), (136, 96, 3), (28672, 4096, 2), (5120, 13824, 8),
])
def test_skinny_gemm_matches_fp32(M, N, K, splits):
    if M > 16 and K // splits > 2048:
        pytest.skip("M > 16 stages at most 2048 k per split")
    torch.manual_seed(M * 7 + N + K + splits)
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N)
    out = skinny(a, w, splits=splits, bias=bias)
    ref = a.float() @ w.float().t() + bias.float()
    assert torch.isfinite(out.float()).all()
    e = rel_l2(out, ref)
    assert e < 4e-3, f"skinny gemm M={M} {N}x{K} splits={splits}: rel-L2 {e:.3e}"
    out32 = skinny(a, w, splits=splits, f32=True)
    e32 = rel_l2(out32, a.float() @ w.float().t())
    assert e32 < 2e-5, f"fp32 out: {e32:.3e}"


def test_skinny_gemm_vs_tcgen05_and_determinism():
    """The library's own dispatch (slime_op_gemm): M <= 32 goes to the weight-streaming kernel, the tcgen05 kernel
    computes the same thing; repeated runs are bit-identical."""
    L = _L()
    lib = L.load()
    torch.manual_seed(11)
    a, w, bias, res = rnd(16, 4096), rnd(6144, 4096, scale=0.05), rnd(6144), rnd(16, 6144)

    def run():
        out = torch.zeros(16, 6144, device="cuda", dtype=torch.bfloat16)
        L.check(lib.slime_op_gemm(L.ptr(a), 4096, L.ptr(w), 4096, 16, 6144, 4096, L.ptr(bias), L.ptr(res), 6144, 0, None,
                                  0, L.ptr(out), None, 6144, L.stream_ptr()), "op_gemm")
        torch.cuda.synchronize()
        return out

    n0 = lib.slime_launch_count()
    o1, o2 = run(), run()
    assert torch.equal(o1, o2)
    try:
        lib.slime_gemm_set_skinny_mode(0)
        o_tc = run()
    finally:
        lib.slime_gemm_set_skinny_mode(-1)
    ref = a.float() @ w.float().t() + bias.float() + res.float()
    assert rel_l2(o1, ref) < 4e-3 and rel_l2(o_tc, ref) < 4e-3
    assert rel_l2(o1, o_tc) < 3e-3
    s1, s2 = skinny(a, w, splits=4, bias=bias, residual=res), skinny(a, w, splits=4, bias=bias, residual=res)
    assert torch.equal(s1, s2), "split-K reduction must be deterministic"
    assert lib.slime_launch_count() > n0


@pytest.mark.parametrize("splits", [1, 4])
def test_skinny_gemm_epilogues(splits):
    L = _L()
    torch.manual_seed(5 + splits)
    M, H, I = 13, 1024, 2048
    a = rnd(M, H)
    gate, up = rnd(I, H, scale=0.05), rnd(I, H, scale=0.05)
    inter = torch.stack([gate, up], dim=1).reshape(2 * I, H).contiguous()  # rows (g0,u0,g1,u1,...)
    out = skinny(a, inter, splits=splits, epi=L.EPI_SWIGLU)
    ref = torch.nn.functional.silu(a.float() @ gate.float().t()) * (a.float() @ up.float().t())
    assert rel_l2(out, ref) < 4e-3
    w, bias = rnd(I, H, scale=0.05), rnd(I)
    pre = a.float() @ w.float().t() + bias.float()
    assert rel_l2(skinny(a, w, splits=splits, bias=bias, epi=L.EPI_GELU_ERF), torch.nn.functional.gelu(pre)) < 4e-3
    assert rel_l2(skinny(a, w, splits=splits, bias=bias, epi=L.EPI_QUICK_GELU), pre * torch.sigmoid(1.702 * pre)) < 4e-3
    # residual in place (out aliases the residual)
    h = rnd(M, I)
    ref_h = pre + h.float()
    out_h = h.clone()
    skinny(a, w, splits=splits, bias=bias, residual=out_h, out=out_h)
    assert rel_l2(out_h, ref_h) < 4e-3


@pytest.mark.parametrize("M,splits", [(1, 1), (7, 4), (16, 8), (24, 8)])
def test_skinny_gemm_fused_rmsnorm(M, splits):
    """out = a W^T + residual (in place) and norm_out = w_n * bf16(out * rsqrt(mean(out^2) + eps)) from ONE finishing
    kernel: out must equal the unfused result bit for bit, norm_out the RMSNorm of that bf16 row."""
    torch.manual_seed(M + splits)
    H, K = 4096, 4096 if M <= 16 else 2048 * 4
    a, w, nw = rnd(M, K), rnd(H, K, scale=0.02), (1.0 + 0.1 * torch.randn(H, device="cuda")).to(torch.bfloat16)
    h = rnd(M, H)
    plain = h.clone()
    skinny(a, w, splits=splits, residual=plain, out=plain)
    fused = h.clone()
    _, normed = skinny(a, w, splits=splits, residual=fused, out=fused, norm_w=nw, eps=1e-5)
    assert torch.equal(fused, plain)
    x = fused.float()
    ref = nw.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16).float()
    assert rel_l2(normed, ref) < 3e-3
    # and against the standalone rmsnorm kernel: same rounding recipe, reduction order may differ in the last bit
    L = _L()
    y = torch.zeros_like(fused)
    L.check(L.load().slime_op_rmsnorm(L.ptr(fused), L.ptr(nw), L.ptr(y), M, H, 1e-5, L.stream_ptr()), "rmsnorm")
    torch.cuda.synchronize()
    assert rel_l2(normed, y) < 2e-3


@pytest.mark.parametrize("M,splits", [(3, 1), (16, 2), (20, 4)])
def test_skinny_gemm_rope_epilogue(M, splits):
    """Rotary embedding on interleaved (i, i + hd/2) column pairs of the q / k heads (same convention as the tcgen05
    GEMM's GEMM_EPI_ROPE), v columns untouched."""
    L = _L()
    torch.manual_seed(M)
    hd, nh, nkv, H, max_pos = 128, 4, 2, 1024, 512
    half = hd // 2
    N = (nh + 2 * nkv) * hd
    a, w = rnd(M, H), rnd(N, H, scale=0.05)
    pos = torch.randint(0, max_pos, (M,), device="cuda", dtype=torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda", dtype=torch.float32) / hd))
    ang = torch.arange(max_pos, device="cuda", dtype=torch.float32)[:, None] * inv[None]
    table = torch.stack([ang.cos(), ang.sin()], -1).contiguous()  # [max_pos, half, 2]
    rcols = (nh + nkv) * hd
    out = skinny(a, w, splits=splits, epi=L.EPI_ROPE, rope=(pos, table, half, rcols))
    y = a.float() @ w.float().t()
    qk = y[:, :rcols].reshape(M, nh + nkv, half, 2)
    c, s = table[pos.long(), :, 0][:, None, :], table[pos.long(), :, 1][:, None, :]
    lo, hi = qk[..., 0], qk[..., 1]
    rot = torch.stack([lo * c - hi * s, hi * c + lo * s], -1).reshape(M, rcols)
    ref = torch.cat([rot, y[:, rcols:]], 1)
    assert rel_l2(out, ref) < 4e-3


def decode_attention(q, kc, vc, lens, heads, kv_heads, splits, counters=None):
    L = _L()
    lib = L.load()
    B, cache_len = kc.shape[0], kc.shape[1]
    hd = 128
    out = torch.zeros(B, heads * hd, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(max(1, B * heads * max(splits, 1) * (hd + 2)), device="cuda", dtype=torch.float32)
    rc = lib.slime_op_decode_attention_fused(L.ptr(q), q.stride(0), L.ptr(kc), L.ptr(vc), cache_len, L.ptr(lens), B, heads,
                                             kv_heads, hd, 1.0 / math.sqrt(hd), L.ptr(out), out.stride(0), splits,
                                             L.ptr(ws), L.ptr(counters), L.stream_ptr())
    L.check(rc, "op_decode_attention")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,splits", [(1, 4096, 4096, 4), (16, 6144, 4096, 2), (7, 4096, 14336, 8), (24, 4096, 8192, 8),
                                          (5, 136, 96, 3), (32, 28672, 4096, 2)])
def test_skinny_gemm_in_kernel_finish(M, N, K, splits):
    """Split-K finished inside the kernel (atomic ticket per 8-column tile, the last warp adds all partials in split
    order): bit-identical with the finishing-kernel path for every epilogue, counters back at zero, repeatable."""
    L = _L()
    torch.manual_seed(M + N + splits)
    a, w, bias, res = rnd(M, K), rnd(N, K, scale=0.05), rnd(N), rnd(M, N)
    cnt = torch.zeros(N // 8, device="cuda", dtype=torch.int32)
    for kw in (dict(bias=bias), dict(bias=bias, residual=res), dict(f32=True), dict(bias=bias, epi=L.EPI_GELU_ERF)):
        plain = skinny(a, w, splits=splits, **kw)
        for _ in range(2):
            fused = skinny(a, w, splits=splits, counters=cnt, **kw)
            assert torch.equal(plain, fused), kw.keys()
            assert int(cnt.abs().sum()) == 0, "tickets must be handed back"
    if N % 16 == 0:
        assert torch.equal(skinny(a, w, splits=splits, epi=L.EPI_SWIGLU), skinny(a, w, splits=splits, epi=L.EPI_SWIGLU, counters=cnt))
    # in-place residual (the decode step's o-/down-projection)
    h1, h2 = res.clone(), res.clone()
    skinny(a, w, splits=splits, residual=h1, out=h1)
    skinny(a, w, splits=splits, residual=h2, out=h2, counters=cnt)
    assert torch.equal(h1, h2) and int(cnt.abs().sum()) == 0


@pytest.mark.parametrize("M,N,K,splits", [(1, 6144, 4096, 2), (16, 28672, 4096, 1), (9, 6144, 4096, 4), (32, 4096, 4096, 4),
                                          (3, 7680, 5120, 2)])
def test_skinny_gemm_rmsnorm_of_the_staged_rows(M, N, K, splits):
    """a_norm_w: every CTA normalises the rows it stages (1/rms over the FULL row, also when it stages one k-split):
    equals rmsnorm (HF: w * (x * rstd).to(dtype)) followed by the plain projection."""
    L = _L()
    lib = L.load()
    torch.manual_seed(M + N + K)
    a, w = rnd(M, K), rnd(N, K, scale=0.05)
    nw = (1.0 + 0.1 * torch.randn(K, device="cuda")).to(torch.bfloat16)
    cnt = torch.zeros(N // 8, device="cuda", dtype=torch.int32)
    normed = torch.empty_like(a)
    L.check(lib.slime_op_rmsnorm(L.ptr(a), L.ptr(nw), L.ptr(normed), M, K, 1e-5, L.stream_ptr()), "rmsnorm")
    want = skinny(normed, w, splits=splits)
    got = skinny(a, w, splits=splits, a_norm_w=nw, counters=cnt if splits > 1 else None)
    e = rel_l2(got, want)
    assert e < 1e-3, f"staged RMSNorm vs rmsnorm kernel + projection: rel-L2 {e:.3e}"  # 1/rms may differ in the last bit
    af = a.float()
    ref = (nw.float() * (af * torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16).float()).to(torch.bfloat16).float() @ w.float().t()
    assert rel_l2(got, ref) < 4e-3
    assert torch.equal(got, skinny(a, w, splits=splits, a_norm_w=nw, counters=cnt if splits > 1 else None))


@pytest.mark.parametrize("heads,kv_heads", [(32, 8), (8, 8), (16, 2)])
@pytest.mark.parametrize("splits", [3, 8, 32])
def test_decode_attention_in_kernel_merge(heads, kv_heads, splits):
    """kv splits merged by the last CTA of a (sequence, kv head) instead of a merge launch: same result, tickets back at zero."""
    torch.manual_seed(heads + splits)
    B, cache_len, hd = 5, 700, 128
    kc, vc = rnd(B, cache_len, kv_heads * hd), rnd(B, cache_len, kv_heads * hd)
    qkv = rnd(B, (heads + 2 * kv_heads) * hd)
    lens = torch.tensor([0, 1, 15, 333, cache_len - 1], device="cuda", dtype=torch.int32)
    cnt = torch.zeros(B * kv_heads, device="cuda", dtype=torch.int32)
    plain = decode_attention(qkv, kc, vc, lens, heads, kv_heads, splits)
    for _ in range(2):
        fused = decode_attention(qkv, kc, vc, lens, heads, kv_heads, splits, counters=cnt)
        assert rel_l2(fused, plain) < 1e-6 and (fused.float() - plain.float()).abs().max() <= 2 ** -8 * plain.float().abs().max()
        assert int(cnt.abs().sum()) == 0


@pytest.mark.parametrize("pname,B", [("small", 1), ("small", 18), ("tiny", 32)])
def test_decode_step_fused_chain_matches_the_finishing_kernels(pname, B):
    """slime_set_decode_fused(7): 5 launches per layer (split reductions finished in the producing kernels, RMSNorm in the
    consumer's staging) against the 9-launch chain with finishing kernels; launch count checked."""
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_state_dict

    cfg = preset(pname)
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(synth_state_dict(cfg))
    lib = eng.lib
    torch.manual_seed(B)
    lens0 = [40 + 3 * (b % 11) for b in range(B)]
    rows = (torch.randn(sum(lens0), cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens0).cumsum(0)), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n) for n in lens0]).to(device="cuda", dtype=torch.int32)
    xs = [(torch.randn(B, cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(3)]
    outs, launches = [], []
    for fused in (7, 0, 7, 1, 2, 4):
        lib.slime_set_decode_fused(fused)
        try:
            eng.attach_kv_cache(B, 128)
            eng.decoder_prefill(rows, cu, pos, lens0)
            lens = torch.tensor(lens0, dtype=torch.int32, device="cuda")
            step = []
            n0 = lib.slime_launch_count()
            for x in xs:
                step.append(eng.decode_step(x, lens).clone())
                lens = lens + 1
            launches.append((lib.slime_launch_count() - n0) / len(xs))
            outs.append(torch.stack(step))
        finally:
            eng.detach_kv_cache()
            lib.slime_set_decode_fused(-1)
    e = rel_l2(outs[0], outs[1])
    print(f"decode step {pname} B={B}: fused vs finishing kernels rel-L2 {e:.3e}; launches per step {launches[0]:.0f} vs {launches[1]:.0f}")
    assert torch.isfinite(outs[0]).all()
    assert e < 3e-3
    assert torch.equal(outs[0], outs[2]), "fused chain must be repeatable bit for bit"
    assert launches[0] <= 5 * cfg.num_hidden_layers + 2 < launches[1]
    # the partial settings (A/B knobs): in-kernel split-K finish / in-kernel kv merge alone do not change a bit
    assert torch.equal(outs[3], outs[1]) and rel_l2(outs[4], outs[1]) < 1e-6 and rel_l2(outs[5], outs[0]) < 1e-6


@pytest.fixture(params=[2, 1], ids=["mma", "cuda_core"])
def attn_mode(request):
    """2 = split-KV kernel on mma.sync (default), 1 = split-KV kernel on CUDA cores."""
    lib = _L().load()
    lib.slime_decode_attention_set_mode(request.param)
    yield request.param
    lib.slime_decode_attention_set_mode(-1)


@pytest.mark.parametrize("heads,kv_heads", [(32, 8), (8, 8), (4, 2), (16, 2)])
@pytest.mark.parametrize("splits", [0, 1, 3, 8, 32])
def test_decode_attention_split_kv(heads, kv_heads, splits, attn_mode):
    if attn_mode == 1 and heads // kv_heads == 8 and splits > 0:
        pytest.skip("the CUDA-core split kernel covers G <= 4 (falls back to the one-CTA-per-head kernel)")
    torch.manual_seed(heads + splits)
    B, cache_len, hd = 5, 700, 128
    kc, vc = rnd(B, cache_len, kv_heads * hd), rnd(B, cache_len, kv_heads * hd)
    # q inside a wider packed row (the decode step passes the qkv buffer)
    qkv = rnd(B, (heads + 2 * kv_heads) * hd)
    lens = torch.tensor([0, 1, 15, 333, cache_len - 1], device="cuda", dtype=torch.int32)
    out = decode_attention(qkv, kc, vc, lens, heads, kv_heads, splits)
    G = heads // kv_heads
    ref = torch.zeros(B, heads * hd, device="cuda")
    for b in range(B):
        n = int(lens[b]) + 1
        qb = qkv[b, :heads * hd].float().reshape(heads, hd)
        k = kc[b, :n].float().reshape(n, kv_heads, hd).repeat_interleave(G, 1)
        v = vc[b, :n].float().reshape(n, kv_heads, hd).repeat_interleave(G, 1)
        s = torch.einsum("hd,nhd->hn", qb, k) / math.sqrt(hd)
        ref[b] = torch.einsum("hn,nhd->hd", s.softmax(-1), v).reshape(-1)
    e = rel_l2(out, ref)
    assert torch.isfinite(out.float()).all()
    assert e < 4e-3, f"decode attention heads={heads}/{kv_heads} splits={splits}: rel-L2 {e:.3e}"
    assert torch.equal(out, decode_attention(qkv, kc, vc, lens, heads, kv_heads, splits)), "must be deterministic"


@pytest.mark.parametrize("B", [1, 18])
def test_decode_step_new_kernels_match_the_tile_kernels(B):
    """One decode step end to end (slime_decoder_decode_fwd): weight-streaming GEMMs + split-KV attention against the
    tcgen05 GEMMs + one-CTA-per-head attention on the same cache."""
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_state_dict

    cfg = preset("small")
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(synth_state_dict(cfg))
    lib = eng.lib
    torch.manual_seed(B)
    lens0 = [40 + 3 * b for b in range(B)]
    rows = (torch.randn(sum(lens0), cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens0).cumsum(0)), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n) for n in lens0]).to(device="cuda", dtype=torch.int32)
    x = (torch.randn(B, cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16)
    outs = []
    for mode in (1, 0):
        lib.slime_gemm_set_skinny_mode(mode)
        lib.slime_decode_attention_set_mode(mode)
        try:
            eng.attach_kv_cache(B, 128)
            eng.decoder_prefill(rows, cu, pos, lens0)
            lens = torch.tensor(lens0, dtype=torch.int32, device="cuda")
            step = []
            for _ in range(3):
                step.append(eng.decode_step(x, lens).clone())
                lens = lens + 1
            outs.append(torch.stack(step))
        finally:
            eng.detach_kv_cache()
            lib.slime_gemm_set_skinny_mode(-1)
            lib.slime_decode_attention_set_mode(-1)
    e = rel_l2(outs[0], outs[1])
    print(f"decode step B={B}: new vs tile kernels rel-L2 {e:.3e}")
    assert torch.isfinite(outs[0]).all()
    assert e < 6e-3


@pytest.mark.parametrize("pname,B", [("small", 1), ("small", 7), ("tiny", 24)])
def test_decode_step_pdl_is_bit_identical(pname, B):
    """Programmatic dependent launch only moves WHEN a kernel's prologue runs: with every split reduction in a fixed
    order the logits of a chain of decode steps must be bit-identical with and without it (a race on the activations,
    the split-K scratch or the cache would show up here), over repeated runs."""
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_state_dict

    cfg = preset(pname)
    eng = SlimeEngine(cfg, 0)
    eng.load_state_dict(synth_state_dict(cfg))
    lib = eng.lib
    torch.manual_seed(B)
    lens0 = [33 + 5 * (b % 9) for b in range(B)]
    rows = (torch.randn(sum(lens0), cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens0).cumsum(0)), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n) for n in lens0]).to(device="cuda", dtype=torch.int32)
    xs = [(torch.randn(B, cfg.hidden_size, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(6)]
    outs = []
    for mode, pf in ((1, 15), (0, 0), (1, 3)):
        lib.slime_set_pdl_mode(mode)
        lib.slime_set_decode_prefetch(pf)  # L2 prefetch duties of the small kernels: no effect on results either
        try:
            eng.attach_kv_cache(B, 96)
            eng.decoder_prefill(rows, cu, pos, lens0)
            lens = torch.tensor(lens0, dtype=torch.int32, device="cuda")
            step = []
            for x in xs:
                step.append(eng.decode_step(x, lens).clone())
                lens = lens + 1
            outs.append(torch.stack(step))
        finally:
            eng.detach_kv_cache()
            lib.slime_set_pdl_mode(-1)
            lib.slime_set_decode_prefetch(-1)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]), "PDL on vs off"
    assert torch.equal(outs[0], outs[2]), "PDL run to run"
