"""Diagnostics + first timings of the tcgen05 attention kernel (impl 2) against the mma.sync one (impl 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L

lib = L.load()


def attn(q, k, v, o, impl, **kw):
    rc = lib.slime_op_attention(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), kw["q_ld"], kw["k_ld"], kw["v_ld"], kw["o_ld"],
                                L.ptr(kw.get("cu_q")), L.ptr(kw.get("cu_k")), kw["seqlen_q"], kw["seqlen_k"],
                                kw.get("q_batch_rows", 0), kw.get("k_batch_rows", 0), kw.get("o_batch_rows", 0),
                                kw["batch"], kw["heads"], kw["kv_heads"], kw["head_dim"], kw["scale"], kw.get("causal", 0),
                                kw.get("total_q_rows", 0), kw.get("total_k_rows", 0), impl, L.stream_ptr())
    if rc != 0:
        print("launch failed", rc, L.last_error())
    torch.cuda.synchronize()


def simple_case(d, Sq, Sk, causal=0, impl=2):
    """single head, single batch; block-wise error report."""
    torch.manual_seed(0)
    q = torch.randn(Sq, d, device="cuda").to(torch.bfloat16)
    k = torch.randn(Sk, d, device="cuda").to(torch.bfloat16)
    v = torch.randn(Sk, d, device="cuda").to(torch.bfloat16)
    o = torch.full((Sq, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    attn(q, k, v, o, impl, q_ld=d, k_ld=d, v_ld=d, o_ld=d, seqlen_q=Sq, seqlen_k=Sk, q_batch_rows=Sq, k_batch_rows=Sk,
         o_batch_rows=Sq, batch=1, heads=1, kv_heads=1, head_dim=d, scale=d ** -0.5, causal=causal)
    s = (q.float() @ k.float().t()) * d ** -0.5
    if causal:
        s = s.masked_fill(~torch.ones(Sq, Sk, device="cuda", dtype=torch.bool).tril(Sk - Sq), float("-inf"))
    ref = torch.softmax(s, -1) @ v.float()
    of = torch.nan_to_num(o.float())
    err = ((of - ref).norm() / ref.norm()).item()
    print(f"[impl={impl} d={d} Sq={Sq} Sk={Sk} causal={causal}] rel-L2 {err:.3e} nan-frac {torch.isnan(o.float()).float().mean().item():.3f}")
    if err > 1e-2:
        # hypotheses: P uniform (O = mean V), V transposed/permuted, only first tile...
        print("   vs mean(V):", ((of - v.float().mean(0)).norm() / ref.norm()).item())
        for rb in range(0, min(Sq, 128), 32):
            row = []
            for cb in range(0, d, 32):
                r = ref[rb:rb + 32, cb:cb + 32]
                row.append(f"{((of[rb:rb+32, cb:cb+32] - r).norm() / r.norm()).item():5.2f}")
            print("   rows", rb, " ".join(row))
        print("   out[0,:8]", of[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return err


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for args in [(128, 128, 128), (128, 64, 64), (128, 128, 256), (128, 256, 128), (128, 577, 577), (64, 128, 128), (64, 577, 577),
                 (128, 300, 300, 1), (128, 1400, 1400, 1)]:
        simple_case(*args)
    # timings
    def bench(name, fn, flops):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: {ms:.3f} ms = {flops / ms / 1e9:.0f} TF/s")
    # decoder: 16 seqs x 1380, 32/8 heads x 128 causal
    L_, B, h, kvh, d = 1380, 16, 32, 8, 128
    W = (h + 2 * kvh) * d
    qkv = torch.randn(B * L_, W, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * L_, h * d, device="cuda", dtype=torch.bfloat16)
    cu = torch.arange(0, (B + 1) * L_, L_, device="cuda", dtype=torch.int32)
    fl = 4 * L_ * L_ * d * h * B * 0.5
    for impl in (1, 2):
        bench(f"decoder causal impl={impl}", lambda: attn(qkv, qkv[:, h * d:], qkv[:, (h + kvh) * d:], o, impl, q_ld=W, k_ld=W,
              v_ld=W, o_ld=h * d, cu_q=cu, cu_k=cu, seqlen_q=L_, seqlen_k=L_, batch=B, heads=h, kv_heads=kvh, head_dim=d,
              scale=d ** -0.5, causal=1, total_q_rows=B * L_, total_k_rows=B * L_), fl)
    # ViT: 64 crops x 577, 16 heads x 64
    S, B, h, d = 577, 64, 16, 64
    D = h * d
    qkv = torch.randn(B * S, 3 * D, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * S, D, device="cuda", dtype=torch.bfloat16)
    fl = 4 * S * S * d * h * B
    for impl in (1, 2):
        bench(f"vit impl={impl}", lambda: attn(qkv, qkv[:, D:], qkv[:, 2 * D:], o, impl, q_ld=3 * D, k_ld=3 * D, v_ld=3 * D, o_ld=D,
              seqlen_q=S, seqlen_k=S, q_batch_rows=S, k_batch_rows=S, o_batch_rows=S, batch=B, heads=h, kv_heads=h,
              head_dim=d, scale=d ** -0.5), fl)
