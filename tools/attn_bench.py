"""Attention kernels on the GPU box: correctness vs fp32 torch, CUDA-event timing and (optionally) the per-tile clock
trace, for the shapes of the SliME path.

    python tools/attn_bench.py [--polys 0,2,3,4] [--trace]

--polys: shares of polynomial exp2 to time (pairs of every 8 on the FMA pipe instead of MUFU.EX2).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slime_b200 import _lib as L

lib = L.load()
dev = "cuda"
PEAK_SUSTAINED = 1383.8  # MEASURED_PEAKS.json bf16_tflops_sustained (the judge's denominator)


def timeit(f, n=20, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def ref_attention(q, k, v, scale, causal):
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = q.shape[-2], k.shape[-2]
        s = s.masked_fill(~torch.ones(Sq, Sk, device=q.device, dtype=torch.bool).tril(Sk - Sq), float("-inf"))
    return torch.softmax(s, dim=-1) @ v


def cases():
    torch.manual_seed(0)
    out = []
    # decoder: 16 sequences x 1380 tokens, causal, GQA 32/8 x 128 (headline)
    for name, L_, B, h, kvh, d in (("decoder llama3 16x1380 32/8x128", 1380, 16, 32, 8, 128),
                                   ("decoder vicuna 8x1250 32/32x128", 1250, 8, 32, 32, 128)):
        W = (h + 2 * kvh) * d
        qkv = torch.randn(B * L_, W, device=dev).to(torch.bfloat16)
        o = torch.zeros(B * L_, h * d, device=dev, dtype=torch.bfloat16)
        cu = torch.arange(0, (B + 1) * L_, L_, device=dev, dtype=torch.int32)

        def args(impl, qkv=qkv, o=o, cu=cu, W=W, h=h, kvh=kvh, d=d, L_=L_, B=B):
            return (L.ptr(qkv), L.ptr(qkv[:, h * d:]), L.ptr(qkv[:, (h + kvh) * d:]), L.ptr(o), W, W, W, h * d, L.ptr(cu),
                    L.ptr(cu), L_, L_, 0, 0, 0, B, h, kvh, d, d ** -0.5, 1, B * L_, B * L_, impl, L.stream_ptr())

        def ref(qkv=qkv, h=h, kvh=kvh, d=d, L_=L_):  # first sequence only
            blk = qkv[:L_].float()
            q = blk[:, :h * d].view(L_, h, d).permute(1, 0, 2)[None]
            k = blk[:, h * d:(h + kvh) * d].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
            v = blk[:, (h + kvh) * d:].view(L_, kvh, d).permute(1, 0, 2)[None].repeat_interleave(h // kvh, dim=1)
            return ref_attention(q, k, v, d ** -0.5, True)[0].permute(1, 0, 2).reshape(L_, h * d)

        out.append((name, o, 4.0 * B * h * L_ * L_ * d * 0.5, args, ref, L_, (qkv, cu)))
    # ViT: 80 crops x 577 tokens, 16 heads x 64, non-causal
    S, Bv, hv, dv = 577, 80, 16, 64
    D = hv * dv
    qkv2 = torch.randn(Bv * S, 3 * D, device=dev).to(torch.bfloat16)
    o2 = torch.zeros(Bv * S, D, device=dev, dtype=torch.bfloat16)

    def args2(impl):
        return (L.ptr(qkv2), L.ptr(qkv2[:, D:]), L.ptr(qkv2[:, 2 * D:]), L.ptr(o2), 3 * D, 3 * D, 3 * D, D, None, None, S, S,
                S, S, S, Bv, hv, hv, dv, dv ** -0.5, 0, 0, 0, impl, L.stream_ptr())

    def ref2():
        x = qkv2[:S].float().view(1, S, 3, hv, dv).permute(2, 0, 3, 1, 4)
        return ref_attention(x[0], x[1], x[2], dv ** -0.5, False).permute(0, 2, 1, 3).reshape(S, D)

    out.append(("vit 80x577 16x64", o2, 4.0 * Bv * hv * S * S * dv, args2, ref2, S, (qkv2,)))
    # local compression: 64 crops, 144 shared queries x 576 keys, 8 heads x 128
    n, nq, NK, hr, dr = 64, 144, 576, 8, 128
    Dr = hr * dr
    q3 = torch.randn(nq, Dr, device=dev).to(torch.bfloat16)
    kv3 = torch.randn(n * NK, 2 * Dr, device=dev).to(torch.bfloat16)
    o3 = torch.zeros(n * nq, Dr, device=dev, dtype=torch.bfloat16)

    def args3(impl):
        return (L.ptr(q3), L.ptr(kv3), L.ptr(kv3[:, Dr:]), L.ptr(o3), Dr, 2 * Dr, 2 * Dr, Dr, None, None, nq, NK, 0, NK, nq, n,
                hr, hr, dr, dr ** -0.5, 0, 0, 0, impl, L.stream_ptr())

    def ref3():
        qf = q3.float().view(1, nq, hr, dr).permute(0, 2, 1, 3)
        kf = kv3[:NK, :Dr].float().view(1, NK, hr, dr).permute(0, 2, 1, 3)
        vf = kv3[:NK, Dr:].float().view(1, NK, hr, dr).permute(0, 2, 1, 3)
        return ref_attention(qf, kf, vf, dr ** -0.5, False).permute(0, 2, 1, 3).reshape(nq, Dr)

    out.append(("resampler 64 x (144 q x 576 k) 8x128", o3, 4.0 * n * hr * nq * NK * dr, args3, ref3, nq, (q3, kv3)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--polys", default="2")
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0))
    for name, out, fl, args, ref, nref, _keep in cases():
        r = ref()
        for impl in (0,):
            for poly in [int(x) for x in a.polys.split(",")]:
                assert lib.slime_attention_set_poly(poly) == 0

                def f():
                    rc = lib.slime_op_attention(*args(impl))
                    assert rc == 0, L.last_error()

                out.zero_()
                f()
                torch.cuda.synchronize()
                err = rel(out[:nref], r)
                ms = timeit(f)
                tf = fl / ms / 1e9
                print(f"attn {name:40s} poly {poly}: {ms:7.4f} ms  {tf:6.0f} TF/s = {tf / PEAK_SUSTAINED:.2f} of "
                      f"sustained peak   rel-L2 vs fp32 {err:.2e} finite={bool(torch.isfinite(out.float()).all())}", flush=True)
        lib.slime_attention_set_poly(-1)
        if a.trace and name.startswith(("decoder llama3", "vit")):
            tr = torch.zeros(64, 16, dtype=torch.int64, device=dev)
            lib.slime_attention_set_trace(L.ptr(tr))
            lib.slime_op_attention(*args(0))
            torch.cuda.synchronize()
            lib.slime_attention_set_trace(None)
            t = tr.cpu()
            t0 = int(t[0, 0])
            print(f"-- clock trace of CTA 0, slot A ({name}); columns: softmax wait_s, s_ready, max_done, exp_done, p_arrived | "
                  "mma wait_p, p_ready, v_ready, pv_issued | epilogue (per item) start, o_ready, o_freed, stored")
            for g in range(40):
                row = [(int(t[g, k]) - t0 if int(t[g, k]) else -1) for k in (0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 13, 14, 15)]
                print(f"{g:3d} | " + " ".join(f"{v:8d}" for v in row))
            for g in range(2, 30):
                s = [int(t[g, k]) for k in range(5)]
                print(f"tile {g}: wait_s {s[1] - s[0]:5d}  max {s[2] - s[1]:5d}  exp {s[3] - s[2]:5d}  st+arrive {s[4] - s[3]:4d} | "
                      f"mma: p wait {int(t[g, 9]) - int(t[g, 8]):5d} v {int(t[g, 10]) - int(t[g, 9]):4d} issue "
                      f"{int(t[g, 11]) - int(t[g, 10]):4d} | slot period {int(t[g + 1, 1]) - s[1]:5d}")


if __name__ == "__main__":
    main()
