#!/usr/bin/env python
"""Decode-step benchmark: ms per generated token and achieved HBM GB/s of
slime_decoder_decode_fwd on SliME-Llama3-8B dimensions, new kernels (weight-streaming GEMM + split-KV attention)
against the tile kernels they replace, plus the isolated projections of one decoder layer.

    python tools/bench_decode.py [--model llama3-8b] [--ctx 1380] [--batches 1,16] [--steps 32] [--out gpurun_out/decode_bench.json]

The step is HBM-bound: algorithmic bytes = every decoder weight once (bf16) + the KV cache rows of every sequence
once per kv head, per step; the figure is compared with MEASURED_PEAKS.json's copy bandwidth."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ctx", type=int, default=1380)
    ap.add_argument("--batches", default="1,16")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--no-projections", action="store_true", help="skip the isolated projection timings")
    ap.add_argument("--quick", action="store_true", help="only the two PDL configurations (mma / CUDA-core attention)")
    ap.add_argument("--variants", type=int, default=0, help="only the first N kernel configurations")
    ap.add_argument("--out", default=None, help="write the rows as JSON here (default: print only)")
    args = ap.parse_args()

    import torch

    from slime_b200 import _lib as L
    from slime_b200.config import preset
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_tensor, weight_specs

    over = {"num_hidden_layers": args.layers} if args.layers else {}
    cfg = preset(args.model, **over)
    dev = torch.device("cuda", 0)
    specs = {n: (s, k) for n, s, k in weight_specs(cfg)}
    eng = SlimeEngine(cfg, 0, max_pos=4096)
    eng.load_weights(lambda n: synth_tensor(n, specs[n][0], specs[n][1], 3407, device=dev, dtype=torch.bfloat16),
                     groups=("llm",))
    lib = eng.lib
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        hbm, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    H, I, V, hd = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.head_dim
    QD, KD = cfg.num_attention_heads * hd, cfg.num_key_value_heads * hd
    QKV = QD + 2 * KD
    w_bytes = cfg.num_hidden_layers * (QKV * H + H * QD + 2 * I * H + H * I) * 2 + V * H * 2

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    results = {"model": cfg.name, "ctx": args.ctx, "layers": cfg.num_hidden_layers, "hbm_peak_gbs": hbm,
               "peak_source": peak_src, "weight_bytes_per_step": w_bytes, "steps": []}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B in [int(b) for b in args.batches.split(",")]:
        lens0 = [args.ctx - (b % 7) for b in range(B)]
        rows = (torch.randn(sum(lens0), H, device=dev) * 0.5).to(torch.bfloat16)
        cu = torch.tensor([0] + list(torch.tensor(lens0).cumsum(0)), dtype=torch.int32, device=dev)
        pos = torch.cat([torch.arange(n) for n in lens0]).to(device=dev, dtype=torch.int32)
        x = (torch.randn(B, H, device=dev) * 0.5).to(torch.bfloat16)
        kv_bytes = cfg.num_hidden_layers * sum(lens0) * 2 * KD * 2
        for mode, amode, pdl, pf, fused, name in (
                (1, 2, 1, 0, 7, "fused chain (5 launches / layer): weight-streaming GEMM with in-kernel split-K finish + RMSNorm in the staging, split-KV attention (mma.sync) with in-kernel merge, PDL"),
                (1, 2, 1, 0, 0, "weight-streaming GEMM + split-KV attention (mma.sync) + finishing kernels, PDL"),
                (1, 2, 1, 0, 1, "finishing kernels, but the QKV split-K sum in kernel (mask 1), PDL"),
                (1, 2, 1, 0, 2, "finishing kernels, but the attention kv merge in kernel (mask 2), PDL"),
                (1, 2, 1, 0, 3, "masks 1 + 2, PDL"),
                (1, 2, 1, 0, 4, "RMSNorm in the staging, split sums by finishing kernels (mask 4), PDL"),
                (1, 2, 1, 0, 5, "masks 1 + 4, PDL"),
                (1, 2, 0, 0, 7, "fused chain, ordinary launches"),
                (1, 1, 1, 0, 0, "weight-streaming GEMM + split-KV attention (CUDA cores) + finishing kernels, PDL"),
                (1, 2, 0, 0, 0, "weight-streaming GEMM + split-KV attention (mma.sync) + finishing kernels, ordinary launches"),
                (0, 0, 0, 0, 0, "tcgen05 tile GEMM + one CTA per head"))[:(args.variants or (2 if args.quick else None))]:
            lib.slime_set_decode_fused(fused)
            lib.slime_gemm_set_skinny_mode(mode)
            lib.slime_decode_attention_set_mode(amode)
            lib.slime_set_pdl_mode(pdl)
            lib.slime_set_decode_prefetch(pf)
            eng.attach_kv_cache(B, args.ctx + 8)
            try:
                eng.decoder_prefill(rows, cu, pos, lens0)
                lens = torch.tensor(lens0, dtype=torch.int32, device=dev)
                for _ in range(3):
                    logits = eng.decode_step(x, lens)
                n0 = lib.slime_launch_count()
                ms = timed(lambda: eng.decode_step(x, lens), args.steps)
                launches = (lib.slime_launch_count() - n0) / args.steps
                # kernel classes through the library's own event profiler (separate pass: the events perturb)
                lib.slime_profile_enable(1)
                for _ in range(4):
                    eng.decode_step(x, lens)
                torch.cuda.synchronize()
                pms, pwork, pl = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
                lib.slime_profile_collect(pms, pwork, pl)
                lib.slime_profile_enable(0)
                row = {"batch": B, "kernels": name, "ms_per_step": ms, "tokens_per_s": B / ms * 1e3,
                       "launches_per_step": launches, "bytes_per_step": w_bytes + kv_bytes,
                       "achieved_gbs": (w_bytes + kv_bytes) / ms / 1e6, "frac_of_hbm_peak": (w_bytes + kv_bytes) / ms / 1e6 / hbm,
                       "profiled_ms_per_step": {"tcgen05_gemm": pms[0] / 4, "attention": pms[1] / 4, "hbm_kernels": pms[2] / 4},
                       "finite": bool(torch.isfinite(logits).all())}
                if mode == 1 and pms[2] > 0:  # (event pairs around each launch serialise the PDL overlap away)
                    row["weight_stream_gbs"] = pwork[2] / (pms[2] / 1e3) / 1e9
                results["steps"].append(row)
                print(json.dumps(row), flush=True)
            finally:
                eng.detach_kv_cache()
                lib.slime_gemm_set_skinny_mode(-1)
                lib.slime_decode_attention_set_mode(-1)
                lib.slime_set_pdl_mode(-1)
                lib.slime_set_decode_prefetch(-1)
                lib.slime_set_decode_fused(-1)

    # ---- isolated projections of one layer at M = 1 / 16 (L2 flushed between launches: weights come from HBM) ----
    gem = []
    shapes = [("qkv", QKV, H), ("o_proj", H, QD), ("gate_up", 2 * I, H), ("down", H, I), ("lm_head", V, H)]
    for M in (() if args.no_projections else (1, 16)):
        for nm, N, K in shapes:
            a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
            w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            ws = torch.empty(8 * M * N, device=dev, dtype=torch.float32)
            row = {"name": nm, "M": M, "N": N, "K": K, "weight_mb": N * K * 2 / 1e6}

            def run_skinny(splits):
                L.check(lib.slime_op_gemm_skinny(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, L.ptr(out), None, N,
                                                 splits, L.ptr(ws), ws.numel(), None, None, 0.0, None, None, 0, 0, 0,
                                                 L.stream_ptr()), "skinny")

            def run_tc():
                L.check(lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, 0, L.ptr(out), None, N,
                                          L.stream_ptr()), "gemm")

            def cold(fn):
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                return sorted(ts)[len(ts) // 2]

            for sp in (1, 2, 4, 8):
                if K // sp < 512 or (K + sp - 1) // sp > 4096:
                    continue
                t = cold(lambda: run_skinny(sp))
                row[f"skinny_s{sp}_us"] = t * 1e3
                row[f"skinny_s{sp}_gbs"] = N * K * 2 / t / 1e6
            lib.slime_gemm_set_skinny_mode(0)
            t = cold(run_tc)
            lib.slime_gemm_set_skinny_mode(-1)
            row["tcgen05_us"] = t * 1e3
            row["tcgen05_gbs"] = N * K * 2 / t / 1e6
            gem.append(row)
            print(json.dumps(row), flush=True)
    results["projections"] = gem
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
