#!/usr/bin/env python
"""bench.py - prefill tokens/s of the B200-native SliME path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # N = 1
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...                      # the reference's CPU path on host cores

A "step" = one prefill pass over one synthetic batch: SliME-Llama3-8B, 672x672 image (1 global + 2x2
local 336 px crops), 256-token prompt, `--batch` samples per GPU (weak scaling: per-GPU work fixed).
  value : prefill tokens/s with pixels/ids already resident in HBM (real spliced tokens, no padding)
  e2e   : the same metric through the public API (SlimeEngine.prefill) from pinned HOST buffers, with
          the host->device copy of pixels+ids and the device->host read of the last-token logits inside
          the timed region
  roofline     : the tcgen05 GEMM kernel (dominant: ~95 % of FLOPs), algorithmic FLOPs / CUDA-event time
                 of its launches inside the timed region, against the measured bf16 peak
  cpu_baseline : the CPU oracle port of the reference algorithm on this box's host cores (bounded sample)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "prefill_tokens_per_sec"
UNIT = "tokens/s"


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tf_burst=d.get("bf16_tflops", 1590.0), tf_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def algorithmic_flops(cfg, n_crops, lengths, n_local_tokens_in):
    """SURVEY.md 8(d) formulas (2 FLOPs per MAC; 23 ViT layers; last-token logits; causal attention halved)."""
    D, I, P, S = cfg.vit_hidden, cfg.vit_mlp, cfg.vit_patches, cfg.vit_tokens
    K = 3 * cfg.vit_patch ** 2
    vit = n_crops * (2 * P * K * D + cfg.vit_layers_used * (2 * S * D * 3 * D + 4 * S * S * D + 2 * S * D * D + 4 * S * D * I))
    H = cfg.hidden_size
    B = len(lengths)
    n_loc_crops = n_crops - B
    q = cfg.mm_resampler_dim
    rs_local = n_loc_crops * (4 * P * D * D + 4 * q * P * D + 2 * q * D * D)
    rs_glob = B * (4 * P * D * D + 4 * 576 * P * D + 2 * 576 * D * D)
    proj_tok = 2 * (D * H + H * H)
    proj = (B * 2 * 576 + n_local_tokens_in) * proj_tok
    nh, nkv, hd, Iq, V = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size, cfg.vocab_size
    per_tok = cfg.num_hidden_layers * (2 * H * (nh + 2 * nkv) * hd + 2 * nh * hd * H + 6 * H * Iq)
    llm = sum(L * per_tok + cfg.num_hidden_layers * 2 * L * L * nh * hd + 2 * H * V for L in lengths)
    return dict(vision=vit, adapter=rs_local + rs_glob + proj, llm=llm, total=vit + rs_local + rs_glob + proj + llm)


# --------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------------
def cpu_baseline_sample(cfg, prompt_len, n_crops, decoder_layers_sampled=2, seed=3407, repeats=1, dtype=None):
    """Bounded sample of the SAME workload on the CPU: one sample (5 crops, T-token prompt) through the full
    vision tower + SliME adapter + router + splice at real dimensions, and `decoder_layers_sampled` of the
    decoder layers + final norm + last-token lm_head; the decoder layer time is scaled to all layers (they are
    identical in shape).  fp32, all host threads.  Returns (tokens/s, description, threads)."""
    import torch

    from oracle import slime_oracle as O
    from slime_b200.synth import synth_inputs, synth_tensor, weight_specs

    # give the CPU path its best shot: probe a few thread counts on a GEMM of the path's shape and keep the fastest
    # (on many-core hosts the largest count is not always the best one for L ~ 1400-row matrices)
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16) if 1 <= c <= ncpu}, reverse=True)
    a = torch.randn(1400, cfg.hidden_size)
    w = torch.randn(cfg.intermediate_size, cfg.hidden_size)
    best_t, best_s = ncpu, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.matmul(a, w.t())
        t0 = time.perf_counter()
        for _ in range(3):
            torch.matmul(a, w.t())
        dt = time.perf_counter() - t0
        if dt < best_s:
            best_t, best_s = c, dt
    del a, w
    threads = best_t
    torch.set_num_threads(threads)
    small = cfg.replace(num_hidden_layers=decoder_layers_sampled)
    sd = {}
    for name, shape, kind in weight_specs(small):
        if name.startswith("model.vision_tower") and f"layers.{cfg.vit_layers - 1}." in name:
            continue  # the last ViT layer is never executed (select_layer = -2)
        if kind in ("linear", "embed", "gate") and len(shape) >= 2 and shape[0] * shape[1] > 1 << 22:
            # cheap N(0, 1/fan_in) fill for the big matrices (values do not matter for timing)
            t = torch.empty(shape, dtype=torch.float32).normal_(0, 1.0 / (shape[-1] ** 0.5))
            sd[name] = t
        else:
            sd[name] = synth_tensor(name, shape, kind, seed)
    px, ids, mask = synth_inputs(small, 1, n_crops, prompt_len, seed=seed)
    if dtype is not None:  # the reference's other CPU configuration (bf16 is ~3x faster on AMX hosts)
        sd = {k: v.to(dtype) for k, v in sd.items()}
        px = px.to(dtype)
    from slime_b200.synth import grid_for_crops

    grids = [grid_for_crops(n_crops - 1)]
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            enc = O.encode_images(sd, small, px, ids, mask, grids)
            emb, am, pid, lab, lens = O.splice(sd["model.embed_tokens.weight"], ids, mask, None, enc["feats"])
            t1 = time.perf_counter()
            # decoder slice: time all-position hidden states through the sampled layers, last-token logits only
            x = emb[0, :lens[0]]
            tl0 = time.perf_counter()
            hidden = _decoder_layers_only(O, sd, small, x)
            tl1 = time.perf_counter()
            last = O.rms_norm(hidden[-1:], sd["model.norm.weight"], small.rms_norm_eps) @ sd["lm_head.weight"].t()
            tl2 = time.perf_counter()
            t_front = t1 - t0
            t_layers = (tl1 - tl0) * (cfg.num_hidden_layers / decoder_layers_sampled)
            t_head = tl2 - tl1
            total = t_front + t_layers + t_head
            if best is None or total < best[0]:
                best = (total, t_front, t_layers, t_head, lens[0])
    total, t_front, t_layers, t_head, L = best
    cpu_baseline_sample.last_seconds = total  # (scaled) seconds of the one-sample step; the reference arm's ms_per_step
    desc = (f"1 sample ({n_crops} crops, T={prompt_len}, L={L}) on {threads} threads, "
            f"{'fp32' if dtype is None else str(dtype).replace('torch.', '')} torch CPU: vision+adapter+router+"
            f"splice {t_front:.2f}s measured in full; {decoder_layers_sampled}/{cfg.num_hidden_layers} decoder layers measured "
            f"and scaled x{cfg.num_hidden_layers // decoder_layers_sampled} = {t_layers:.2f}s; final norm + last-token "
            f"lm_head {t_head:.2f}s")
    return L / total, desc, threads


def _decoder_layers_only(O, sd, cfg, x):
    """The layer loop of oracle.llama_prefill without the all-position lm_head (same arithmetic)."""
    import math

    import torch
    import torch.nn.functional as F

    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    L = x.shape[0]
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    ang = torch.arange(L, dtype=torch.float32)[:, None] * inv[None]
    cos, sin = torch.cat([ang, ang], -1).cos()[None].to(x.dtype), torch.cat([ang, ang], -1).sin()[None].to(x.dtype)
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)  # noqa: E731
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        h = O.rms_norm(x, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = (h @ sd[p + "self_attn.q_proj.weight"].t()).view(L, nh, hd).transpose(0, 1)
        k = (h @ sd[p + "self_attn.k_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        v = (h @ sd[p + "self_attn.v_proj.weight"].t()).view(L, nkv, hd).transpose(0, 1)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        k, v = k.repeat_interleave(nh // nkv, 0), v.repeat_interleave(nh // nkv, 0)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        a = torch.softmax(s.masked_fill(~causal, float("-inf")), -1) @ v
        x = x + a.transpose(0, 1).reshape(L, nh * hd) @ sd[p + "self_attn.o_proj.weight"].t()
        h = O.rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = F.silu(h @ sd[p + "mlp.gate_proj.weight"].t()) * (h @ sd[p + "mlp.up_proj.weight"].t())
        x = x + g @ sd[p + "mlp.down_proj.weight"].t()
    return x


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------
def run_reference_arm(args, cfg):
    """The reference's CPU implementation of the path on this box's host cores.  The reference is a Python
    repo that cannot travel to the GPU box (/root/reference is absent there), so this times the CPU oracle
    port of its algorithm (oracle/slime_oracle.py, pinned to the reference's golden vectors)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    vals, secs, desc, threads = [], [], "", 1
    for _ in range(max(1, min(args.steps, 2))):
        v, desc, threads = cpu_baseline_sample(cfg, args.prompt_len, args.crops, decoder_layers_sampled=2)
        vals.append(v)
        secs.append(getattr(cpu_baseline_sample, "last_seconds", None))
    value = max(vals)
    step_ms = secs[vals.index(value)] * 1e3 if secs[vals.index(value)] is not None else None
    bf16_value, bf16_desc = None, None
    try:  # the reference's faster CPU configuration on AMX hosts; the headline stays the fp32 figure BASELINE.json names
        bf16_value, bf16_desc, _ = cpu_baseline_sample(cfg, args.prompt_len, args.crops, decoder_layers_sampled=2,
                                                       dtype=torch.bfloat16)
    except Exception as e:  # noqa: BLE001
        bf16_desc = f"failed: {e!r}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "value_bf16": bf16_value, "sample_bf16": bf16_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cfg):
    return {"workload": f"{cfg.name} prefill, 672x672 image = {args.crops} crops of 336 px (1 global + 2x2 local, spatial "
                        f"merge), {args.prompt_len}-token prompt, batch {args.batch} per GPU, top-p {cfg.mm_resampler_topp}",
            "model": cfg.name, "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus,
            "crops_per_image": args.crops, "prompt_len": args.prompt_len, "top_p": cfg.mm_resampler_topp,
            "parallelism": f"dp{args.gpus}", "l2": "inputs (16 GB of weights streamed every step) exceed the 126 MB L2"}


# --------------------------------------------------------------------------------------------
# main
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU per step")
    ap.add_argument("--crops", type=int, default=5)
    ap.add_argument("--prompt-len", type=int, default=256)
    ap.add_argument("--topp", type=float, default=None, help="override mm_resampler_topp (default 0.95)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=None, help="debug only: truncate the decoder (INVALID as a bench number)")
    args = ap.parse_args()

    from slime_b200.config import preset

    over = {}
    if args.topp is not None:
        over["mm_resampler_topp"] = args.topp
    if args.layers is not None:
        over["num_hidden_layers"] = args.layers
    cfg = preset(args.model, **over)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch
    import torch.distributed as dist

    from slime_b200 import _lib as L
    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import grid_for_crops, synth_inputs, synth_tensor, weight_specs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("SLIME_NCCL_DEBUG", "WARN")  # stdout = the one JSON line (no banner)
        dist.init_process_group("nccl", device_id=dev)

    # ---- model: random-init weights of the named architecture, generated on the GPU ----
    specs = {name: (shape, kind) for name, shape, kind in weight_specs(cfg)}
    eng = SlimeEngine(cfg, local_rank, max_pos=4096)
    eng.load_weights(lambda name: synth_tensor(name, specs[name][0], specs[name][1], 3407, device=dev,
                                               dtype=torch.bfloat16))
    lib = eng.lib

    # ---- inputs: a different synthetic batch shard per rank, pinned on the host ----
    B = args.batch
    px_h, ids_h, mask_h = synth_inputs(cfg, B, args.crops, args.prompt_len, seed=3407 + rank)
    px_h = px_h.to(torch.bfloat16).pin_memory()
    ids_h = ids_h.pin_memory()
    mask_h = mask_h.pin_memory()
    grids = [grid_for_crops(args.crops - 1)] * B
    px_d, ids_d, mask_d = px_h.to(dev), ids_h.to(dev), mask_h.to(dev)
    gathered = torch.empty(world * B, cfg.vocab_size, dtype=torch.float32, device=dev) if world > 1 else None
    logits_h = torch.empty(B, cfg.vocab_size, dtype=torch.float32).pin_memory()

    def step_device():
        res = eng.prefill(px_d, ids_d, mask_d, grids=grids)
        if world > 1:  # the one exchange step of the path: gather the last-token logits
            dist.all_gather_into_tensor(gathered, res.logits_last)
        return res

    def step_e2e():
        px = px_h.to(dev, non_blocking=True)
        ids = ids_h.to(dev, non_blocking=True)
        mask = mask_h.to(dev, non_blocking=True)
        res = eng.prefill(px, ids, mask, grids=grids)
        if world > 1:
            dist.all_gather_into_tensor(gathered, res.logits_last)
        logits_h.copy_(res.logits_last, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.slime_launch_count()
        e0.record()
        toks = 0
        for _ in range(steps):
            toks += fn().total_tokens
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.slime_launch_count() - launches0
        t = torch.tensor([ms, float(toks)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, toks = float(tmax[0]), float(tsum[1])
        return ms, toks, launches

    for _ in range(max(args.warmup, 3)):
        res = step_device()
    torch.cuda.synchronize()
    lengths = res.lengths

    # ---- timed region 1: device-resident inputs (value) with the library's event profiler on ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.slime_profile_enable(1)
    ms, toks, launches = timed(step_device, args.steps)
    import ctypes as C

    pms, pwork, pl = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
    lib.slime_profile_collect(pms, pwork, pl)
    lib.slime_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = toks / (ms / 1e3)

    # ---- timed region 2: end to end from pinned host memory through the public API ----
    for _ in range(2):
        step_e2e()
    ms_e2e, toks_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = toks_e2e / (ms_e2e / 1e3)
    h2d = px_h.numel() * px_h.element_size() + ids_h.numel() * 8 + mask_h.numel() * 8
    d2h = logits_h.numel() * 4 + (B + 1) * 4

    # ---- timed region 3: from raw RGB bytes - GPU pre-processing (process_images) + prefill ----
    raw = None
    if args.crops == 5:
        import numpy as np

        from slime_b200.preprocess import preprocess_images

        rng = np.random.default_rng(3407 + rank)
        imgs = [rng.integers(0, 256, (672, 672, 3), dtype=np.uint8) for _ in range(B)]  # 672^2 -> 1 + 2x2 crops

        def step_raw():
            crops, _ = preprocess_images(imgs, "anyres", dtype=torch.bfloat16, device=dev)
            ids = ids_h.to(dev, non_blocking=True)
            mask = mask_h.to(dev, non_blocking=True)
            res = eng.prefill(torch.stack(crops), ids, mask, grids=grids)
            if world > 1:
                dist.all_gather_into_tensor(gathered, res.logits_last)
            logits_h.copy_(res.logits_last, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return res

        for _ in range(2):
            step_raw()
        ms_raw, toks_raw, _ = timed(step_raw, args.steps)
        raw = {"value": toks_raw / (ms_raw / 1e3), "unit": UNIT, "ms_per_step": ms_raw / args.steps,
               "h2d_bytes_per_step": B * 672 * 672 * 3 + ids_h.numel() * 8 + mask_h.numel() * 8,
               "note": "raw 672x672 RGB bytes on the host -> slime_preprocess_fwd (Pillow-exact resize, tiling, "
                       "CLIP normalise) -> prefill -> logits on the host"}

    # ---- ViT crops/s (secondary metric of BASELINE.json) ----
    flat = px_d.flatten(0, 1)
    for _ in range(2):
        eng.vision_tower(flat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.vision_tower(flat)
    e1.record()
    torch.cuda.synchronize()
    crops_per_s = 5 * flat.shape[0] / (e0.elapsed_time(e1) / 1e3) * world

    # ---- decode step after this prefill (secondary, HBM-bound): ms per generated token ----
    decode = None
    try:
        n_dec = 16
        eng.attach_kv_cache(B, max(lengths) + n_dec + 8)
        res_d = eng.prefill(px_d, ids_d, mask_d, grids=grids)  # fills the cache
        lens_d = torch.tensor(res_d.lengths, dtype=torch.int32, device=dev)
        table = eng.weights["llm.embed"]
        x_d = table[res_d.logits_last.argmax(-1)]
        for _ in range(3):
            eng.decode_step(x_d, lens_d)  # (same slot rewritten: timing only)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.slime_launch_count()
        e0.record()
        for _ in range(n_dec):
            eng.decode_step(x_d, lens_d)
        e1.record()
        torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / n_dec
        w_bytes = 2 * (cfg.num_hidden_layers * (cfg.qkv_dim * cfg.hidden_size + cfg.hidden_size * cfg.num_attention_heads * cfg.head_dim
                                                 + 3 * cfg.intermediate_size * cfg.hidden_size) + cfg.vocab_size * cfg.hidden_size)
        kv_bytes = cfg.num_hidden_layers * sum(res_d.lengths) * 2 * cfg.num_key_value_heads * cfg.head_dim * 2
        decode = {"batch": B, "ms_per_step": dms, "tokens_per_s": B / dms * 1e3 * world, "context": float(statistics.mean(lengths)),
                  "launches_per_step": (lib.slime_launch_count() - n0) / n_dec,
                  "hbm_bytes_per_step": w_bytes + kv_bytes, "achieved_gbs": (w_bytes + kv_bytes) / dms / 1e6}
    except Exception as e:  # noqa: BLE001 - never lose the prefill line over the secondary figure
        decode = {"error": repr(e)}
    finally:
        try:
            eng.detach_kv_cache()
        except Exception:  # noqa: BLE001
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    if decode is not None and "achieved_gbs" in decode:
        decode["frac_of_hbm_peak"] = decode["achieved_gbs"] / peaks["hbm"]
        decode["kernels"] = "weight-streaming mma.sync GEMM (csrc/gemm_skinny.cu) + split-KV mma.sync attention (csrc/decode_attn.cu), PDL"
    fl = algorithmic_flops(cfg, args.crops * B, lengths, (args.crops - 1) * B * cfg.mm_resampler_dim)
    gemm_tf = pwork[0] / (pms[0] / 1e3) / 1e12 if pms[0] > 0 else 0.0
    step_ms = ms / args.steps
    traffic, traffic_note = None, None
    try:  # DRAM bytes per launch of the dominant GEMM from the committed ncu --set full capture (tools/ncu_summary.py)
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f)["gemm"]
        traffic, traffic_note = tr["dram_bytes_per_launch"], f"mean over {tr['launches']} captured launches, {tr['source']}"
    except Exception:
        pass
    roofline = {
        "kernel": "gemm_bf16_tn_2cta_kernel / gemm_bf16_tn_kernel (tcgen05.mma cta_group::2 256x256x16 resp. 128xBNx16, TMA; "
                  "slime_b200/csrc/gemm2_sm100.cu, gemm_sm100.cu) - every dense contraction of the path",
        "bound": "tensor", "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tf / peaks["tf_sustained"], "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
        "launches_per_step": pl[0] / args.steps, "gemm_ms_per_step": pms[0] / args.steps,
        "gemm_share_of_step": (pms[0] / args.steps) / step_ms,
        "attention_ms_per_step": pms[1] / args.steps, "attention_launches_per_step": pl[1] / args.steps,
        "whole_step_tflops": fl["total"] / (step_ms / 1e3) / 1e12,
        "whole_step_frac_of_peak": fl["total"] / (step_ms / 1e3) / 1e12 / peaks["tf_sustained"],
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, cfg),
        "tokens_per_step": toks / args.steps, "mean_kept_local_tokens": float(statistics.mean(lengths)) - (args.prompt_len - 1) - 577,
        "vit_crops_per_sec": crops_per_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "e2e_from_rgb_bytes": raw,
        "decode_step": decode,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "algorithmic_tflop_per_step": {k: v / 1e12 for k, v in fl.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, desc, threads = cpu_baseline_sample(cfg, args.prompt_len, args.crops)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
            try:  # the same sample in bf16 (the faster CPU configuration where the host has AMX / AVX512-BF16)
                vb, descb, _ = cpu_baseline_sample(cfg, args.prompt_len, args.crops, dtype=torch.bfloat16)
                line["cpu_baseline"]["value_bf16"] = vb
                line["cpu_baseline"]["sample_bf16"] = descb
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"]["value_bf16"] = None
                line["cpu_baseline"]["sample_bf16"] = f"failed: {e!r}"
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
