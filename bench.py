#!/usr/bin/env python
"""bench.py - prefill tokens/s of the B200-native SliME path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # N = 1
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...                      # the reference's CPU path on host cores

A "step" = one prefill pass over one synthetic batch: SliME-Llama3-8B, 672x672 image (1 global + 2x2
local 336 px crops), 256-token prompt, `--batch` samples per GPU (weak scaling: per-GPU work fixed).
  value : prefill tokens/s with pixels/ids already resident in HBM (real spliced tokens, no padding)
  e2e   : the same metric through the public API (SlimeEngine.prefill) from pinned HOST buffers, with
          the host->device copy of pixels+ids and the device->host read of the (gathered) last-token
          logits inside the timed region
  roofline     : the tcgen05 GEMM kernel (dominant: ~95 % of FLOPs), algorithmic FLOPs / CUDA-event time
                 of its launches inside the timed region, against the measured bf16 peak
  cpu_baseline : the CPU oracle port of the reference algorithm on this box's host cores (bounded sample)
  secondary    : short runs of the other modes / BASELINE.json configurations (batch-1 latency, top-p 1.0
                 fixed-length mode, the fp16 build, and at 8 GPUs the per-GPU shapes of configs 4 and 5)
Multi-GPU: the only exchange step is the all-gather of the last-token logits; it runs on a side stream
(slime_b200/parallel.py LogitsGather) and `gather_check` verifies, un-timed, that the gathered blocks are
bit-identical to what a single GPU computes for the same shard.
Prints ONE JSON line on rank 0 (stdout); NCCL's own log, if NCCL_DEBUG is set, goes to stderr.
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
class CpuBaseline:
    """ONE sample of the benchmark workload (n crops, T-token prompt) through oracle/slime_oracle.py on the host
    cores, at the real dimensions and with EVERY stage executed in full: vision tower + SliME adapter + router +
    splice, all `num_hidden_layers` decoder layers (oracle.llama_hidden), final norm + last-token lm_head.  To keep
    host memory bounded the decoder holds `layers_held` layers' weights and cycles through them (layer l runs with the
    weights of layer l % layers_held; 0.9 GB per fp32 layer is far beyond any cache, so this does not flatter the CPU).
    Nothing is extrapolated."""

    def __init__(self, cfg, prompt_len, n_crops, dtype=None, layers_held=2, seed=3407, threads=None):
        import torch

        from slime_b200.synth import grid_for_crops, synth_inputs, synth_tensor, weight_specs

        self.torch, self.cfg, self.dtype = torch, cfg, dtype
        self.threads = threads or self.pick_threads(cfg)
        torch.set_num_threads(self.threads)
        held = cfg.replace(num_hidden_layers=layers_held)
        self.layer_map = [l % layers_held for l in range(cfg.num_hidden_layers)]
        sd = {}
        for name, shape, kind in weight_specs(held):
            if name.startswith("model.vision_tower") and f"layers.{cfg.vit_layers - 1}." in name:
                continue  # the last ViT layer is never executed (select_layer = -2)
            if kind in ("linear", "embed", "gate") and len(shape) >= 2 and shape[0] * shape[1] > 1 << 22:
                # cheap N(0, 1/fan_in) fill for the big matrices (values do not matter for timing)
                sd[name] = torch.empty(shape, dtype=torch.float32).normal_(0, 1.0 / (shape[-1] ** 0.5))
            else:
                sd[name] = synth_tensor(name, shape, kind, seed)
        px, self.ids, self.mask = synth_inputs(held, 1, n_crops, prompt_len, seed=seed)
        if dtype is not None:  # the reference's other CPU configuration (bf16 is ~3x faster on AMX hosts)
            sd = {k: v.to(dtype) for k, v in sd.items()}
            px = px.to(dtype)
        self.sd, self.px = sd, px
        self.grids = [grid_for_crops(n_crops - 1)]
        self.n_crops, self.prompt_len = n_crops, prompt_len

    @staticmethod
    def pick_threads(cfg):
        """Give the CPU path its best shot: probe a few thread counts on a GEMM of the path's shape, keep the fastest."""
        import torch

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
        return best_t

    def step(self):
        """One full pass; returns (seconds, spliced tokens, (front, decoder, head) seconds)."""
        from oracle import slime_oracle as O

        torch, sd, cfg = self.torch, self.sd, self.cfg
        with torch.no_grad():
            t0 = time.perf_counter()
            enc = O.encode_images(sd, cfg, self.px, self.ids, self.mask, self.grids)
            emb, _, _, _, lens = O.splice(sd["model.embed_tokens.weight"], self.ids, self.mask, None, enc["feats"])
            t1 = time.perf_counter()
            hidden = O.llama_hidden(sd, cfg, emb[0, :lens[0]], layers=self.layer_map)
            t2 = time.perf_counter()
            last = O.rms_norm(hidden[-1:], sd["model.norm.weight"], cfg.rms_norm_eps) @ sd["lm_head.weight"].t()
            t3 = time.perf_counter()
        assert last.shape[-1] == cfg.vocab_size
        return t3 - t0, lens[0], (t1 - t0, t2 - t1, t3 - t2)

    def describe(self, secs, L, parts):
        dt = "fp32" if self.dtype is None else str(self.dtype).replace("torch.", "")
        return (f"1 sample ({self.n_crops} crops, T={self.prompt_len}, L={L}) on {self.threads} threads, {dt} torch CPU, "
                f"every stage in full: vision+adapter+router+splice {parts[0]:.2f}s, all {self.cfg.num_hidden_layers} "
                f"decoder layers {parts[1]:.2f}s (weights of {len(set(self.layer_map))} layers cycled), final norm + "
                f"last-token lm_head {parts[2]:.2f}s")


def cpu_baseline_block(cfg, prompt_len, n_crops):
    """cpu_baseline object of the main line: one bf16 pass (the CPU's faster configuration, AMX / AVX512-BF16) and one
    fp32 pass (the dtype BASELINE.json's config 1 names), ~15-20 s of CPU work."""
    import torch

    out = {"unit": UNIT, "kind": "port"}
    try:
        cb = CpuBaseline(cfg, prompt_len, n_crops, dtype=torch.bfloat16)
        cb.step()  # first touch / oneDNN primitive creation
        s, L, parts = cb.step()
        out.update(value=L / s, cores=cb.threads, dtype="bf16", sample=cb.describe(s, L, parts))
        threads = cb.threads
        del cb
        c32 = CpuBaseline(cfg, prompt_len, n_crops, dtype=None, threads=threads)
        s, L, parts = c32.step()
        out.update(value_fp32=L / s, sample_fp32=c32.describe(s, L, parts))
    except Exception as e:  # noqa: BLE001
        out.setdefault("value", None)
        out.setdefault("cores", os.cpu_count())
        out["sample"] = out.get("sample", "") + f" failed: {e!r}"
    return out


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------
def run_reference_arm(args, cfg):
    """The reference's CPU implementation of the path on this box's host cores.  The reference is a Python repo that
    cannot travel to the GPU box (/root/reference is absent there), so this times the CPU oracle port of its algorithm
    (oracle/slime_oracle.py, pinned to the reference's golden vectors): `--warmup` untimed and exactly `--steps` timed
    steps, a step = ONE sample of the workload's batch (samples are independent) executed in full, in the CPU's
    fastest dtype (bf16 where the host has AMX / AVX512-BF16 - 2-3x faster than fp32; the fp32 figure is measured
    once and reported next to it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cb = CpuBaseline(cfg, args.prompt_len, args.crops, dtype=torch.bfloat16)
    for _ in range(args.warmup):
        cb.step()
    secs, toks, parts = [], 0, None
    for _ in range(args.steps):
        s, L, parts = cb.step()
        secs.append(s)
        toks += L
    total = sum(secs)
    value = toks / total
    desc = cb.describe(secs[-1], L, parts)
    threads = cb.threads
    del cb
    fp32_value, fp32_desc = None, None
    try:
        c32 = CpuBaseline(cfg, args.prompt_len, args.crops, dtype=None, threads=threads)
        s, L32, p32 = c32.step()
        fp32_value, fp32_desc = L32 / s, c32.describe(s, L32, p32)
    except Exception as e:  # noqa: BLE001
        fp32_desc = f"failed: {e!r}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, cfg),
        "step_samples": 1,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "dtype": "bf16",
                         "sample": f"each of the {args.steps} timed steps = " + desc,
                         "value_fp32": fp32_value, "sample_fp32": fp32_desc},
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
# GPU arm
# --------------------------------------------------------------------------------------------
class Bench:
    """One engine + one synthetic batch shard per rank, and the timing loops over it."""

    def __init__(self, torch, dist, eng, cfg, B, crops, T, rank, world, dev, flat=False):
        from slime_b200.parallel import LogitsGather
        from slime_b200.synth import grid_for_crops, synth_inputs

        self.torch, self.dist, self.eng, self.cfg = torch, dist, eng, cfg
        self.B, self.crops, self.T, self.rank, self.world, self.dev = B, crops, T, rank, world, dev
        self.grids = None if flat else [grid_for_crops(crops - 1)] * B
        self.set_inputs(3407 + rank)
        self.gather = LogitsGather(B, cfg.vocab_size, dev) if world > 1 else None
        self.logits_h = [torch.empty(world * B, cfg.vocab_size, dtype=torch.float32).pin_memory() for _ in range(2)]

    def set_inputs(self, seed):
        from slime_b200.synth import synth_inputs

        px_h, ids_h, mask_h = synth_inputs(self.cfg, self.B, self.crops, self.T, seed=seed)
        self.px_h = px_h.to(self.eng.dtype).pin_memory()
        self.ids_h, self.mask_h = ids_h.pin_memory(), mask_h.pin_memory()
        self.px_d, self.ids_d, self.mask_d = self.px_h.to(self.dev), self.ids_h.to(self.dev), self.mask_h.to(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def step_device(self):
        res = self.eng.prefill(self.px_d, self.ids_d, self.mask_d, grids=self.grids)
        if self.gather is not None:  # the one exchange step of the path, on the side stream
            self.last_slot = self.gather.submit(res.logits_last)
        return res

    def step_e2e(self):
        """Pinned host pixels + ids -> H2D -> SlimeEngine.prefill -> (all-gather) -> D2H of the result logits; the host
        reads step k's logits while step k+1 is in flight (double-buffered pinned result buffers)."""
        torch = self.torch
        px = self.px_h.to(self.dev, non_blocking=True)
        ids = self.ids_h.to(self.dev, non_blocking=True)
        mask = self.mask_h.to(self.dev, non_blocking=True)
        res = self.eng.prefill(px, ids, mask, grids=self.grids)
        k = getattr(self, "_e2e_k", 0)
        self._e2e_k = k + 1
        prev = getattr(self, "_e2e_ev", None)
        if self.gather is not None:
            slot = self.gather.submit(res.logits_last)
            with torch.cuda.stream(self.gather.stream):
                self.logits_h[k & 1].copy_(self.gather.bufs[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.gather.stream)
        else:
            self.logits_h[k & 1].copy_(res.logits_last, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
        if prev is not None:
            prev.synchronize()  # the host holds step k-1's logits now
        self._e2e_ev = ev
        return res

    def timed(self, fn, steps):
        torch, dist = self.torch, self.dist
        lib = self.eng.lib
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.slime_launch_count()
        e0.record()
        toks = 0
        for _ in range(steps):
            toks += fn().total_tokens
        if self.gather is not None:
            self.gather.drain()  # every gather (and result copy) issued in the region is inside the timing
        ev = getattr(self, "_e2e_ev", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.slime_launch_count() - launches0
        if self.world > 1:
            t = torch.tensor([ms, float(toks)], dtype=torch.float64, device=self.dev)
            tmax, tsum = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, toks = float(tmax[0]), float(tsum[1])
        return ms, toks, launches

    def gather_check(self):
        """Un-timed, N > 1 (VERDICT r1 row i; the reference's multi-GPU story is N independent processes whose outputs
        are concatenated, scripts/llama/eval/gqa.sh:20-43).  (1) every rank computes the SAME shard: all gathered row
        blocks must be bit-identical;  (2) each rank computes its own shard, rank 0 then re-computes rank 1's shard
        on its own GPU and compares it bit for bit with block 1 of the gather."""
        torch = self.torch
        if self.gather is None:
            return None
        B = self.B
        self.set_inputs(3407)  # the same inputs everywhere
        res = self.step_device()
        g = self.gather.result(self.last_slot).clone()
        same = all(torch.equal(g[:B], g[r * B:(r + 1) * B]) for r in range(1, self.world)) and torch.equal(g[:B], res.logits_last)
        self.set_inputs(3407 + self.rank)  # own shard again
        res = self.step_device()
        g = self.gather.result(self.last_slot).clone()
        own = torch.equal(g[self.rank * B:(self.rank + 1) * B], res.logits_last)
        cross = True
        if self.rank == 0:
            self.set_inputs(3407 + 1)
            r1 = self.eng.prefill(self.px_d, self.ids_d, self.mask_d, grids=self.grids)
            cross = torch.equal(g[B:2 * B], r1.logits_last)
            self.set_inputs(3407 + self.rank)
        ok = torch.tensor([int(same and own and cross)], device=self.dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)
        torch.cuda.synchronize()
        return "ok" if int(ok) == 1 else f"MISMATCH (same-shard {same}, own-block {own}, rank0-recomputes-rank1 {cross})"


def short_run(torch, dist, eng, cfg, B, crops, T, rank, world, dev, steps, flat=False, warm=3):
    """A short secondary measurement on one engine: (tokens/s whole job, ms per step, whole-step TFLOP/s per GPU)."""
    b = Bench(torch, dist, eng, cfg, B, crops, T, rank, world, dev, flat=flat)
    for _ in range(warm):
        res = b.step_device()
    ms, toks, _ = b.timed(b.step_device, steps)
    fl = algorithmic_flops(cfg, crops * B, res.lengths, (crops - 1) * B * cfg.mm_resampler_dim)
    step_ms = ms / steps
    return dict(tokens_per_s=toks / (ms / 1e3), ms_per_step=step_ms, batch_per_gpu=B, crops=crops, prompt_len=T,
                whole_step_tflops_per_gpu=fl["total"] / (step_ms / 1e3) / 1e12,
                mean_len=float(statistics.mean(res.lengths)))


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the short secondary measurements")
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

    # NCCL's log (NCCL_DEBUG=INFO from the driver) must not land between stdout's one JSON line: send it to stderr
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"

    import ctypes as C

    import torch
    import torch.distributed as dist

    from slime_b200.engine import SlimeEngine
    from slime_b200.synth import synth_tensor, weight_specs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def make_engine(c, dtype=torch.bfloat16):
        specs = {name: (shape, kind) for name, shape, kind in weight_specs(c)}
        e = SlimeEngine(c, local_rank, max_pos=4096, dtype=dtype)
        e.load_weights(lambda name: synth_tensor(name, specs[name][0], specs[name][1], 3407, device=dev, dtype=torch.bfloat16))
        return e

    # ---- model: random-init weights of the named architecture, generated on the GPU ----
    eng = make_engine(cfg)
    lib = eng.lib
    B = args.batch
    bench = Bench(torch, dist, eng, cfg, B, args.crops, args.prompt_len, rank, world, dev)

    for _ in range(max(args.warmup, 3)):
        res = bench.step_device()
    torch.cuda.synchronize()
    lengths = res.lengths

    # ---- timed region 1: device-resident inputs (value) with the library's event profiler on ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.slime_profile_enable(1)
    ms, toks, launches = bench.timed(bench.step_device, args.steps)
    pms, pwork, pl = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_longlong * 3)()
    lib.slime_profile_collect(pms, pwork, pl)
    lib.slime_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = toks / (ms / 1e3)

    # ---- timed region 2: end to end from pinned host memory through the public API ----
    for _ in range(2):
        bench.step_e2e()
    ms_e2e, toks_e2e, _ = bench.timed(bench.step_e2e, args.steps)
    e2e_value = toks_e2e / (ms_e2e / 1e3)
    h2d = bench.px_h.numel() * bench.px_h.element_size() + bench.ids_h.numel() * 8 + bench.mask_h.numel() * 8
    d2h = bench.logits_h[0].numel() * 4 + (B + 1) * 4
    bench._e2e_ev = None

    # ---- un-timed: the gathered logits equal what a single GPU computes (bit for bit) ----
    gather_check = bench.gather_check()

    # ---- timed region 3: from raw RGB bytes - GPU pre-processing (process_images) + prefill ----
    raw = None
    if args.crops == 5:
        import numpy as np

        from slime_b200.preprocess import preprocess_images

        rng = np.random.default_rng(3407 + rank)
        imgs = [rng.integers(0, 256, (672, 672, 3), dtype=np.uint8) for _ in range(B)]  # 672^2 -> 1 + 2x2 crops
        host_out = torch.empty(B, cfg.vocab_size, dtype=torch.float32).pin_memory()

        def step_raw():
            crops, _ = preprocess_images(imgs, "anyres", dtype=torch.bfloat16, device=dev)
            ids = bench.ids_h.to(dev, non_blocking=True)
            mask = bench.mask_h.to(dev, non_blocking=True)
            r = eng.prefill(torch.stack(crops), ids, mask, grids=bench.grids)
            if bench.gather is not None:
                bench.gather.submit(r.logits_last)
            host_out.copy_(r.logits_last, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return r

        for _ in range(2):
            step_raw()
        ms_raw, toks_raw, _ = bench.timed(step_raw, args.steps)
        raw = {"value": toks_raw / (ms_raw / 1e3), "unit": UNIT, "ms_per_step": ms_raw / args.steps,
               "h2d_bytes_per_step": B * 672 * 672 * 3 + bench.ids_h.numel() * 8 + bench.mask_h.numel() * 8,
               "note": "raw 672x672 RGB bytes on the host -> slime_preprocess_fwd (Pillow-exact resize, tiling, "
                       "CLIP normalise) -> prefill -> logits on the host"}

    # ---- ViT crops/s (secondary metric of BASELINE.json) ----
    flat = bench.px_d.flatten(0, 1)
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
    def decode_timing(e, c, px, ids, mask, grids, n_dec=16, fused_mask=-1):
        try:
            e.lib.slime_set_decode_fused(fused_mask)
            nb = px.shape[0]
            r0 = e.prefill(px, ids, mask, grids=grids)
            e.attach_kv_cache(nb, max(r0.lengths) + n_dec + 8)
            res_d = e.prefill(px, ids, mask, grids=grids)  # fills the cache
            lens_d = torch.tensor(res_d.lengths, dtype=torch.int32, device=dev)
            x_d = e.weights["llm.embed"][res_d.logits_last.argmax(-1)]
            for _ in range(3):
                e.decode_step(x_d, lens_d)  # (same slot rewritten: timing only)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = e.lib.slime_launch_count()
            t0.record()
            for _ in range(n_dec):
                e.decode_step(x_d, lens_d)
            t1.record()
            torch.cuda.synchronize()
            dms = t0.elapsed_time(t1) / n_dec
            w_bytes = 2 * (c.num_hidden_layers * (c.qkv_dim * c.hidden_size + c.hidden_size * c.num_attention_heads * c.head_dim
                                                  + 3 * c.intermediate_size * c.hidden_size) + c.vocab_size * c.hidden_size)
            kv_bytes = c.num_hidden_layers * sum(res_d.lengths) * 2 * c.num_key_value_heads * c.head_dim * 2
            return {"batch": nb, "ms_per_step": dms, "tokens_per_s": nb / dms * 1e3 * world,
                    "context": float(statistics.mean(res_d.lengths)),
                    "launches_per_step": (e.lib.slime_launch_count() - n0) / n_dec,
                    "hbm_bytes_per_step": w_bytes + kv_bytes, "achieved_gbs": (w_bytes + kv_bytes) / dms / 1e6}
        except Exception as ex:  # noqa: BLE001 - never lose the prefill line over a secondary figure
            return {"error": repr(ex)}
        finally:
            e.lib.slime_set_decode_fused(-1)
            try:
                e.detach_kv_cache()
            except Exception:  # noqa: BLE001
                pass

    decode = decode_timing(eng, cfg, bench.px_d, bench.ids_d, bench.mask_d, bench.grids)
    decode_b1 = decode_timing(eng, cfg, bench.px_d[:1], bench.ids_d[:1], bench.mask_d[:1], bench.grids[:1])
    # the 5-launches-per-layer chain (split reductions finished inside the producing kernels; opt-in because it measures
    # slower than the PDL chain with finishing kernels, profiles/r02_decode_experiments.txt) for the record
    decode_fused = {f"batch_{nb}": {k: v for k, v in decode_timing(eng, cfg, bench.px_d[:nb], bench.ids_d[:nb], bench.mask_d[:nb],
                                                                   bench.grids[:nb], fused_mask=7).items()
                                    if k in ("ms_per_step", "launches_per_step", "error")} for nb in (1, bench.px_d.shape[0])}

    # ---- secondary measurements (short step counts; same kernels, other modes / BASELINE.json configurations) ----
    peaks = measured_peaks()
    secondary = {}
    if not args.no_secondary and args.layers is None:
        def guarded(name, fn):
            try:
                secondary[name] = fn()
            except Exception as ex:  # noqa: BLE001
                secondary[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()

        def lat(e, c, crops, T):
            r = short_run(torch, None, e, c, 1, crops, T, rank, 1, dev, steps=20)
            out = {"ms": r["ms_per_step"], "tokens_per_s": r["tokens_per_s"],
                   "whole_step_frac": r["whole_step_tflops_per_gpu"] / peaks["tf_sustained"], "mean_len": r["mean_len"]}
            try:  # the same step captured in a CUDA graph (slime_b200.engine.GraphedPrefill: no host sync, one graph launch)
                from slime_b200.engine import GraphedPrefill
                b1 = Bench(torch, None, e, c, 1, crops, T, rank, 1, dev)
                g = GraphedPrefill(e, 1, crops, T, grids=b1.grids)
                for _ in range(3):
                    g(b1.px_d, b1.ids_d, b1.mask_d)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                t0.record()
                for _ in range(20):
                    g(b1.px_d, b1.ids_d, b1.mask_d)
                t1.record()
                torch.cuda.synchronize()
                out["ms_graph"] = t0.elapsed_time(t1) / 20
                del g
            except Exception as ex:  # noqa: BLE001
                out["ms_graph_error"] = repr(ex)
            return out

        # batch-1 latency (the reference's eval loop is batch 1, llava/eval/model_vqa_loader.py:103-119)
        guarded("latency_b1", lambda: {"headline_llama3_8b_T256": lat(eng, cfg, args.crops, args.prompt_len)})
        # fixed-length mode: top-p 1.0 keeps (almost) every local token (SURVEY.md 8d defines the roofline figure there)
        def topp1():
            e1 = eng.clone(mm_resampler_topp=1.0)
            r = short_run(torch, dist, e1, e1.cfg, B, args.crops, args.prompt_len, rank, world, dev, steps=4)
            r["whole_step_frac"] = r["whole_step_tflops_per_gpu"] / peaks["tf_sustained"]
            return r
        guarded("topp_1.0", topp1)
        if world > 1 and args.model == "llama3-8b":
            # BASELINE.json config 4 at its per-GPU shape (batch 32 over 8 GPUs): 8B, 8-frame video = 8 crops (flat merge),
            # T 256, 4 samples / GPU - at 8 GPUs this IS config 4, at fewer GPUs the same per-GPU work (weak scaling)
            def cfg4():
                e4 = eng.clone(mm_patch_merge_type="flat")
                return short_run(torch, dist, e4, e4.cfg, 4, 8, 256, rank, world, dev, steps=6, flat=True)
            guarded(f"config4_8b_8crops_b4_per_gpu_dp{world}", cfg4)
        # the float16 build (the reference's inference dtype, llava/model/builder.py:43) on the headline shape
        def fp16():
            e16 = make_engine(cfg, torch.float16)
            r = short_run(torch, dist, e16, cfg, B, args.crops, args.prompt_len, rank, world, dev, steps=4)
            r["whole_step_frac"] = r["whole_step_tflops_per_gpu"] / peaks["tf_sustained"]
            e16.close()
            return r
        guarded("fp16_build", fp16)
        # the other models need their own weights: release the 8B first
        if args.model == "llama3-8b":
            bench = None
            eng.close()
            del eng
            torch.cuda.empty_cache()
            def cfg2():
                c7 = preset("vicuna-7b")
                e7 = make_engine(c7)
                out = lat(e7, c7, 5, 128)
                e7.close()
                return out
            guarded("latency_b1_config2_vicuna7b_T128", cfg2)
            if world > 1:
                # BASELINE.json config 5 at its per-GPU shape: Vicuna-13B, 1344 px = 17 crops (flat), T 512, batch 64 over
                # 8 GPUs = 8 / GPU
                def cfg5():
                    c13 = preset("vicuna-13b", mm_patch_merge_type="flat")
                    e13 = make_engine(c13)
                    out = short_run(torch, dist, e13, c13, 8, 17, 512, rank, world, dev, steps=3, flat=True, warm=2)
                    e13.close()
                    return out
                guarded(f"config5_13b_17crops_b8_per_gpu_dp{world}", cfg5)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    for d in (decode, decode_b1):
        if d is not None and "achieved_gbs" in d:
            d["frac_of_hbm_peak"] = d["achieved_gbs"] / peaks["hbm"]
    if "achieved_gbs" in decode:
        decode["kernels"] = "weight-streaming mma.sync GEMM (csrc/gemm_skinny.cu) + split-KV mma.sync attention (csrc/decode_attn.cu) + finishing kernels, PDL"
        decode["batch_1"] = decode_b1
        decode["fused_chain_5_launches_per_layer"] = decode_fused
    fl = algorithmic_flops(cfg, args.crops * B, lengths, (args.crops - 1) * B * cfg.mm_resampler_dim)
    gemm_tf = pwork[0] / (pms[0] / 1e3) / 1e12 if pms[0] > 0 else 0.0
    step_ms = ms / args.steps
    traffic, traffic_note = None, None
    try:  # DRAM bytes per launch of the dominant GEMM: NOT measured by this run (ncu cannot run inside a timed bench) -
        # the committed `ncu --set full` capture of the same kernel on the same shape (tools/ncu_summary.py)
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f)["gemm"]
        traffic = tr["dram_bytes_per_launch"]
        traffic_note = (f"static: read from profiles/roofline_traffic.json (ncu --set full capture, {tr['launches']} "
                        f"launch(es), {tr['source']}), not measured in this run" + (f"; {tr['note']}" if tr.get("note") else ""))
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
        "whole_step_tflops": fl["total"] / (step_ms / 1e3) / 1e12 * world,
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
        "secondary": secondary,
    }
    if world > 1:
        line["gather_check"] = gather_check
        line["gather"] = "all_gather of [B_local, V] fp32 last-token logits on a side stream (slime_b200/parallel.py LogitsGather)"
    if "latency_b1" in secondary:
        line["latency_b1"] = dict(secondary["latency_b1"])
        if "latency_b1_config2_vicuna7b_T128" in secondary:
            line["latency_b1"]["config2_vicuna7b_T128"] = secondary["latency_b1_config2_vicuna7b_T128"]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(cfg, args.prompt_len, args.crops)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
