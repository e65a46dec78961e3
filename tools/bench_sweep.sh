#!/bin/bash
# per-GPU batch sweep {1, 8, 16, 32} at top-p 0.95, and the fixed-length mode (top-p 1.0) at the
# headline batch.  One JSON line per run in gpurun_out/bench_sweep.jsonl, a table on stdout.
mkdir -p gpurun_out
: > gpurun_out/bench_sweep.jsonl
for cfg in "--batch 1" "--batch 8" "--batch 16" "--batch 32" "--batch 16 --topp 1.0" "--model vicuna-7b --prompt-len 128 --batch 16"; do
  timeout 600 python bench.py --no-cpu-baseline --steps 6 --warmup 3 $cfg >> gpurun_out/bench_sweep.jsonl 2>> gpurun_out/bench_sweep.err
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_sweep.jsonl"):
    d = json.loads(l)
    c = d["config"]
    print(f'{c["model"]:16s} B={c["per_gpu_batch"]:<3d} T={c["prompt_len"]:<4d} top_p={c["top_p"]:<5} '
          f'{d["value"]:9.0f} tok/s  {d["ms_per_step"]:8.2f} ms/step  e2e {d["e2e"]["value"]:9.0f}  '
          f'vit {d["vit_crops_per_sec"]:6.0f} crops/s  gemm frac {d["roofline"]["frac"]:.3f}  '
          f'whole-step frac {d["roofline"]["whole_step_frac_of_peak"]:.3f}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
