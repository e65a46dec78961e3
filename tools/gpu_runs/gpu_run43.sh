#!/bin/bash
# round-1 closing run (final state): full GPU suite, smoke, default bench, the other BASELINE configurations at their per-GPU shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  raw {d["e2e_from_rgb_bytes"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz  whole {r["whole_step_frac_of_peak"]:.3f}')
print("cpu:", {k: d["cpu_baseline"].get(k) for k in ("value", "value_bf16", "cores")}, " decode:", {k: d["decode_step"].get(k) for k in ("ms_per_step", "tokens_per_s", "frac_of_hbm_peak")})
PY
cfgs=("--model vicuna-7b --crops 5 --prompt-len 128 --batch 1" "--model llama3-8b --crops 10 --prompt-len 256 --batch 8" "--model llama3-8b --crops 8 --prompt-len 256 --batch 4" "--model vicuna-13b --crops 17 --prompt-len 512 --batch 8" "--model llama3-8b --crops 5 --prompt-len 256 --batch 1")
i=2
for c in "${cfgs[@]}"; do
  timeout 400 python bench.py $c --steps 8 --no-cpu-baseline > gpurun_out/bench_cfg_$i.json 2> gpurun_out/bench_cfg.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg_$i.json")); r=d["roofline"]
print("cfg: $c |", f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms/step  tokens/step {d["tokens_per_step"]:.0f}  gemm {r["achieved"]:.0f} TF/s  whole {r["whole_step_frac_of_peak"]:.3f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz  decode {d["decode_step"].get("ms_per_step")}')
PY
  i=$((i+1))
done
