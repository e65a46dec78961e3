#!/bin/bash
# full GPU suite + default bench on the current defaults (attention variant 5, fused RoPE, pipelined bias loads)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  raw {d["e2e_from_rgb_bytes"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz  whole {r["whole_step_frac_of_peak"]:.3f} cpu {d.get("cpu_baseline",{}).get("value")}')
PY
