#!/bin/bash
# round 2: bench A/B of the minimum rasterisation group (8 m-tiles for long reductions) - two pairs on one box
mkdir -p gpurun_out
for gm in 1 8 1 8; do
SLIME_GEMM_GROUP_MIN=$gm timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_gmin.json 2> gpurun_out/r2_bench_gmin.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_gmin.json")); r=d["roofline"]
print("group min $gm", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k "gemm or invariants" 2>&1 | tail -2
