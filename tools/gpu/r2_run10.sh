#!/bin/bash
# round 2, call 10: half-width 2-CTA GEMM tiles for nearly empty last waves (batch-1 latency), norm fold at batch 1, poly shares
mkdir -p gpurun_out
echo "== kernel tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x 2>&1 | tail -3
SLIME_GEMM2_BN=128 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm or rope" 2>&1 | tail -3
echo "== attention polys"
timeout 600 python tools/attn_bench.py --polys 0,2,3,4 > gpurun_out/r2_attn_bench7.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench7.log | head -40
echo "== batch-1 latency A/B (graph test prints eager / graph ms)"
for cfg in "256 1" "0 1" "0 0"; do
  set -- $cfg
  echo "-- SLIME_GEMM2_BN=$1 SLIME_NORM_FOLD=$2"
  SLIME_GEMM2_BN=$1 SLIME_NORM_FOLD=$2 timeout 600 python -m pytest tests/test_graph_gpu.py -q -s -m gpu -k fullsize 2>&1 | grep -E "latency|passed|failed" 
done
echo "== GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_gpu.py 2>&1 | tail -4 | tee gpurun_out/r2_suite.log
echo "== bench"
timeout 900 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_b.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
print(json.dumps(d.get("latency_b1"), indent=1))
PY
