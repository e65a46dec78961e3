#!/bin/bash
# round 2, call 6: attention_tc2 with early S release at hd 64 (+ chunk masking, wait reordering); GEMM L2-hint DRAM traffic
mkdir -p gpurun_out
echo "== attention kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x -k attention 2>&1 | tail -4 | tee gpurun_out/r2_attn_tests.log
echo "== attention bench + trace"
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2,3,4 --trace > gpurun_out/r2_attn_bench4.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench4.log | head -40
grep -E "^tile (5|6|7|8|9|1[0-5]):" gpurun_out/r2_attn_bench4.log | head -24
echo "== GEMM gate/up DRAM traffic vs L2 hints"
for hint in 0 1 2 3; do
  echo "-- SLIME_GEMM_L2HINT=$hint"
  SLIME_GEMM_L2HINT=$hint timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  SLIME_GEMM_L2HINT=$hint timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time" | head -6
done 2>&1 | tee gpurun_out/r2_gemm_l2hint.log
echo "== GPU suite (impl 3), minus full-size"
SLIME_ATTN_IMPL=3 timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -4 | tee gpurun_out/r2_suite_impl3.log
echo "== bench (impl 3)"
SLIME_ATTN_IMPL=3 timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_impl3.json 2> gpurun_out/r2_bench_impl3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_impl3.json")); r=d["roofline"]
print("impl 3", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
