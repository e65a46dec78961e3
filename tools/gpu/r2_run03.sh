#!/bin/bash
# round 2, call 4: warp-uniform MMA issue (attention_tc2 + both GEMM kernels): kernel tests, attention bench + trace, GEMM timings, bench A/B
mkdir -p gpurun_out
echo "== kernel tests (gemm + attention)"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2_kernel_tests.log
echo "== attention bench + trace"
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2,3,4 --trace > gpurun_out/r2_attn_bench2.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench2.log | head -40
grep -E "^tile (5|6|7|8|9|1[0-5]):" gpurun_out/r2_attn_bench2.log | head -24
echo "== gemm timings"
timeout 600 python tools/ab_kernels.py gemm 2>&1 | grep -E "^gemm" | cut -c1-150 | tee gpurun_out/r2_gemm_ab.log
echo "== bench A/B impl 2 vs 3"
for impl in 2 3; do
  SLIME_ATTN_IMPL=$impl timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_impl$impl.json 2> gpurun_out/r2_bench_impl$impl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_impl$impl.json")); r=d["roofline"]
print("impl $impl", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
echo "== GPU suite (impl 3), minus full-size"
SLIME_ATTN_IMPL=3 timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -4 | tee gpurun_out/r2_suite_impl3.log
