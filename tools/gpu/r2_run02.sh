#!/bin/bash
# round 2, call 2: first light of the two-query-tile attention kernel (impl 3) + full-size parity log
mkdir -p gpurun_out
echo "== attention kernel tests (impl 3 and 2)"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -k attention -x 2>&1 | tail -15 | tee gpurun_out/r2_attn_tests.log
echo "== attention bench + trace"
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2,3,4 --trace > gpurun_out/r2_attn_bench.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench.log | head -40
echo "== whole GPU suite with the two-tile kernel as the default (minus full-size)"
SLIME_ATTN_IMPL=3 timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -8 | tee gpurun_out/r2_suite_impl3.log
echo "== fullsize parity (impl 3)"
SLIME_ATTN_IMPL=3 timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -s -m gpu > gpurun_out/r2_fullsize_impl3.log 2>&1
grep -E "rel-L2|16 bit|passed|failed|Error" gpurun_out/r2_fullsize_impl3.log | cut -c1-250
echo "== bench A/B impl 2 vs 3"
for impl in 2 3; do
  SLIME_ATTN_IMPL=$impl timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_impl$impl.json 2> gpurun_out/r2_bench_impl$impl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_impl$impl.json")); r=d["roofline"]
print("impl $impl", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
