#!/bin/bash
# round 2, call 7: Veltkamp bf16 packing in the softmax (F2FP off the XU pipe); diagnose the decode test failure with impl 3
mkdir -p gpurun_out
echo "== decode test (impl 3)"
SLIME_ATTN_IMPL=3 timeout 600 python -m pytest tests/test_decode_gpu.py -q -m gpu -x -s 2>&1 | grep -v Warning | tail -30 | tee gpurun_out/r2_decode_test.log
echo "== attention kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x -k attention 2>&1 | tail -4
echo "== attention bench + trace"
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2,3,4 --trace > gpurun_out/r2_attn_bench5.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench5.log | head -40
grep -E "^tile (5|6|7|8|9|1[0-5]):" gpurun_out/r2_attn_bench5.log | cut -c1-110 | head -24
echo "== GPU suite (impl 3), minus full-size"
SLIME_ATTN_IMPL=3 timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2_suite_impl3.log
echo "== bench (impl 3)"
SLIME_ATTN_IMPL=3 timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_impl3.json 2> gpurun_out/r2_bench_impl3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_impl3.json")); r=d["roofline"]
print("impl 3", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
