#!/bin/bash
# round 2, call 8: attention_tc2 default (compile-time mask split), RMSNorm folding on by default: suite, full-size parity, A/B
mkdir -p gpurun_out
echo "== attention tests + bench"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2 --trace > gpurun_out/r2_attn_bench6.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench6.log | head -40
grep -E "^tile (5|6|7|8|9|1[0-5]):" gpurun_out/r2_attn_bench6.log | cut -c1-110 | head -24
echo "== GPU suite (defaults: impl 3, norm fold on), minus full-size"
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2_suite.log
echo "== fullsize parity (defaults)"
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -s -m gpu > gpurun_out/r2_fullsize_fold.log 2>&1
grep -E "rel-L2|16 bit|passed|failed|Error|assert" gpurun_out/r2_fullsize_fold.log | sed 's/^\.*//' | cut -c1-230
echo "== bench A/B norm fold"
for fold in 0 1; do
  SLIME_NORM_FOLD=$fold timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_fold$fold.json 2> gpurun_out/r2_bench_fold$fold.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_fold$fold.json")); r=d["roofline"]
print("fold $fold", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
