#!/bin/bash
# Run D: kv-split softmax (two groups of 4 warps alternate kv tiles) - parity, A/B timing, trace, bench
mkdir -p gpurun_out
SLIME_ATTN_VARIANT=21 timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_kvsplit.log 2>&1; echo "pytest attention (variant 21) rc=$?"; tail -12 gpurun_out/pytest_kvsplit.log | cut -c1-400
SLIME_ATTN_VARIANT=21 timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_variants_gpu.py tests/test_decode_gpu.py tests/test_fp16_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_kvsplit2.log 2>&1; echo "pytest stages/variants/decode/fp16 (variant 21) rc=$?"; tail -6 gpurun_out/pytest_kvsplit2.log | cut -c1-400
timeout 300 python tools/ab_kernels.py attn > gpurun_out/ab_attn.log 2>&1; echo "ab rc=$?"; cat gpurun_out/ab_attn.log
SLIME_ATTN_VARIANT=21 timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace_v21.log 2>&1; echo "trace v21 rc=$?"; grep -A16 "per-tile deltas" gpurun_out/attn_trace_v21.log; head -30 gpurun_out/attn_trace_v21.log
for var in 5 21 25; do SLIME_ATTN_VARIANT=$var timeout 400 python bench.py --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_var$var.json 2> gpurun_out/bench_var$var.err; echo "bench var=$var rc=$?"; tail -2 gpurun_out/bench_var$var.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_var$var.json")); r=d["roofline"]
print(f'  {d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
