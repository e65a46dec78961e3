#!/bin/bash
# Run B: staged GEMM epilogue (SLIME_GEMM_EPI_MODE=1) and attention softmax variants (SLIME_ATTN_VARIANT) - parity, A/B timing, trace
mkdir -p gpurun_out
SLIME_GEMM_EPI_MODE=1 SLIME_ATTN_VARIANT=5 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_stages_gpu.py tests/test_variants_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_new_paths.log 2>&1; echo "pytest (epi mode 1, attn variant 5) rc=$?"; tail -6 gpurun_out/pytest_new_paths.log | cut -c1-300
SLIME_GEMM_EPI_MODE=1 SLIME_ATTN_VARIANT=9 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -x -q -m gpu -p no:cacheprovider -k "attention or gemm" > gpurun_out/pytest_new_paths2.log 2>&1; echo "pytest (epi mode 1, attn variant 9, + fp16) rc=$?"; tail -4 gpurun_out/pytest_new_paths2.log | cut -c1-300
timeout 600 python tools/ab_kernels.py > gpurun_out/ab_kernels.log 2>&1; echo "ab rc=$?"; cat gpurun_out/ab_kernels.log
SLIME_ATTN_VARIANT=0 timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace_v0.log 2>&1; echo "trace v0 rc=$?"; grep -A12 "per-tile deltas" gpurun_out/attn_trace_v0.log
SLIME_ATTN_VARIANT=5 timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace_v5.log 2>&1; echo "trace v5 rc=$?"; grep -A12 "per-tile deltas" gpurun_out/attn_trace_v5.log
SLIME_ATTN_VARIANT=9 timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace_v9.log 2>&1; echo "trace v9 rc=$?"; grep -A12 "per-tile deltas" gpurun_out/attn_trace_v9.log
for cfg in "0 0" "1 0" "1 5" "1 7"; do set -- $cfg; SLIME_GEMM_EPI_MODE=$1 SLIME_ATTN_VARIANT=$2 timeout 400 python bench.py --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_epi$1_var$2.json 2> gpurun_out/bench_epi$1_var$2.err; echo "bench epi=$1 var=$2 rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_epi$1_var$2.json")); r=d["roofline"]
print(f'  {d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
