#!/bin/bash
# Run C: new kernel tests (staged epilogue bit-identity, softmax variants, fused QKV+RoPE), the whole GPU suite with
# SLIME_FUSED_ROPE=1, bench A/B of the fused RoPE, and the GEMM rasterisation group-size sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "staged or variants or qkv_rope" > gpurun_out/pytest_new_tests.log 2>&1; echo "pytest new kernel tests rc=$?"; tail -8 gpurun_out/pytest_new_tests.log | cut -c1-400
SLIME_FUSED_ROPE=1 timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_fused_rope.log 2>&1; echo "pytest all (fused rope) rc=$?"; tail -8 gpurun_out/pytest_fused_rope.log | cut -c1-400
for rows in 2048 4096 8192 16384 32768; do SLIME_GEMM_GROUP_ROWS=$rows timeout 120 python tools/prof_gemm.py 2>&1 | tail -1; done
for cfg in "0 0" "1 0" "1 8192" "1 16384"; do set -- $cfg; SLIME_FUSED_ROPE=$1 SLIME_GEMM_GROUP_ROWS=$2 timeout 400 python bench.py --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_rope$1_rows$2.json 2> gpurun_out/bench_rope$1_rows$2.err; echo "bench fused_rope=$1 group_rows=$2 rc=$?"; tail -2 gpurun_out/bench_rope$1_rows$2.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_rope$1_rows$2.json")); r=d["roofline"]
print(f'  {d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
