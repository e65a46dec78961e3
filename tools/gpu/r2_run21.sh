#!/bin/bash
# round 2: DRAM traffic of gate/up vs (a) evict_first output stores, (b) serpentine n-sweep; bench for each
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "-- SLIME_GEMM_STORE_HINT=$1 SLIME_GEMM_SNAKE=$2"
  SLIME_GEMM_STORE_HINT=$1 SLIME_GEMM_SNAKE=$2 timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  SLIME_GEMM_STORE_HINT=$1 SLIME_GEMM_SNAKE=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -5
done
echo "-- hint 1 snake 1 group 6144"
SLIME_GEMM_GROUP_ROWS=6144 SLIME_GEMM_STORE_HINT=1 SLIME_GEMM_SNAKE=1 timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
SLIME_GEMM_GROUP_ROWS=6144 SLIME_GEMM_STORE_HINT=1 SLIME_GEMM_SNAKE=1 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -5
} 2>&1 | tee gpurun_out/r2_gemm_store_hint.log
echo "== kernel tests with both on"
SLIME_GEMM_STORE_HINT=1 SLIME_GEMM_SNAKE=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k gemm 2>&1 | tail -2
echo "== bench"
for cfg in "0 0" "1 1" "1 0" "0 0" "1 1"; do
set -- $cfg
SLIME_GEMM_STORE_HINT=$1 SLIME_GEMM_SNAKE=$2 timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_hint.json 2> gpurun_out/r2_bench_hint.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_hint.json")); r=d["roofline"]
print("hint $1 snake $2", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
