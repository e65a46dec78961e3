#!/bin/bash
# round 2: DRAM traffic of gate/up vs the shape of a wave (sub-blocks of n-tiles inside an m-group)
mkdir -p gpurun_out
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for cfg in "1 4096" "2 4096" "4 4096" "9 4096" "4 6144" "9 6144" "4 8192" "18 4096" "1 4096"; do
  set -- $cfg
  echo "-- SLIME_GEMM_SUBN=$1 SLIME_GEMM_GROUP_ROWS=$2"
  SLIME_GEMM_SUBN=$1 SLIME_GEMM_GROUP_ROWS=$2 timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  SLIME_GEMM_SUBN=$1 SLIME_GEMM_GROUP_ROWS=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
done
} 2>&1 | tee gpurun_out/r2_gemm_subn.log
