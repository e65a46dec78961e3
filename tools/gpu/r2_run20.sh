#!/bin/bash
# round 2: box survey - DRAM traffic / time of the gate/up GEMM vs rasterisation group size on whatever box this call gets
mkdir -p gpurun_out
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,serial,uuid,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for rows in 1024 2048 4096 6144; do
  echo "-- SLIME_GEMM_GROUP_ROWS=$rows"
  SLIME_GEMM_GROUP_ROWS=$rows timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  SLIME_GEMM_GROUP_ROWS=$rows timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -5
done
} 2>&1 | tee gpurun_out/r2_box_survey_$(date +%H%M%S).log
