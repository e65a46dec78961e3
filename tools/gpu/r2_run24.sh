#!/bin/bash
# round 2: DRAM traffic / time of the other decoder GEMMs vs rasterisation group size
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
run() {
  echo "-- $1 SLIME_GEMM_GROUP_ROWS=$2"
  SLIME_GEMM_GROUP_ROWS=$2 timeout 120 python tools/prof_gemm.py $1 2>&1 | tail -1
  SLIME_GEMM_GROUP_ROWS=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py $1 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -5
}
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for rows in 0 2048 2304 3072 4096; do run down $rows; done
for rows in 0 2048 8192; do run o $rows; done
for rows in 0 2048 8192; do run qkv $rows; done
} 2>&1 | tee gpurun_out/r2_gemm_other_shapes.log
