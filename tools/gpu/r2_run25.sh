#!/bin/bash
# round 2: is the DRAM traffic of the GEMM paid once per die?  Same GEMM on 74 / 37 / 18 clusters (cluster launches fill
# (drives SLIME_GEMM_MAX_CLUSTERS, an experiment knob that capped `max_clusters` in gemm2_sm100.cu launch2s; not kept)
# SM ids contiguously, clusters never straddle a die)
mkdir -p gpurun_out
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for which in gate_up down; do
for cl in 0 56 37 30 18; do
  echo "-- $which SLIME_GEMM_MAX_CLUSTERS=$cl"
  SLIME_GEMM_MAX_CLUSTERS=$cl timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py $which 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
done
done
} 2>&1 | tee gpurun_out/r2_gemm_half_grid.log
