#!/bin/bash
# round 2: SM -> die probe; DRAM traffic of gate/up and down with die-aware tile lists (static schedule, isolated launches)
# (needs the experimental csrc/topology.cu + die-aware tile lists described in profiles/r02_gemm_experiments.txt item 9; not kept)
mkdir -p gpurun_out
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
timeout 120 python tools/probe_dies.py
timeout 120 python tools/probe_dies.py | head -1
for which in gate_up down; do
for da in 0 1; do
  echo "-- $which SLIME_GEMM_DIE_AWARE=$da"
  SLIME_GEMM_DIE_AWARE=$da timeout 120 python tools/prof_gemm.py $which 2>&1 | tail -1
  SLIME_GEMM_DIE_AWARE=$da timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py $which 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
done
done
echo "-- down, die aware, group rows 2048"
SLIME_GEMM_GROUP_ROWS=2048 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py down 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
} 2>&1 | tee gpurun_out/r2_gemm_die_aware.log
echo "== kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k gemm 2>&1 | tail -2
