#!/bin/bash
# round 2: L2 policies of the operand streams on the down-projection (W = 117 MB is re-read every wave, A streams)
# SLIME_GEMM_L2HINT = kind_A + 4 * kind_W + 16 * pct   (kind 1 evict_last, 2 evict_first, 3 evict_last for pct % / evict_first rest)
# (drives SLIME_GEMM_L2HINT, an experiment knob in gemm2_sm100.cu that is not kept: profiles/r02_gemm_experiments.txt item 11)
mkdir -p gpurun_out
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
{
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv,noheader
for h in 0 2 4 6 $((12+16*25)) $((12+16*33)) $((12+16*50)) $((2+12+16*33)) $((3+16*33)) 0; do
  echo "-- down SLIME_GEMM_L2HINT=$h  (A kind $((h&3)), W kind $(((h>>2)&3)), pct $((h>>4)))"
  SLIME_GEMM_L2HINT=$h timeout 120 python tools/prof_gemm.py down 2>&1 | tail -1
  SLIME_GEMM_L2HINT=$h timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py down 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
done
for h in $((12+16*33)) $((12+16*50)); do
  echo "-- gate_up SLIME_GEMM_L2HINT=$h"
  SLIME_GEMM_L2HINT=$h timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py gate_up 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -4
done
} 2>&1 | tee gpurun_out/r2_gemm_down_l2hint.log
