#!/bin/bash
# round 2: launch list of one batch-1 step with the final build (prefill PDL, 192-column tiles for the N = 4096 projections)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b1_final.csv python tools/step_profile.py --model llama3-8b --batch 1 > gpurun_out/r2_launches_b1_final.log 2>&1; tail -1 gpurun_out/r2_launches_b1_final.log
