#!/bin/bash
# round 2: does a start-up stagger of the GEMM clusters (operand sharers one after the other instead of in k-lockstep) cut the
# DRAM re-reads of gate/up?  DRAM bytes from a metrics-only ncu pass, time from CUDA events outside ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
run() {  # label, env...
  echo "-- $*"
  env "$@" timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  env "$@" timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time|per_second" | head -6
}
{
run SLIME_GEMM_STAGGER_NS=0
run SLIME_GEMM_STAGGER_NS=200
run SLIME_GEMM_STAGGER_NS=500
run SLIME_GEMM_STAGGER_NS=1000
run SLIME_GEMM_STAGGER_NS=2000
run SLIME_GEMM_STAGGER_NS=500 SLIME_GEMM_GROUP_ROWS=8192
run SLIME_GEMM_STAGGER_NS=0 SLIME_GEMM_GROUP_ROWS=8192
run SLIME_GEMM_STAGGER_NS=500 SLIME_GEMM_GROUP_ROWS=2048
} 2>&1 | tee gpurun_out/r2_gemm_stagger.log
echo "== --set full on the same box (stagger 0): does the full set report the same DRAM bytes as the metrics pass?"
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes_read.sum |dram__bytes_write.sum |DRAM Throughput|Duration|L2 Hit" | head -8 | tee gpurun_out/r2_gemm_setfull_check.log
echo "== bench, stagger 0 / 500 / 1000"
for ns in 0 500 1000; do
SLIME_GEMM_STAGGER_NS=$ns timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_stagger$ns.json 2> gpurun_out/r2_bench_stagger$ns.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_stagger$ns.json")); r=d["roofline"]
print("stagger $ns", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
