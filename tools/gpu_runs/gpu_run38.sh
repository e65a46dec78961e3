#!/bin/bash
# GEMM evict-first output stores (SLIME_GEMM_STREAM_OUT_MB): DRAM bytes of the gate/up GEMM under ncu, isolated timing, bench A/B; smoke()
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for mb in 0 64; do
  echo "== SLIME_GEMM_STREAM_OUT_MB=$mb"
  SLIME_GEMM_STREAM_OUT_MB=$mb timeout 120 python tools/prof_gemm.py 2>&1 | tail -1
  SLIME_GEMM_STREAM_OUT_MB=$mb timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 2 python tools/prof_gemm.py 2>&1 | grep -E "dram__bytes|gpu__time" | head -6
done
for mb in 0 64 0 64; do
  SLIME_GEMM_STREAM_OUT_MB=$mb timeout 400 python bench.py --steps 6 --no-cpu-baseline > gpurun_out/bench_so_$mb.json 2> gpurun_out/bench_so.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_so_$mb.json")); r=d["roofline"]
print("stream_out_mb=$mb", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
