#!/bin/bash
# round 2, call 5: attention with alternating exp phases + chunk masking; sync-free / graphed prefill; ncu of attention; B=1 launch list
mkdir -p gpurun_out
echo "== attention kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py -q -m gpu -x -k attention 2>&1 | tail -4 | tee gpurun_out/r2_attn_tests.log
echo "== attention bench + trace"
timeout 600 python tools/attn_bench.py --impls 2,3 --polys 0,2,3,4 --trace > gpurun_out/r2_attn_bench3.log 2>&1
grep -E "^attn|Error|error" gpurun_out/r2_attn_bench3.log | head -40
grep -E "^tile (5|6|7|8|9|1[0-5]):" gpurun_out/r2_attn_bench3.log | head -24
echo "== sync-free / graph tests"
timeout 900 python -m pytest tests/test_graph_gpu.py -q -s -m gpu -x 2>&1 | grep -E "latency|passed|failed|Error|error" | head -20 | tee gpurun_out/r2_graph_tests.log
echo "== ncu attention (impl 3): vit + decoder"
for which in vit decoder; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn2q -s 3 -c 1 -o gpurun_out/r2_attn2q_$which -f python tools/prof_attn.py $which 3 > gpurun_out/r2_ncu_$which.log 2>&1
  tail -2 gpurun_out/r2_ncu_$which.log
done
echo "== B=1 launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b1.csv python tools/step_profile.py --model llama3-8b --batch 1 > gpurun_out/r2_launches_b1.log 2>&1
tail -2 gpurun_out/r2_launches_b1.log
echo "== bench (impl 3)"
SLIME_ATTN_IMPL=3 timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_impl3.json 2> gpurun_out/r2_bench_impl3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_impl3.json")); r=d["roofline"]
print("impl 3", f'{d["value"]:.0f} tok/s  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
PY
