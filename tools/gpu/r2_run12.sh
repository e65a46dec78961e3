#!/bin/bash
# round 2: evidence run on one B200 - full GPU suite (twice), smoke(), ncu --set full of the dominant kernels, launch list of
# one bench step, bench.py as the driver runs it, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "== pytest -m gpu (run 1)"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_1.log
echo "== pytest -m gpu (run 2)"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_2.log
echo "== measured stage errors (small configurations)"
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_variants_gpu.py tests/test_qformer_router_gpu.py tests/test_shims_gpu.py -q -s -m gpu 2>&1 | grep -E "rel-L2" | sed 's/^\.*//' > gpurun_out/r2_stage_errors.txt
sort -t: -k2 -g gpurun_out/r2_stage_errors.txt | awk '{print}' | tail -12
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== ncu --set full"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 -o gpurun_out/r2_ncu_gemm_gate_up -f python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm.log 2>&1; tail -1 gpurun_out/r2_ncu_gemm.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_tn_2cta -s 6 -c 1 -o gpurun_out/r2_ncu_gemm_vit_fc1 -f python tools/prof_gemm_vit.py > gpurun_out/r2_ncu_gemm_vit.log 2>&1; tail -1 gpurun_out/r2_ncu_gemm_vit.log
for which in vit decoder; do
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:attn2q -s 3 -c 1 -o gpurun_out/r2_ncu_attn2q_$which -f python tools/prof_attn.py $which > gpurun_out/r2_ncu_attn_$which.log 2>&1; tail -1 gpurun_out/r2_ncu_attn_$which.log
done
echo "== launch list of one headline step (batch 16)"
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b16.csv python tools/step_profile.py --model llama3-8b --batch 16 > gpurun_out/r2_launches_b16.log 2>&1; tail -1 gpurun_out/r2_launches_b16.log
echo "== bench.py as the driver runs it"
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n1.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s, frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  whole {r["whole_step_frac_of_peak"]:.3f}  sm {d["clocks"]["sm_mhz"]} MHz {d["clocks"]["reasons"]}')
print(json.dumps(d.get("latency_b1")))
print(json.dumps(d.get("cpu_baseline"))[:600])
print(json.dumps(d.get("decode_step"))[:700])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cut -c1-900 gpurun_out/r2_bench_ref.json
