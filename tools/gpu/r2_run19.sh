#!/bin/bash
# round 2, closing evidence on one B200: GPU suite, smoke, ncu --set full of gate/up (traffic for roofline_traffic.json),
# launch list of one step, bench.py as the driver runs it, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_final.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2_smoke_final.log
echo "== ncu --set full, gate/up"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_tn_2cta -s 8 -c 1 -o gpurun_out/r2_ncu_gemm_gate_up_final -f python tools/prof_gemm.py > gpurun_out/r2_ncu_gemm_final.log 2>&1; tail -1 gpurun_out/r2_ncu_gemm_final.log
echo "== launch list of one headline step (batch 16)"
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b16_final.csv python tools/step_profile.py --model llama3-8b --batch 16 > gpurun_out/r2_launches_b16_final.log 2>&1; tail -1 gpurun_out/r2_launches_b16_final.log
echo "== bench.py as the driver runs it"
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
tail -c 300 gpurun_out/r2_bench_n1_final.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_n1_final.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s, frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  whole {r["whole_step_frac_of_peak"]:.3f}  sm {d["clocks"]["sm_mhz"]} MHz {d["clocks"]["reasons"]}')
print(json.dumps(d.get("latency_b1")))
print(json.dumps(d.get("decode_step"))[:1200])
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
cut -c1-300 gpurun_out/r2_bench_ref_final.json
