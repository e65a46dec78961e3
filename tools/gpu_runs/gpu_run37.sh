#!/bin/bash
# round-1 closing run: full GPU suite, default bench (prefill + decode figure), decode bench with isolated projections,
# ncu launch list + full captures of the decode kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -14 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  raw {d["e2e_from_rgb_bytes"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f} ms  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches {d["gpu_launches"]}  sm {d["clocks"]["sm_mhz"]} MHz  whole {r["whole_step_frac_of_peak"]:.3f} cpu {d.get("cpu_baseline",{}).get("value")}')
print("decode:", d.get("decode_step"))
PY
timeout 600 python tools/bench_decode.py --batches 1,4,16,32 --steps 32 > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"; grep '"batch"' gpurun_out/decode_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", d['kernels'][:80])
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"skinny|decode_attn|rmsnorm" -c 120 --csv --log-file gpurun_out/decode_launches.csv python tools/bench_decode.py --layers 2 --batches 1,16 --steps 1 --no-projections --quick > gpurun_out/decode_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny_kernel" -s 6 -c 6 -o gpurun_out/prof_skinny -f python tools/bench_decode.py --layers 1 --batches 16 --steps 1 --no-projections --quick > gpurun_out/ncu_skinny.out 2>&1; echo "ncu skinny rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"decode_attn_mma_kernel|skinny_finish_norm" -s 2 -c 3 -o gpurun_out/prof_decode_attn -f python tools/bench_decode.py --layers 1 --batches 16 --steps 1 --no-projections --quick > gpurun_out/ncu_dattn.out 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
