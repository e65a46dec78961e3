#!/bin/bash
# round 2: full GPU suite + smoke + short bench after the in-place decoder input and the split vision-tower output
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_3.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench (short)"
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err
tail -c 400 gpurun_out/r2_bench_short.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_short.json")); r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s, frac {r["frac"]:.3f})  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  launches/step {d["gpu_launches"]/d["steps"]:.0f}  sm {d["clocks"]["sm_mhz"]} MHz')
print(json.dumps(d.get("latency_b1")))
print(json.dumps(d.get("decode_step"))[:900])
PY
