#!/bin/bash
# round 2, call 1: new full-size parity tests (configs 3/4/5, bf16 floor, full-size decode), tail-split GEMM validation,
# the rewritten bench.py at N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "== fullsize parity"
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -x -q -s -m gpu 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r2_fullsize.log
echo "== bench N=1"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 3000 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench1.json"))
r=d["roofline"]
print(f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["gemm_ms_per_step"]:.2f} ms ({r["achieved"]:.0f} TF/s)  attn {r["attention_ms_per_step"]:.2f}  vit {d["vit_crops_per_sec"]:.0f} crops/s  sm {d["clocks"]["sm_mhz"]} MHz')
print(json.dumps(d.get("secondary"), indent=1)[:3000])
print(json.dumps(d.get("cpu_baseline"), indent=1))
print(json.dumps(d.get("decode_step"), indent=1)[:1500])
PY
echo "== tail-split gemm validation"
SLIME_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gemm_tail_split_gpu.py -x -q -s -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_tail.log
