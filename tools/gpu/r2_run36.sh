#!/bin/bash
# round 2: 192-column cluster tiles for small problems - bit identity, invariants at full size, batch-1 latency A/B
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_graph_gpu.py tests/test_stages_gpu.py -q -m gpu -x 2>&1 | tail -2
SLIME_GEMM2_BN=192 timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_fullsize_gpu.py tests/test_fp16_gpu.py -q -m gpu -x 2>&1 | tail -2
echo "== batch-1 latency A/B (tools/step_profile-free: bench secondaries only)"
for bn in 256 0 256 0; do
SLIME_GEMM2_BN=$bn timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_bn.json 2> gpurun_out/r2_bench_bn.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_bn.json")); l=d.get("latency_b1",{})
print("tile width $bn (0 = by shape)", f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  B=1 {l.get("headline_llama3_8b_T256",{}).get("ms",0):.2f} ms / config2 {l.get("config2_vicuna7b_T128",{}).get("ms",0):.2f} ms  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
