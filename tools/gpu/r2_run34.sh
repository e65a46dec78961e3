#!/bin/bash
# round 2: programmatic dependent launch for the prefill chain - bit-identity tests, then bench A/B (e2e and batch-1 latency
# are measured with the library's event profiler off; `value` has it on, which serialises the launches)
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -3
SLIME_PREFILL_PDL=1 timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_fullsize_gpu.py tests/test_variants_gpu.py tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== bench A/B"
for pdl in 0 1 0 1; do
SLIME_PREFILL_PDL=$pdl timeout 600 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/r2_bench_pdl.json 2> gpurun_out/r2_bench_pdl.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_pdl.json")); r=d["roofline"]; l=d.get("latency_b1",{})
print("prefill pdl $pdl", f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f} ({d["e2e"]["ms_per_step"]:.2f} ms)  B=1 {l.get("headline_llama3_8b_T256",{}).get("ms",0):.2f} ms / config2 {l.get("config2_vicuna7b_T128",{}).get("ms",0):.2f} ms  topp1 {d["secondary"].get("topp_1.0",{}).get("ms_per_step",0):.2f} ms  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
