#!/bin/bash
# decode: swizzled 2-stage mma attention (fits the GEMM's carve-out), 8-row staging for M <= 8, carve-out knob
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_decode.log | cut -c1-220
run() {
  echo "== $1"
  env $1 timeout 300 python tools/bench_decode.py --batches 1,4,16 --steps 32 --no-projections --quick --out gpurun_out/decode_x.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", f\"attn {d['profiled_ms_per_step']['attention']:.3f} gemm {d['profiled_ms_per_step']['hbm_kernels']:.3f}\", d['kernels'][:70])
"
}
run "SLIME_X=0"
run "SLIME_SKINNY_ROWS8=0"
run "SLIME_CARVEOUT_PCT=58"
run "SLIME_CARVEOUT_PCT=44"
