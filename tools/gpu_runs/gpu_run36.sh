#!/bin/bash
# decode: preferred carve-out sweep for the chain kernels
mkdir -p gpurun_out
run() {
  echo "== $1"
  env $1 timeout 300 python tools/bench_decode.py --batches 1,16 --steps 32 --no-projections --quick --out gpurun_out/decode_x.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", f\"attn {d['profiled_ms_per_step']['attention']:.3f} gemm {d['profiled_ms_per_step']['hbm_kernels']:.3f}\", d['kernels'][:70])
"
}
run "SLIME_CARVEOUT_PCT=30"
run "SLIME_CARVEOUT_PCT=16"
run "SLIME_CARVEOUT_PCT=0"
