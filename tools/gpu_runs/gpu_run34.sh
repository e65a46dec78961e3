#!/bin/bash
# decode chain with one shared-memory carve-out for all kernels (SLIME_CARVEOUT=1) vs the driver's per-kernel choice
mkdir -p gpurun_out
for co in 1 0; do for st in 3 2; do
  echo "== carveout=$co mma stages=$st"
  SLIME_CARVEOUT=$co SLIME_DECODE_ATTN_STAGES=$st timeout 300 python tools/bench_decode.py --batches 1,16 --steps 32 --no-projections --quick --out gpurun_out/decode_co_${co}_${st}.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", f\"attn {d['profiled_ms_per_step']['attention']:.3f}\", d['kernels'][:70])
"
done; done
