#!/bin/bash
# decode attention A/B: mma kernel with 2 / 3 stages, with / without the PDL attribute; new merge kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_decode.log | cut -c1-220
for st in 3 2; do for pdl in 1 0; do
  echo "== mma stages=$st attn_pdl=$pdl"
  SLIME_DECODE_ATTN_STAGES=$st SLIME_DECODE_ATTN_PDL=$pdl timeout 300 python tools/bench_decode.py --batches 1,16 --steps 32 --no-projections --quick --out gpurun_out/decode_ab_${st}_${pdl}.json 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", f\"attn {d['profiled_ms_per_step']['attention']:.3f}\", d['kernels'][:70])
"
done; done
