#!/bin/bash
# decode GEMM for M <= 8: only the real rows staged, 256-thread CTAs (two per SM) vs 512-thread CTAs; tests + bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_decode.log | cut -c1-220
run() {
  echo "== $1"
  env $1 timeout 300 python tools/bench_decode.py --batches 1,2,4,8 --steps 32 --no-projections --quick 2>&1 | grep '"batch"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", f\"attn {d['profiled_ms_per_step']['attention']:.3f} gemm {d['profiled_ms_per_step']['hbm_kernels']:.3f}\", d['kernels'][:70])
"
}
run "SLIME_SKINNY_SMALL_CTA=1"
run "SLIME_SKINNY_SMALL_CTA=0"
run "SLIME_SKINNY_SMALL_CTA=1 SLIME_CARVEOUT_PCT=30"
