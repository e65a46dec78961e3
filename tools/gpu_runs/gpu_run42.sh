#!/bin/bash
# decode attention: split merge done by the last-arriving CTA (no merge kernel) - tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py tests/test_fp16_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_decode.log | cut -c1-220
run() {
  echo "== $1"
  env $1 timeout 300 python tools/bench_decode.py --batches 1,4,16 --steps 32 --no-projections --quick 2>&1 | grep '"batch"' | grep mma | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", d['launches_per_step'], f\"attn {d['profiled_ms_per_step']['attention']:.3f}\")
"
}
run "SLIME_DECODE_ATTN_FUSED_MERGE=1"
run "SLIME_DECODE_ATTN_FUSED_MERGE=0"
run "SLIME_DECODE_ATTN_FUSED_MERGE=1"
run "SLIME_DECODE_ATTN_FUSED_MERGE=0"
