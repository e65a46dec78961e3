#!/bin/bash
# round 2: fused decode chain (in-kernel split-K finish / attention merge, RMSNorm in the staging) - tests + decode bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
echo "== decode kernel tests"
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -m gpu -x -s 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | tail -30 | tee gpurun_out/r2_decode_fused_tests.log
echo "== full-size decode parity + shims"
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_shims_gpu.py tests/test_graph_gpu.py -q -m gpu -x -k "decode or shim or cache or graph" -s 2>&1 | grep -E "rel-L2|passed|failed|Error|error" | tail -12
echo "== decode bench"
timeout 900 python tools/bench_decode.py --batches 1,4,16,32 --steps 32 --no-projections --out gpurun_out/r2_decode_bench.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: r = json.loads(line)
    except Exception: print(line.rstrip()[:200]); continue
    print(f\"B={r['batch']:>2} {r['ms_per_step']:.3f} ms  {r['achieved_gbs']:.0f} GB/s  frac {r['frac_of_hbm_peak']:.3f}  launches {r['launches_per_step']:.0f}  {r['kernels'][:70]}\")
"
