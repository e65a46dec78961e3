#!/bin/bash
# round 2: fused decode chain, second pass (batched RMSNorm staging) + which of the three fusions costs what
mkdir -p gpurun_out
echo "== decode kernel tests"
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -m gpu -x -s 2>&1 | grep -E "fused vs|passed|failed|Error|error|assert" | tail -12
echo "== decode bench"
timeout 900 python tools/bench_decode.py --batches 1,16 --steps 32 --variants 8 --no-projections --out gpurun_out/r2_decode_bench2.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: r = json.loads(line)
    except Exception: print(line.rstrip()[:200]); continue
    print(f\"B={r['batch']:>2} {r['ms_per_step']:.3f} ms  {r['achieved_gbs']:.0f} GB/s  frac {r['frac_of_hbm_peak']:.3f}  launches {r['launches_per_step']:.0f}  {r['kernels'][:90]}\")
"
