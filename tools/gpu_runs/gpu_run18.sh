#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_diag.py > gpurun_out/attn_diag.log 2>&1; echo "attn_diag rc=$?"; tail -14 gpurun_out/attn_diag.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fp16_gpu.py tests/test_decode_gpu.py -q -m gpu -p no:cacheprovider -k "attention or decode or tiny" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -5 gpurun_out/pytest_attn.log | cut -c1-300
timeout 300 python tools/attn_trace.py > gpurun_out/attn_trace.log 2>&1; echo "trace rc=$?"; head -30 gpurun_out/attn_trace.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], r['attention_ms_per_step'], r['gemm_ms_per_step'], d['clocks'])"
