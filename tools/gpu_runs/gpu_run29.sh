#!/bin/bash
# decode step with PDL + deeper attention prefetch + KV append fused into the QKV epilogue: tests, decode bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -p no:cacheprovider --durations=5 -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_decode.log | cut -c1-220
timeout 600 python tools/bench_decode.py --batches 1,16 --steps 32 > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"; grep '"batch"' gpurun_out/decode_bench.log | cut -c1-420
