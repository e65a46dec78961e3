#!/bin/bash
# decode-step kernels (weight-streaming GEMM, split-KV attention): unit + decode tests, memcheck on a subset, decode bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py "tests/test_kernels_gpu.py::test_qkv_rope_fused_epilogue" "tests/test_kernels_gpu.py::test_gemm_plain" -q -p no:cacheprovider --durations=5 > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_decode.log | cut -c1-220
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_decode_kernels_gpu.py -q -p no:cacheprovider -k "fused_rmsnorm or rope_epilogue or (split_kv and 3-4-2) or (matches_fp32 and 136 and 17)" > gpurun_out/memcheck_decode.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/memcheck_decode.log; tail -4 gpurun_out/memcheck_decode.log | cut -c1-200
timeout 600 python tools/bench_decode.py --batches 1,16 --steps 32 > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"; tail -16 gpurun_out/decode_bench.log | cut -c1-600
