#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm2_diag.py > gpurun_out/gemm2_diag.log 2>&1; echo "gemm2_diag rc=$?"; tail -25 gpurun_out/gemm2_diag.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "2cta" > gpurun_out/pytest_2cta.log 2>&1; echo "pytest 2cta rc=$?"; tail -8 gpurun_out/pytest_2cta.log | cut -c1-300
SLIME_GEMM_2CTA=2 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err; echo "bench 2cta rc=$?"; cat gpurun_out/bench_2cta.json | cut -c1-400; tail -3 gpurun_out/bench_2cta.err
