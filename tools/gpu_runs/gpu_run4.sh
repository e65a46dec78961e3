#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/pytest_all.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; tail -8 gpurun_out/gemm_diag.log
