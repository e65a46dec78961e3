#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/attn_diag.py > gpurun_out/attn_diag.log 2>&1; echo "attn_diag rc=$?"; tail -22 gpurun_out/attn_diag.log
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attention or activations or swiglu" > gpurun_out/pytest_attn.log 2>&1; rc=$?; echo "pytest attention rc=$rc"; tail -8 gpurun_out/pytest_attn.log | cut -c1-300
timeout 600 python -m pytest tests/test_variants_gpu.py tests/test_shims_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_var.log 2>&1; echo "pytest variants+shims rc=$?"; tail -8 gpurun_out/pytest_var.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-200
SLIME_ATTN_IMPL=p timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_pp.json 2> gpurun_out/bench_pp.err; echo "bench pingpong rc=$?"; cat gpurun_out/bench_pp.json | cut -c1-200
