#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn.log 2>&1; rc=$?; echo "pytest attention rc=$rc"; tail -5 gpurun_out/pytest_attn.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "attention failed: skipping the rest"; exit 1; fi
timeout 200 python tools/attn_diag.py > gpurun_out/attn_diag.log 2>&1; echo "attn_diag rc=$?"; tail -5 gpurun_out/attn_diag.log
timeout 600 python -m pytest tests/test_decode_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_decode.log 2>&1; echo "pytest decode rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/pytest_decode.log | head -20; tail -15 gpurun_out/pytest_decode.log | cut -c1-250
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_decode_gpu.py > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_all.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SLIME_ATTN_IMPL=fa2 SLIME_GEMM_2CTA=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err; echo "bench(old kernels) rc=$?"; cat gpurun_out/bench_old.json | cut -c1-300
