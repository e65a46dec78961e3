#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_all.log | cut -c1-300
timeout 300 python tools/attn_diag.py > gpurun_out/attn_diag.log 2>&1; echo "attn_diag rc=$?"; tail -5 gpurun_out/attn_diag.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
