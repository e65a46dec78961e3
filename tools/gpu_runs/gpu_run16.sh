#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp16_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_fp16.log 2>&1; echo "pytest fp16 rc=$?"; grep -E "rel-L2|passed|failed|Error|error" gpurun_out/pytest_fp16.log | cut -c1-250 | tail -40
timeout 900 python -m pytest tests/test_shims_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_shims.log 2>&1; echo "pytest shims rc=$?"; tail -6 gpurun_out/pytest_shims.log | cut -c1-300
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -p no:cacheprovider -s -k "against_fp32" > gpurun_out/pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/pytest_fullsize.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 300 python tools/prof_preprocess.py > gpurun_out/prof_preprocess.log 2>&1; echo "prof_preprocess rc=$?"; tail -8 gpurun_out/prof_preprocess.log | cut -c1-400
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_all.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
