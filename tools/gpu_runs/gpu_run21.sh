#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_preprocess_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess rc=$?"; tail -8 gpurun_out/pytest_pre.log | cut -c1-300
timeout 200 python tools/prof_preprocess.py > gpurun_out/prof_preprocess.log 2>&1; echo "prof_preprocess rc=$?"; tail -5 gpurun_out/prof_preprocess.log | cut -c1-400
