#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_pre.log 2>&1; echo "pytest preprocess rc=$?"; tail -15 gpurun_out/pytest_pre.log | cut -c1-300
timeout 300 python tools/prof_preprocess.py > gpurun_out/prof_preprocess.log 2>&1; echo "prof_preprocess rc=$?"; tail -8 gpurun_out/prof_preprocess.log | cut -c1-400
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_preprocess_gpu.py > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_all.log | cut -c1-300
