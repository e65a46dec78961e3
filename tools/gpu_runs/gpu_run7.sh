#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shims_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_shims.log 2>&1; echo "pytest shims rc=$?"; tail -5 gpurun_out/pytest_shims.log | cut -c1-300
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/pytest_fullsize.log | head -20; tail -30 gpurun_out/pytest_fullsize.log | cut -c1-250
