#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_stages.log 2>&1; echo "pytest stages rc=$?"
grep -E "rel-L2|passed|failed|Error|error" gpurun_out/pytest_stages.log | head -80
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
