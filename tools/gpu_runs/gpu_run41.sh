#!/bin/bash
# qformer (cross-attention) text-guided router: GPU parity tests + regression of the router / stage / shim suites
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qformer_router_gpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_qformer.log 2>&1; echo "qformer rc=$?"; grep -E "rel-L2|passed|failed|Error|error|assert" gpurun_out/pytest_qformer.log | head -30
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_variants_gpu.py tests/test_shims_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_regr.log 2>&1; echo "regression rc=$?"; tail -3 gpurun_out/pytest_regr.log
