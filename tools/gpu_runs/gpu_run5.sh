#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -40 gpurun_out/pytest_attn.log | cut -c1-300
timeout 600 python -m pytest tests/test_shims_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_shims.log 2>&1; echo "pytest shims rc=$?"; tail -30 gpurun_out/pytest_shims.log | cut -c1-300
timeout 300 python tools/attn_diag.py > gpurun_out/attn_diag.log 2>&1; echo "attn_diag rc=$?"; tail -40 gpurun_out/attn_diag.log
