#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -8 gpurun_out/pytest_attn.log | cut -c1-300
