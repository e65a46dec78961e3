#!/bin/bash
# first GPU call: kernel-level diagnostics + unit parity tests; every step bounded by its own timeout
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; echo "gemm_diag rc=$?"
for grp in gemm attention layernorm; do
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider > gpurun_out/pytest_$grp.log 2>&1; echo "pytest $grp rc=$?"
  tail -5 gpurun_out/pytest_$grp.log
done
tail -40 gpurun_out/gemm_diag.log
