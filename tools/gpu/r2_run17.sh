#!/bin/bash
# round 2: compute-sanitizer passes over the whole path on a small configuration (smoke) and the kernel unit tests
mkdir -p gpurun_out
which compute-sanitizer || export PATH=$PATH:/usr/local/cuda/bin
echo "== memcheck: smoke()"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|\[smoke\]" gpurun_out/r2_sanitizer_memcheck_smoke.log | head -20
echo "== memcheck: kernel unit tests (GEMM, attention, decode kernels; small shapes)"
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_decode_kernels_gpu.py -q -m gpu -x -k "not fullsize and not 28672 and not 14336" > gpurun_out/r2_sanitizer_memcheck_kernels.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/r2_sanitizer_memcheck_kernels.log | head -20
echo "== initcheck: smoke()"
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_initcheck_smoke.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|Uninitialized|\[smoke\]" gpurun_out/r2_sanitizer_initcheck_smoke.log | head -20
echo "== racecheck: stage tests on the tiny configuration"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_stages_gpu.py -q -m gpu -x -k "tiny_spatial_b2" > gpurun_out/r2_sanitizer_racecheck_stages.log 2>&1; echo "rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/r2_sanitizer_racecheck_stages.log | head -20
