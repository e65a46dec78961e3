#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -3 gpurun_out/pytest_attn.log | cut -c1-300
timeout 120 python tools/attn_trace.py > gpurun_out/attn_trace.log 2>&1; echo "attn_trace rc=$?"; tail -34 gpurun_out/attn_trace.log
timeout 200 python tools/attn_diag.py 2>&1 | tail -4
timeout 120 python tools/prof_gemm_vit.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 3 -c 1 -o gpurun_out/prof_gemm_vit -f python tools/prof_gemm_vit.py > gpurun_out/ncu_gemm_vit.out 2>&1; echo "ncu gemm vit rc=$?"
timeout 900 python bench.py --no-cpu-baseline --batch 32 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench b32 rc=$?"; cat gpurun_out/bench_b32.json | cut -c1-200
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-200
