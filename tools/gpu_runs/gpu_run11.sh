#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -3 gpurun_out/pytest_gemm.log | cut -c1-300
for rows in 1024 2048 4096 8192; do SLIME_GEMM_GROUP_ROWS=$rows timeout 120 python tools/prof_gemm.py; done
timeout 120 python tools/prof_gemm.py
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_2cta -s 3 -c 1 -o gpurun_out/prof_gemm2_default -f python tools/prof_gemm.py > gpurun_out/ncu_gemm2.out 2>&1; echo "ncu gemm default rc=$?"
SLIME_GEMM_GROUP_ROWS=1024 timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16_tn_2cta -s 3 -c 1 -o gpurun_out/prof_gemm2_rows1024 -f python tools/prof_gemm.py > gpurun_out/ncu_gemm2b.out 2>&1; echo "ncu gemm rows1024 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn_decoder -f python tools/prof_attn.py decoder > gpurun_out/ncu_attn.out 2>&1; echo "ncu attn decoder rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn_vit -f python tools/prof_attn.py vit > gpurun_out/ncu_attn2.out 2>&1; echo "ncu attn vit rc=$?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-200
SLIME_GEMM_GROUP_ROWS=1024 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_rows1024.json 2> gpurun_out/bench_b.err; echo "bench rows1024 rc=$?"; cat gpurun_out/bench_rows1024.json | cut -c1-200
ls -la gpurun_out | tail -12
