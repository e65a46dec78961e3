#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stages_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_stages.log 2>&1; echo "pytest stages rc=$?"; tail -3 gpurun_out/pytest_stages.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
KREG='regex:gemm_bf16|attn_fwd|layernorm|rmsnorm|rope_|im2col|clip_embed|copy_rows|gate_mix|router_|splice_|text_|last_rows|add_rows'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 16 > gpurun_out/ncu_list.out 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 150 -c 4 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 16 > gpurun_out/ncu_full.out 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
