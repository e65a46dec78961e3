#!/bin/bash
# Run A of session 2: verify HEAD on a fresh box (full -m gpu suite), default bench line, refreshed ncu launch list
# (current kernels: 2-CTA GEMM + tcgen05 attention), and ncu --set full captures of the HBM-bound kernels.
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
KREG='regex:gemm_bf16|attn_|layernorm|rmsnorm|rope_|im2col|clip_embed|copy_rows|gate_mix|router_|splice_|text_|last_rows|add_rows|scatter_rows|cache_rows|resize_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.out 2>&1; echo "ncu list rc=$?"
SMALL='regex:rmsnorm|rope_kernel|clip_embed|im2col|splice_gather|router_score|router_select|text_dir|text_inv|gate_mix|copy_rows|splice_plan'
timeout 400 ncu --set full --clock-control none --import-source on -k "$SMALL" -c 26 -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 1 > gpurun_out/ncu_small.out 2>&1; echo "ncu small rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 6 -c 2 -o gpurun_out/prof_layernorm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 1 > gpurun_out/ncu_ln.out 2>&1; echo "ncu layernorm rc=$?"
ls -la gpurun_out
