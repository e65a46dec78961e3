#!/bin/bash
# round 2: N=8 bench exactly as the driver launches it (gather_check, side-stream gather, NCCL log to stderr)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "rc=$?"
echo "stdout lines: $(wc -l < gpurun_out/r2_bench_n8.json)"
grep -c "NCCL INFO" gpurun_out/r2_bench_n8.err
grep -E "NCCL INFO (comm|ncclCommInitRank|Connected|NVLS|Channel 00)" gpurun_out/r2_bench_n8.err | head -8 | cut -c1-200
grep -v "NCCL INFO" gpurun_out/r2_bench_n8.err | tail -15
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(f'N={d["n_gpus"]} {d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gather_check={d.get("gather_check")}  sm {d["clocks"]["sm_mhz"]} MHz')
print(json.dumps(d.get("secondary"), indent=1)[:2500])
PY
