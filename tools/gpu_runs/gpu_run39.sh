#!/bin/bash
# N=2 torchrun bench on the round's final state (weak scaling, logits all-gather, decode figure on every rank)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n2.json") if l.startswith("{")][-1]); r=d["roofline"]
print(f'N={d["n_gpus"]} {d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  gemm {r["achieved"]:.0f} TF/s  sm {d["clocks"]["sm_mhz"]} MHz decode {d.get("decode_step")}')
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
