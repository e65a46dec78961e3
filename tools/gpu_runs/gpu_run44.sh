#!/bin/bash
# N-GPU torchrun bench on the round's final state (weak scaling; N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1]); r=d["roofline"]
print(f'N={d["n_gpus"]} {d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  per-GPU {d["value"]/d["n_gpus"]:.0f}  gemm {r["achieved"]:.0f} TF/s  sm {d["clocks"]["sm_mhz"]} MHz  decode {d["decode_step"].get("tokens_per_s")}')
PY
