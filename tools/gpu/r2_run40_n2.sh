#!/bin/bash
# round 2: N=2 A/B of the prefill PDL on one 2-GPU box (short: headline only)
mkdir -p gpurun_out
for pdl in 0 1 0 1; do
SLIME_PREFILL_PDL=$pdl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pdl \
  bench.py --gpus 2 --steps 8 --warmup 3 --no-secondary > gpurun_out/r2_bench_n2_pdl.json 2> gpurun_out/r2_bench_n2_pdl.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n2_pdl.json").read().strip().splitlines()[-1])
print("N=2 prefill pdl $pdl", f'{d["value"]:.0f} tok/s  e2e {d["e2e"]["value"]:.0f}  {d["ms_per_step"]:.2f} ms  e2e {d["e2e"]["ms_per_step"]:.2f} ms  gather_check={d.get("gather_check")}  sm {d["clocks"]["sm_mhz"]} MHz')
PY
done
