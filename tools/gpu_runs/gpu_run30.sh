#!/bin/bash
# decode step: L2 prefetch duties + both register groups before the PDL wait; tests, decode bench A/B, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_kernels_gpu.py tests/test_decode_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decode.log | cut -c1-220
timeout 600 python tools/bench_decode.py --batches 1,16 --steps 32 --no-projections > gpurun_out/decode_bench.log 2>&1; echo "decode bench rc=$?"; grep '"batch"' gpurun_out/decode_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['batch'], f\"{d['ms_per_step']:.3f} ms\", f\"{d['frac_of_hbm_peak']:.3f}\", d['launches_per_step'], d['kernels'])
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/decode_launches.csv python tools/bench_decode.py --layers 2 --batches 1 --steps 1 --no-projections > gpurun_out/decode_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/decode_launches.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-24:]:
    print(r[4][:60].ljust(60), r[-1], r[-2])
PY
