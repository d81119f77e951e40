#!/bin/bash
for zc in 16 8 4; do
  timeout 300 python bench.py --mesh hex --order 8 --rings 120 --size 1 1 100 --no-e2e --no-solve --steps 5 --opts "{\"z_chunk\":$zc}" > gpurun_out/bench_hex.json 2> gpurun_out/bench_hex.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_hex.json").read().strip().splitlines()[-1])
    print("z_chunk $zc | value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "launches", d["roofline"]["launches_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_hex.err").read()[-1500:])
PY
done
