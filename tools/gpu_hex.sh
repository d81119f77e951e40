#!/bin/bash
for args in "--order 8 --rings 120 --size 1 1 100" "--order 12 --rings 80 --size 1 1 100 --groups 16"; do
  timeout 300 python bench.py --mesh hex $args --no-e2e --steps 5 > gpurun_out/bench_hex.json 2> gpurun_out/bench_hex.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_hex.json").read().strip().splitlines()[-1])
    print(d["config"]["workload"], "| value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "launches", d["roofline"]["launches_per_step"], "solve", d["keff_solve"], d["config"]["l2_policy"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_hex.json").read()[-1500:]); print(open("gpurun_out/bench_hex.err").read()[-1500:])
PY
  cp gpurun_out/bench_hex.json "gpurun_out/bench_hex_$(echo $args | tr ' ' '_' | tr -d '-').json"
done
