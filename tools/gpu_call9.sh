#!/bin/bash
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for o in '{}' '{"no_graph":1}'; do
  timeout 300 python bench.py --mesh hex --order 8 --rings 120 --size 1 1 100 --no-e2e --no-solve --steps 5 --opts "$o" > gpurun_out/bench_hex.json 2> gpurun_out/bench_hex.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_hex.json").read().strip().splitlines()[-1])
    print('$o', "| value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "kernel %.3f"%d["roofline"]["kernel_ms_per_launch"], "launches", d["roofline"]["launches_per_step"], d["gpu_launches"], "k", d["config"]["keff_after_steps"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_hex.err").read()[-1500:])
PY
done
python smoke_run.py 2>&1 | tail -2
