#!/bin/bash
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for o in '{}' '{"group_merge":8}'; do
  timeout 200 python bench.py --no-cpu-baseline --no-solve --opts "$o" > gpurun_out/bench_t.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print('$o', "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "k", d["config"]["keff_after_steps"], "e2e %.4g"%d["e2e"]["value"], d["e2e"]["phases_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_t.json").read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches3.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-solve > gpurun_out/launches3.log 2>&1
grep -o '"sn_[a-z_]*\|void sn_[a-z_<0-9, >]*\|"ns","[0-9]*"' gpurun_out/launches3.csv | paste - - | tail -9
