#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for gm in 4 8 2 1; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-solve --opts "{\"group_merge\":$gm}" > gpurun_out/bench_gm$gm.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_gm$gm.json").read().strip().splitlines()[-1])
    print("gm=$gm", "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "k", d["config"]["keff_after_steps"])
except Exception as e:
    print("gm=$gm failed", e); print(open("gpurun_out/bench_gm$gm.json").read()[-2000:])
PY
done
