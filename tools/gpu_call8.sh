#!/bin/bash
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python tools/cmp_gm.py 64 64 216 '[{"wave_launch":1},{},{"dbg":16},{"inline_edges":1},{"store_psi":0}]' || echo "cmp failed rc=$?"
for o in '{}' '{"inline_edges":1}' '{"store_psi":0}'; do
  timeout 200 python bench.py --no-cpu-baseline --no-solve --no-e2e --opts "$o" > gpurun_out/bench_t.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print('$o', "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "sweep %.2f"%d["roofline"]["sweep_ms_per_step"], "k", d["config"]["keff_after_steps"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_t.json").read()[-1500:])
PY
done
