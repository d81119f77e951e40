#!/bin/bash
timeout 120 python tools/cmp_gm.py 64 64 216 '[{"wave_launch":1},{"group_merge":8},{"group_merge":8,"dbg":16},{"group_merge":8,"dbg":16},{"group_merge":4},{"group_merge":1,"dbg":16},{"group_merge":3,"store_psi":0}]' || echo "cmp failed/hung rc=$?"
for dbg in 0 16 64 128; do
  PAMPA_SN_DBG=$dbg timeout 120 python bench.py --no-cpu-baseline --no-e2e --no-solve > gpurun_out/bench_dbg$dbg.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_dbg$dbg.json").read().strip().splitlines()[-1])
    print("dbg=$dbg", "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms_per_launch"], "k", d["config"]["keff_after_steps"])
except Exception as e:
    print("dbg=$dbg failed", e); print(open("gpurun_out/bench_dbg$dbg.json").read()[-1500:])
PY
done
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
